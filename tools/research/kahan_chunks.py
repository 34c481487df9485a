"""Research prototype (CPU only): chunked evaluation of the integer model in kahan_automaton.py.

A chunk's effect on A is A_end - A_start = sum(X) + (rounding corrections), and the corrections depend on A_start only
through  l0 = A_start mod 512  - except at addends t = 255, whose grid depends on the sign of c, i.e. on where A sits
inside its own ulp.  So every chunk can be summarised INDEPENDENTLY (in parallel on a GPU) for each of the 512 values
of l0 from a speculative start A_spec = E_start + (small correction congruent to l0): the summary stays valid for the
true start A_spec + 512 m as long as |512 m| is below the smallest distance of any t = 255 decision to its threshold
(the `margin`).  A cheap sequential pass then chains the summaries and re-runs the few chunks whose margin is too
small (start of a chain, where ulp(sum) is comparable to the accumulated corrections).
The 512 speculative runs collapse quickly in practice (after an addend with ulp 2^b only the bits >= b of the state
survive), which is what a GPU version would exploit; this prototype only demonstrates exactness."""
import numpy as np

from kahan_automaton import G, X_T, kahan_fp, kahan_int, rn53, step_int, to_A

M = 512


def centered(v):
    v %= M
    return v - M if v > M // 2 else v


def decision_margin(A):
    """Distance from A to the nearest point where the t = 255 branch (sign of A - RN53(A), or ulp(A)) changes."""
    n = A.bit_length()
    U = 1 << max(n - 53, 0)
    r = A % U
    d = min(r, abs(r - U // 2), U - r) if U > 1 else 0
    d = min(d, A - (1 << (n - 1)), (1 << n) - A)      # binade edges
    return d


def summarize_chunk(ts, E_start):
    """For every l0: (A_spec, A_end_spec - A_spec, margin)."""
    table = []
    for l0 in range(M):
        A = E_start + centered(l0 - E_start)
        assert A % M == l0
        A0, margin = A, None
        for t in ts:
            t = int(t)
            if t == 255:
                dm = decision_margin(A)
                margin = dm if margin is None else min(margin, dm)
            A = step_int(A, t)
        table.append((A0, A - A0, margin))
    return table


def chained(ts, A, chunk):
    """Exact A after all of ts, using per-chunk summaries built without knowledge of the true A (only E)."""
    E = A                                # the trajectory without any rounding correction
    reruns = 0
    for a in range(0, len(ts), chunk):
        part = ts[a:a + chunk]
        table = summarize_chunk(part, E)              # <- parallel over chunks on a GPU: depends on E only
        A_spec, delta, margin = table[A % M]
        shift = A - A_spec
        assert shift % M == 0
        if margin is None or abs(shift) < margin:
            A = A + delta
        else:                                          # the speculation does not cover the true start: run it
            A = kahan_int([int(t) for t in part], A)
            reruns += 1
        E += sum(X_T[int(t)] for t in part)
    return A, reruns


def check(seed=3):
    rng = np.random.default_rng(seed)
    bad = total_reruns = total_chunks = 0
    for trial in range(24):
        n, chunk = int(rng.integers(2000, 9000)), int(rng.choice([64, 256, 1000]))
        kind = trial % 4
        ts = [rng.integers(0, 256, n), rng.choice([255, 254, 0, 1, 2, 200], n), rng.integers(0, 40, n),
              rng.choice([255, 128, 127, 64, 63], n)][kind]
        s = c = 0.0
        k = 0
        while k < n and s < 4.0:
            s, c = kahan_fp(ts[k:k + 1], s, c)
            k += 1
        A0 = to_A(s, c)
        want = kahan_int([int(t) for t in ts[k:]], A0)
        s_fp, c_fp = kahan_fp(ts[k:], s, c)
        assert to_A(s_fp, c_fp) == want
        got, reruns = chained(ts[k:], A0, chunk)
        total_reruns += reruns
        total_chunks += (n - k + chunk - 1) // chunk
        if got != want:
            bad += 1
            print(f"trial {trial} kind {kind} chunk {chunk}: MISMATCH")
        assert rn53(got)[0] * 2.0 ** -G == s_fp or got != want
    print(f"24 chains: {bad} mismatches; {total_reruns} of {total_chunks} chunks had to be re-run sequentially")
    return bad


if __name__ == "__main__":
    raise SystemExit(1 if check() else 0)
