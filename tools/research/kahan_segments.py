"""Research prototype (CPU only): the chunked evaluation of kahan_chunks.py with FOUR speculative branches instead of 512.

After an addend with 128 <= t <= 254 (ulp 2^7 in units of g) the state A is a multiple of 128, so A mod 512 is one of
{0, 128, 256, 384}.  Cutting a chain at such ANCHORS therefore leaves four possible entry residues per segment, and the
four runs differ only where bit 7 or 8 of the state matters (ties at ulp-2^7 addends, addends t = 255); everything else
is shared.  A GPU version would: (1) pick an anchor every ~C members per (cell, dimension) chain, (2) sum the X of
every segment exactly (128-bit) and prefix them, (3) simulate every segment for the four residues in parallel, recording
the margin of its t = 255 decisions, (4) chain the segment summaries sequentially (one short loop per chain), re-running
the rare segment whose margin does not cover the true start.  Chains (or stretches) without anchors - dark regions -
stay sequential.  This file checks the scheme against the plain integer model."""
import numpy as np

from kahan_automaton import X_T, kahan_fp, kahan_int, step_int, to_A
from kahan_chunks import M, centered, decision_margin

RESIDUES = (0, 128, 256, 384)


def pick_anchors(ts, spacing):
    """Indices i such that the segment boundary sits right AFTER element i (128 <= ts[i] <= 254), >= spacing apart."""
    out, last = [], -spacing
    for i, t in enumerate(ts):
        if 128 <= t <= 254 and i - last >= spacing:
            out.append(i)
            last = i
    return out


def summarize_segment(part, E_start):
    """E_start: the no-rounding trajectory value right after the anchor (a multiple of 128 is NOT required of it)."""
    table = {}
    for r in RESIDUES:
        A = E_start + centered(r - E_start)
        A0, margin = A, None
        for t in part:
            t = int(t)
            if t == 255:
                dm = decision_margin(A)
                margin = dm if margin is None else min(margin, dm)
            A = step_int(A, t)
        table[r] = (A0, A - A0, margin)
    return table


def evaluate(ts, A, spacing):
    anchors = pick_anchors(ts, spacing)
    if not anchors:
        return kahan_int([int(t) for t in ts], A), 0, 0
    # head: up to and including the first anchor, sequentially
    first = anchors[0]
    A = kahan_int([int(t) for t in ts[:first + 1]], A)
    assert A % 128 == 0
    E = A
    bounds = anchors[1:] + [len(ts) - 1]
    start, reruns = first + 1, 0
    for end in bounds:                                   # segments (start .. end], in parallel on a GPU
        part = ts[start:end + 1]
        if len(part) == 0:
            continue
        A_spec, delta, margin = summarize_segment(part, E)[A % M]
        shift = A - A_spec
        assert shift % M == 0
        if margin is None or abs(shift) < margin:
            A += delta
        else:
            A = kahan_int([int(t) for t in part], A)
            reruns += 1
        E += sum(X_T[int(t)] for t in part)
        start = end + 1
    return A, reruns, len(bounds)


def check(seed=4, trials=24):
    rng = np.random.default_rng(seed)
    bad = reruns = segs = 0
    for trial in range(trials):
        n, spacing = int(rng.integers(3000, 12000)), int(rng.choice([32, 128, 512]))
        kind = trial % 4
        ts = [rng.integers(0, 256, n), rng.choice([255, 254, 0, 1, 2, 200], n),
              np.where(rng.random(n) < 0.03, rng.integers(128, 255, n), rng.integers(0, 40, n)),   # dark with rare anchors
              rng.choice([255, 128, 127, 64, 63], n)][kind]
        s = c = 0.0
        k = 0
        while k < n and s < 4.0:
            s, c = kahan_fp(ts[k:k + 1], s, c)
            k += 1
        A0 = to_A(s, c)
        want = kahan_int([int(t) for t in ts[k:]], A0)
        got, r, sg = evaluate(ts[k:], A0, spacing)
        reruns += r
        segs += sg
        if got != want:
            bad += 1
            print(f"trial {trial} kind {kind} spacing {spacing}: MISMATCH")
    print(f"{trials} chains: {bad} mismatches; {reruns} of {segs} segments re-run sequentially")
    return bad


if __name__ == "__main__":
    raise SystemExit(1 if check() else 0)
