"""Research prototype (CPU only, not part of the product): an exact INTEGER model of the reference's compensated
sum (Solution::sumInArea, /root/reference/src/Quantizer.cpp:59-70) on SCALED lattice values t/255.0, as the basis
of a parallel (chunked) bit-exact centroid sum.  See DESIGN.md 4.6 "what would make exactness cheap".

Model.  g = 2^-60 (the ulp of the smallest non-zero lattice value 1/255).  X_t = fl(t/255)/g is an integer, a multiple
of u_t = ulp(fl(t/255))/g = 2^(floor(log2 t)) for 1 <= t <= 254 and 2^8 for t = 255.  Once sum >= 2 (so that Fast2Sum is
exact and ulp(sum) >= 2 ulp(y)) the pair (sum, c) is a function of ONE exact integer A = (sum - c)/g:
        sum = RN53(A),  c = sum - A
and one loop iteration with addend t is
        t == 0        : A unchanged
        1 <= t <= 254 : A <- X_t + rm(A, u_t)          rm = round to a multiple of u_t, ties so that (X_t + rm)/u_t is even
        t == 255      : the grid of y = fl(1 - c) is 2^8 if c < 0 (y > 1), 2^7 if c > 0 (y < 1), exact if c == 0
The final sum is RN53(A).  Everything below ulp(sum) matters only through A mod 2^9 - except the sign test of t = 255.
This file checks the model against the real floating-point loop; the chunked evaluation is in kahan_chunks.py."""
import numpy as np

G = 60


def lattice_tables():
    X, U = [0] * 256, [0] * 256
    for t in range(1, 256):
        x = float(t) / 255.0
        m, e = np.frexp(x)                    # x = m * 2^e, 0.5 <= m < 1
        X[t] = int(x * 2.0 ** G)              # exact: x has 53 bits, the product is an integer < 2^61
        assert X[t] * 2.0 ** -G == x
        U[t] = 1 << (int(e) - 1 - 52 + G)     # ulp(x) / g
        assert X[t] % U[t] == 0
    return X, U


X_T, U_T = lattice_tables()


def kahan_fp(ts, s=0.0, c=0.0):
    """The reference's loop in IEEE double (numpy scalars; no FMA)."""
    s, c = np.float64(s), np.float64(c)
    for t in ts:
        x = np.float64(t) / np.float64(255.0)
        y = x - c
        tt = s + y
        c = (tt - s) - y
        s = tt
    return float(s), float(c)


def rn53(A):
    """Round the non-negative integer A (units of g) to 53 significant bits, ties to even. Returns (rounded, U)."""
    if A == 0:
        return 0, 1
    n = A.bit_length()
    if n <= 53:
        return A, 1
    U = 1 << (n - 53)
    q, r = divmod(A, U)
    if r * 2 > U or (r * 2 == U and (q & 1)):
        q += 1
    return q * U, U          # (q may have become 2^53: still exact as a double)


def rm_even(A, u, X):
    """A rounded to a multiple of u; on a tie pick the neighbour that makes (X + result)/u even."""
    r = A % u
    lo = A - r
    if r * 2 < u:
        return lo
    if r * 2 > u:
        return lo + u
    return lo if ((X + lo) // u) % 2 == 0 else lo + u


def step_int(A, t):
    if t == 0:
        return A
    X = X_T[t]
    if t < 255:
        return X + rm_even(A, U_T[t], X)
    s, _ = rn53(A)
    d = A - s                         # = -c / g
    if d == 0:
        return X + A                  # y = 1 exactly... A' = sum + y = A + X
    u = 1 << 8 if d > 0 else 1 << 7   # y = fl(1 + d): above 1 the grid is 2^-52, below 1 it is 2^-53
    # y = fl(X + d) on that grid, ties to even mantissa; A' = sum + y = (A - d) + y
    v = X + d
    r = v % u
    lo = v - r
    if r * 2 < u:
        y = lo
    elif r * 2 > u:
        y = lo + u
    else:
        y = lo if (lo // u) % 2 == 0 else lo + u
    return (A - d) + y


def kahan_int(ts, A):
    for t in ts:
        A = step_int(A, t)
    return A


def to_A(s, c):
    """(sum, c) doubles -> the exact integer A = (sum - c)/g."""
    from fractions import Fraction
    v = (Fraction(s) - Fraction(c)) * (1 << G)
    assert v.denominator == 1
    return int(v)


def check(n_trials=300, seed=1):
    rng = np.random.default_rng(seed)
    bad = 0
    for trial in range(n_trials):
        kind = trial % 6
        n = int(rng.integers(50, 3000))
        if kind == 0:
            ts = rng.integers(0, 256, n)
        elif kind == 1:
            ts = rng.integers(0, 8, n)                     # dark: tiny ulps
        elif kind == 2:
            ts = rng.choice([255, 254, 1, 0, 128, 127], n)    # binade edges and exact 1.0
        elif kind == 3:
            ts = np.full(n, int(rng.integers(1, 256)))
        elif kind == 4:
            ts = rng.choice([255, 3], n)
        else:
            ts = rng.integers(120, 136, n)
        # prefix in floating point until the regime holds (sum >= 2 and at least a few terms)
        s = c = 0.0
        k = 0
        while k < n and s < 4.0:
            s, c = kahan_fp(ts[k:k + 1], s, c)
            k += 1
        A = to_A(s, c)
        assert rn53(A)[0] * 2.0 ** -G == s
        # step by step comparison
        for j in range(k, n):
            s, c = kahan_fp(ts[j:j + 1], s, c)
            A = step_int(A, int(ts[j]))
            if to_A(s, c) != A:
                bad += 1
                print(f"trial {trial} kind {kind}: diverged at element {j} (t={ts[j]}), sum={s!r}")
                break
    print(f"{n_trials} chains, {bad} diverged from the floating-point loop")
    return bad


if __name__ == "__main__":
    raise SystemExit(1 if check() else 0)
