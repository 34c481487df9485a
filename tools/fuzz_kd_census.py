"""CPU fuzz of the census behind the auto centroid mode (no GPU needed):  python tools/fuzz_kd_census.py [seconds] [seed]

The auto mode keeps a train's integer-sum centroids only if no decision of the train can depend on the last bits in
which those centroids differ from the reference's compensated sums.  Two pieces of that reasoning live in
host code and in a few lines of the resolver that are restated here:
  1. KdHostTree::min_margin (kd_host.cpp): above 1e-9 the KD tree built from the integer-sum codebook must have the
     SHAPE and point order of the tree the reference builds from its own codebook;
  2. the visiting order among exactly tied, bit-reproducible candidates (resolve_bruteforce_kernel): "order_safe" must
     imply that the first tied candidate the reference's walk visits is the one our descent picks.
The fuzz builds codebooks the way a split level does - cells of 8-bit vectors; centroid = ((sum t)/255)/n (ours) or the
compensated sum of t/255 divided by n (reference); children 1.2c | 0.8c - with palettes small enough that spreads, planes
and distances tie all the time, builds both trees through libqb200 (qb200_debug_kd_tree) and checks 1. and 2.
Exit code 1 on any violation (the failing codebook is written to /tmp/kd_census_fail_<n>.npz)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quant_b200 import _lib

NODE = np.dtype([("c1", "<i4"), ("c2", "<i4"), ("a", "<i4"), ("b", "<i4"), ("lo", "<f8"), ("hi", "<f8")])
FEAT_MASK, LOW_EXACT, HIGH_EXACT, NODE_FRAGILE = 0xFFFF, 1 << 16, 1 << 17, 1 << 18
PER_PATH = os.environ.get("KD_FUZZ_PER_PATH", "1") == "1"   # 0: the one-margin-per-level rule the per-path rule replaced
SMALL = 8
IGNORE_BITS = os.environ.get("KD_FUZZ_IGNORE_PLANE_BITS") == "1"   # self-test: the census as it was before the plane bits


def build(lib, pts, flags):
    pts = np.ascontiguousarray(pts, np.float64)
    K, dim = pts.shape
    nodes = np.zeros(2 * K + 8, NODE)
    order = np.zeros(K, np.uint32)
    n, m = C.c_int(), C.c_double()
    f = None if flags is None else np.ascontiguousarray(flags, np.uint8).ctypes.data_as(C.c_void_p)
    rc = lib.qb200_debug_kd_tree(pts.ctypes.data_as(C.c_void_p), K, dim, f, nodes.ctypes.data_as(C.c_void_p), nodes.size,
                                 C.byref(n), order.ctypes.data_as(C.c_void_p), C.byref(m))
    assert rc == 0
    return nodes[:n.value], order, m.value


def kahan_mean(ts):
    """The reference's centroid of a cell: compensated sum of t/255 in member order, divided by the count."""
    s = np.zeros(ts.shape[1])
    c = np.zeros(ts.shape[1])
    for t in ts:
        x = t.astype(np.float64) / 255.0
        y = x - c
        tt = s + y
        c = (tt - s) - y
        s = tt
    return s / float(len(ts))


def make_level(rng):
    """One split level: (ours, reference, flags) for 2 * cells codevectors."""
    dim = int(rng.choice([3, 3, 6, 9, 12, 12, 27, 48]))
    cells = int(rng.choice([8, 16, 32, 64, 128, 256]))
    n_colours = int(rng.choice([2, 3, 4, 6, 16, 256]))
    palette = np.sort(rng.choice(256, n_colours, replace=False)) if n_colours < 256 else np.arange(256)
    ours = np.zeros((cells, dim))
    ref = np.zeros((cells, dim))
    flag = np.zeros(cells, np.uint8)
    shared = palette[rng.integers(0, len(palette), dim)]   # a vector many cells are built around
    flat = rng.random() < 0.35
    for k in range(cells):
        kind = rng.random()
        if kind < 0.08:                                   # dead cell
            flag[k] = 1
            continue
        n = int(rng.choice([1, 1, 2, 2, 3, 5, 8, 9, 10, 12, 16, 40, 150]))
        ts = palette[rng.integers(0, len(palette), (n, dim))].astype(np.int64)
        if flat:                                          # a flat image with outliers: copies of one vector, few strays
            stray = ts.copy()
            ts[:] = shared
            for m in range(n):
                if rng.random() < (0.5 if n == 1 else 0.12):
                    px = int(rng.integers(0, dim // 3)) * 3   # one stray pixel (3 coordinates)
                    ts[m, px:px + 3] = stray[m, px:px + 3]
        elif rng.random() < 0.5:                          # most coordinates equal to the shared vector's
            keep = rng.random(dim) < 0.7
            ts[:, keep] = shared[keep]
        if rng.random() < 0.2:
            ts[:] = ts[0]                                 # all members one vector
        r = kahan_mean(ts)
        if n <= SMALL:
            o = r.copy()
            flag[k] = 1
        elif (ts == ts[0]).all():
            o = (float(n) * (ts[0].astype(np.float64) / 255.0)) / float(n)   # closed form of the one-vector cell
            assert np.array_equal(o, r), "closed form of a one-vector cell differs from the compensated sum"
            flag[k] = 1
        else:
            o = (ts.sum(0).astype(np.float64) / 255.0) / float(n)
        ours[k], ref[k] = o, r
    f_up, f_dn = 1 + 0.2, 1 - 0.2
    return (np.concatenate([ours * f_up, ours * f_dn]), np.concatenate([ref * f_up, ref * f_dn]),
            np.concatenate([flag, flag]), palette, ours)


def same_shape(a, b, oa, ob):
    return (len(a) == len(b) and np.array_equal(a["c1"], b["c1"]) and np.array_equal(a["c2"], b["c2"])
            and np.array_equal(a["a"] & FEAT_MASK, b["a"] & FEAT_MASK) and np.array_equal(a["b"], b["b"])
            and np.array_equal(oa, ob))


def l2(x, c):
    d = x - c
    return float(np.add.reduce(d * d))     # any fixed order: the tied candidates are the same numbers in both codebooks


def first_visited(nodes, order, x, cand_pos):
    """Position (in `order`) of the first candidate a near-child-first depth-first walk reaches."""
    stack = [0]
    while stack:
        i = stack.pop()
        nd = nodes[i]
        if nd["c1"] < 0:
            for p in range(nd["a"], nd["b"]):
                if p in cand_pos:
                    return p
            continue
        val = x[nd["a"] & FEAT_MASK]
        side = (val - nd["lo"]) + (val - nd["hi"])
        near, far = (nd["c1"], nd["c2"]) if side < 0 else (nd["c2"], nd["c1"])
        stack.append(far)
        stack.append(near)
    return -1


def census_descent(nodes, order, inv, pts, x, cands):
    """resolve_bruteforce_kernel's descent among exactly tied candidates: (winner, fragile descent step, fragile node
    on the path from the root to the winner's leaf)."""
    pos = {int(inv[k]) for k in cands}
    node, lo, hi, fragile, path_fragile = 0, 0, len(order), False, False
    while True:
        nd = nodes[node]
        if nd["c1"] < 0:
            break
        path_fragile = path_fragile or bool(nd["a"] & NODE_FRAGILE)
        mid = int(nd["b"])
        in1 = {p for p in pos if lo <= p < mid}
        in2 = {p for p in pos if mid <= p < hi}
        if in1 and in2:
            feat = int(nd["a"]) & FEAT_MASK
            val = x[feat]
            side = (val - nd["lo"]) + (val - nd["hi"])
            go1 = side < 0
            if abs(side) <= 1e-9 * (abs(val) + abs(nd["lo"]) + abs(nd["hi"])):
                lo_ok = (IGNORE_BITS or bool(nd["a"] & LOW_EXACT)) and any(pts[order[p], feat] == nd["lo"] for p in in1)
                hi_ok = (IGNORE_BITS or bool(nd["a"] & HIGH_EXACT)) and any(pts[order[p], feat] == nd["hi"] for p in in2)
                fragile = fragile or not (lo_ok and hi_ok)
        else:
            go1 = bool(in1)
        if go1:
            node, hi, pos = int(nd["c1"]), mid, in1
        else:
            node, lo, pos = int(nd["c2"]), mid, in2
    return int(order[min(pos)]), fragile, path_fragile


def our_level(L, assign, K_prev):
    """(codebook of the next split level in OUR arithmetic, reproducibility flags) from a level's assignment.
    L: N x dim values t = 0..255 (SCALED value t / 255); assign: index per vector into K_prev cells."""
    dim = L.shape[1]
    cent = np.zeros((K_prev, dim))
    flag = np.zeros(K_prev, np.uint8)
    order = np.argsort(assign, kind="stable")
    bounds = np.searchsorted(assign[order], np.arange(K_prev + 1))
    for k in range(K_prev):
        mem = L[order[bounds[k]:bounds[k + 1]]]          # members in ascending vector order
        n = len(mem)
        if n == 0:
            flag[k] = 1
        elif n <= SMALL:
            cent[k] = kahan_mean(mem)
            flag[k] = 1
        elif (mem == mem[0]).all():
            cent[k] = (float(n) * (mem[0].astype(np.float64) / 255.0)) / float(n)
            flag[k] = 1
        else:
            cent[k] = (mem.sum(0).astype(np.float64) / 255.0) / float(n)
    f_up, f_dn = 1 + 0.2, 1 - 0.2
    return np.concatenate([cent * f_up, cent * f_dn]), np.concatenate([flag, flag])


def check_level(lib, ours, flags, ref, queries, stats=None):
    """The three census checks on one level; returns (robust, order-safe ties checked) or raises AssertionError."""
    same_bits = (ours.view(np.uint64) == ref.view(np.uint64)).all(axis=1)
    assert same_bits[flags == 1].all(), "a codevector flagged reproducible differs from the reference's bits"
    no, oo, margin = build(lib, ours, flags)
    nr, orr, _ = build(lib, ref, None)
    robust = margin > 1e-9
    if robust:
        assert same_shape(no, nr, oo, orr), f"census margin {margin:.3g} but the trees differ"
    K = ours.shape[0]
    inv_o = np.empty(K, np.int64)
    inv_o[oo] = np.arange(K)
    inv_r = np.empty(K, np.int64)
    inv_r[orr] = np.arange(K)
    d = ((queries[:, None, :] - ours[None, :, :]) ** 2).sum(2)
    dmin, dmax = d.min(1), d.max(1)
    band = d <= (dmin + dmax * 5.6843418860808015e-14)[:, None]
    ties = 0
    for q in np.nonzero(band.sum(1) > 1)[0]:
        cands = np.nonzero(d[q] == dmin[q])[0]
        in_band = np.nonzero(band[q])[0]
        if len(cands) < 2 or len(cands) > 32 or len(cands) != len(in_band) or not flags[in_band].all():
            continue
        win, fragile, path_fragile = census_descent(no, oo, inv_o, ours, queries[q], cands)
        if stats is not None:   # ties the per-level rule / the per-path rule would send to the exact repeat
            stats["unsafe_per_level"] = stats.get("unsafe_per_level", 0) + int(fragile or not robust)
            stats["unsafe_per_path"] = stats.get("unsafe_per_path", 0) + int(fragile or not (robust or not path_fragile))
        if fragile or not (robust or (PER_PATH and not path_fragile)):
            continue
        dr = np.array([l2(queries[q], c) for c in ref])
        cr = np.nonzero(dr == dr.min())[0]
        ref_win = int(orr[first_visited(nr, orr, queries[q], {int(inv_r[k]) for k in cr})])
        assert ref_win == win, f"tie among {cands.tolist()} called safe: ours {win}, reference {ref_win}"
        ties += 1
    return robust, ties


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lib = _lib.load()
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    levels = robust = shape_bad = ties = tie_safe = tie_bad = 0
    while time.time() < t_end:
        ours, ref, flags, palette, centroids = make_level(rng)
        K, dim = ours.shape
        no, oo, margin = build(lib, ours, flags)
        nr, orr, _ = build(lib, ref, None)
        levels += 1
        is_robust = margin > 1e-9
        if is_robust:
            robust += 1
            if not same_shape(no, nr, oo, orr):
                shape_bad += 1
                np.savez(f"/tmp/kd_census_fail_{shape_bad + tie_bad}.npz", ours=ours, ref=ref, flags=flags)
                print(f"SHAPE differs although the census reports margin {margin:.3g}: K={K} dim={dim}", flush=True)
        inv_o = np.empty(K, np.int64)
        inv_o[oo] = np.arange(K)
        inv_r = np.empty(K, np.int64)
        inv_r[orr] = np.arange(K)
        # queries: lattice vectors from the palette (members of the cells), looking for exact ties of the minimum
        for _ in range(24):
            x = palette[rng.integers(0, len(palette), dim)].astype(np.float64) / 255.0
            if rng.random() < 0.7:                       # a cell's own centroid: midway between its children 1.2c / 0.8c
                x = centroids[int(rng.integers(0, K // 2))]   # (a one-vector cell's centroid is its member)
            d_o = np.array([l2(x, c) for c in ours])
            dmin = d_o.min()
            band = d_o <= dmin + d_o.max() * 5.6843418860808015e-14
            if band.sum() < 2:
                continue
            cands = np.nonzero(d_o == dmin)[0]
            if len(cands) < 2 or len(cands) > 32 or not flags[np.nonzero(band)[0]].all() or (d_o[band] != dmin).any():
                continue                                  # the census calls these sensitive without looking at the order
            ties += 1
            win, fragile, path_fragile = census_descent(no, oo, inv_o, ours, x, cands)
            if fragile or not (is_robust or (PER_PATH and not path_fragile)):
                continue
            tie_safe += 1
            # the reference: its own codebook, its own tree; the tied candidates are the same numbers there
            d_r = np.array([l2(x, c) for c in ref])
            cr = np.nonzero(d_r == d_r.min())[0]
            ref_win = int(orr[first_visited(nr, orr, x, {int(inv_r[k]) for k in cr})])
            if ref_win != win:
                tie_bad += 1
                np.savez(f"/tmp/kd_census_fail_{shape_bad + tie_bad}.npz", ours=ours, ref=ref, flags=flags, x=x)
                print(f"TIE ORDER differs although the census calls it safe: K={K} dim={dim} ours {win} reference {ref_win} "
                      f"candidates {cands.tolist()} / {cr.tolist()}", flush=True)
    print(f"kd census fuzz: {levels} levels ({robust} reported robust, {shape_bad} of them with a different shape), "
          f"{ties} exact ties among reproducible candidates ({tie_safe} called safe, {tie_bad} of them decided differently) - seed {seed}")
    sys.exit(1 if shape_bad or tie_bad else 0)


if __name__ == "__main__":
    main()
