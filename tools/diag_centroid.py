"""Diagnostic (GPU box): where and why a default-mode SCALED train leaves the oracle on a duplicate-heavy image.
Replays the library's own schedule level by level (integer-sum centroids) next to the oracle's and prints the first
diverging query with both candidates."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_b200 as qb
from oracle.pyoracle import PortLib
P = PortLib(); ctx = qb.Context(0)
rng = np.random.default_rng(5)
found = 0
for trial in range(3000):
    xs, ys, w, h, nbits = int(rng.integers(20, 160)), int(rng.integers(20, 120)), int(rng.integers(1, 3)), int(rng.integers(1, 3)), int(rng.integers(4, 9))
    pal = rng.integers(0, 256, (5, 3)); img = pal[rng.integers(0, 5, (ys, xs))]
    rgb = np.clip(img, 0, 255).astype(np.uint8)
    X = P.blocks(rgb, xs, ys, w, h, 1)
    cb_o, a_o, d_o, cb0, lv = P.quantize(X, nbits, levels=True)
    ctx.set_image(rgb, xs, ys, w, h, 1)
    cb, d, _ = ctx.train(nbits)
    a = ctx.get_assign().astype(np.uint64)
    if np.array_equal(a, a_o): continue
    found += 1
    print(f"trial {trial} N={X.shape[0]} dim={X.shape[1]} nbits={nbits}: final mismatch {int((a != a_o).sum())}")
    mine = np.ascontiguousarray(cb0, np.float64).reshape(1, -1).copy()
    for li, l in enumerate(lv):
        mine = P.split(mine)
        r = ctx.assign_accumulate(mine)
        am = r["assign"].astype(np.uint64)
        bits = int((mine != l["cb_pre"]).sum())
        mism = np.where(am != l["assign"])[0]
        print(f"   K={l['K']:4d}: my codebook differs from the oracle's in {bits} elements; {len(mism)} indices differ")
        if len(mism):
            i = int(mism[0]); x = X[i]; k1, k2 = int(am[i]), int(l["assign"][i])
            for name, C in (("mine", mine), ("oracle", l["cb_pre"])):
                d1 = float(((x - C[k1]) ** 2).sum()); d2 = float(((x - C[k2]) ** 2).sum())
                print(f"      {name}: query {i}: d(k={k1})={d1!r} d(k={k2})={d2!r}  diff {d1 - d2:.3e}")
            par = l["K"] // 2
            print(f"      k1 % parent = {k1 % par}, k2 % parent = {k2 % par}; members of the parent cell at the previous level:",
                  "n/a" if li == 0 else (int((lv[li - 1]["assign"] == k1 % par).sum()), len(np.unique(X[lv[li - 1]["assign"] == k1 % par], axis=0))))
            print("      x == parent centroid (oracle)?", "n/a" if li == 0 else bool((x == lv[li - 1]["cb_post"][k1 % par]).all()),
                  " max|x - c|:", "n/a" if li == 0 else float(np.abs(x - lv[li - 1]["cb_post"][k1 % par]).max()))
            break
        mine, _, _ = qb.finalize_level(1, X.shape[0], r["count"], r["sum"], r["sqsum"], mine)
    if found >= 4: break
print("diverging trials:", found)
