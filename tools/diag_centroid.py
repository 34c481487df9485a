"""Diagnostic (GPU box): where do full trains on degenerate images leave the oracle?  Per level, with the ORACLE's
pre-fix codebook as input, compare indices (must be exact) and centroids (bits)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_b200 as qb
from oracle.pyoracle import PortLib
P = PortLib(); ctx = qb.Context(0)
rng = np.random.default_rng(5)
found = 0
for trial in range(400):
    xs, ys, w, h, nbits = int(rng.integers(20, 160)), int(rng.integers(20, 120)), 2, 2, 7
    kind = trial % 4
    if kind == 0: img = rng.integers(0, 256, (ys, xs, 3))
    elif kind == 1:
        yy, xx = np.mgrid[0:ys, 0:xs]; base = (xx * 3 + yy * 2) % 256
        img = np.stack([base, (base + 40) % 256, 255 - base], -1) + rng.integers(-2, 3, (ys, xs, 3))
    elif kind == 2:
        pal = rng.integers(0, 256, (5, 3)); img = pal[rng.integers(0, 5, (ys, xs))]
    else:
        img = np.full((ys, xs, 3), int(rng.integers(0, 256)))
        for _ in range(10): img[rng.integers(0, ys), rng.integers(0, xs)] = rng.integers(0, 256, 3)
    rgb = np.clip(img, 0, 255).astype(np.uint8)
    X = P.blocks(rgb, xs, ys, w, h, 1)
    cb_o, a_o, d_o, cb0, lv = P.quantize(X, nbits, levels=True)
    ctx.set_image(rgb, xs, ys, w, h, 1)
    cb, d, _ = ctx.train(nbits)
    a = ctx.get_assign().astype(np.uint64)
    if np.array_equal(a, a_o): continue
    found += 1
    msg = [f"trial {trial} kind {kind} N={X.shape[0]}: final mismatch {int((a != a_o).sum())}"]
    for l in lv:
        r = ctx.assign_accumulate(l["cb_pre"])
        mism = int((r["assign"].astype(np.uint64) != l["assign"]).sum())
        post, d0, d1 = qb.finalize_level(1, X.shape[0], r["count"], r["sum"], r["sqsum"], l["cb_pre"])
        nb = int((post != l["cb_post"]).sum())
        rel = float(np.max(np.abs(post - l["cb_post"]) / np.maximum(np.abs(l["cb_post"]), 1e-300)))
        msg.append(f"  K={l['K']:4d} idx mism (oracle cb in) {mism}  centroid elements not bit-equal {nb}/{post.size}  max rel {rel:.1e}")
    print("\n".join(msg), flush=True)
    if found >= 4: break
print("diverging trials:", found)
