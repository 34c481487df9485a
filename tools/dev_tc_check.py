"""Developer check on a GPU box: tensor-core filter vs CUDA-core filter vs the C oracle (exact)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_b200 as qb
from oracle.pyoracle import PortLib, SCALED

P = PortLib()
ctx = qb.Context(0)
bad = 0
rng = np.random.default_rng(7)
for (xs, ys, w, h) in [(512, 256, 2, 2), (512, 256, 4, 4), (256, 256, 1, 1), (300, 200, 3, 3), (256, 128, 2, 4)]:
    rgb = rng.integers(0, 256, (ys, xs, 3), dtype=np.uint8)
    dim = 3 * w * h
    ctx.set_image(rgb, xs, ys, w, h, SCALED)
    X = P.blocks(rgb, xs, ys, w, h, SCALED)
    for K in (16, 48, 64, 256, 1000, 1024, 4096):
        cb = rng.random((K, dim))
        if K >= 64:
            cb[K - 8:] = 0.0          # dead cells: exact duplicates
            cb[5] = cb[4]             # an exact duplicate pair
        ctx.set_tensor_cores(True)
        t = time.time(); r1 = ctx.assign_accumulate(cb, want_stats=False); t1 = time.time() - t
        ctx.set_tensor_cores(False)
        t = time.time(); r0 = ctx.assign_accumulate(cb, want_stats=False); t0 = time.time() - t
        n = min(X.shape[0], 6000)
        want = P.assign(X[:n], cb)
        m_tc = int((r1["assign"][:n].astype(np.uint64) != want).sum())
        m_cc = int((r0["assign"][:n].astype(np.uint64) != want).sum())
        m_x = int((r1["assign"] != r0["assign"]).sum())
        print(f"dim {dim:3d} K {K:5d} N {X.shape[0]:7d}: tc-vs-oracle {m_tc} cc-vs-oracle {m_cc} tc-vs-cc {m_x} "
              f"flagged tc {r1['flagged']} cc {r0['flagged']}  ({t1*1e3:.1f} / {t0*1e3:.1f} ms)", flush=True)
        bad += m_tc + m_cc + m_x
print("TOTAL BAD", bad)
sys.exit(1 if bad else 0)
