"""GPU box: time the exact-centroid mode and the FP64-vector path (CIE1931) at config-2 size, and check a mid-size
CIE train against the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, quant_b200 as qb
from oracle.pyoracle import PortLib
P = PortLib()
rng = np.random.default_rng(0)
ctx = qb.Context(0)
rgb = rng.integers(0, 256, (1024, 1024, 3), dtype=np.uint8)
X = P.blocks(rgb, 1024, 1024, 2, 2, 2)
t = time.time(); cb_o, a_o, d_o = P.quantize(X, 8); t_cpu = time.time() - t
ctx.set_image(rgb, 1024, 1024, 2, 2, 2)
t = time.time(); cb, d, rep = ctx.train(8); t_gpu = time.time() - t
print(f"CIE 1024^2 2x2 K=256: indices equal {np.array_equal(ctx.get_assign().astype(np.uint64), a_o)}, codebook bits equal "
      f"{cb.tobytes() == cb_o.tobytes()}, oracle {t_cpu:.2f} s (1 thread), GPU {t_gpu:.3f} s", flush=True)
rgb = rng.integers(0, 256, (4096, 4096, 3), dtype=np.uint8)
for cs, name in ((1, "SCALED exact-centroid mode"), (2, "CIE1931 (FP64 vectors)")):
    ctx.set_exact_centroids(cs == 1)
    ctx.set_image(rgb, 4096, 4096, 2, 2, cs)
    ctx.train(10)
    t = time.time(); cb, d, rep = ctx.train(10); dt = time.time() - t
    print(f"{name}: 4096^2 2x2 K=1024 train {dt * 1e3:.1f} ms; last level: assign {rep[-1]['ms_assign']:.2f} ms, "
          f"resolve {rep[-1]['ms_resolve']:.2f} ms, flagged {rep[-1]['flagged']}", flush=True)
