import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, quant_b200 as qb
rng = np.random.default_rng(0)
rgb = rng.integers(0, 256, (4096, 4096, 3), dtype=np.uint8)
ctx = qb.Context(0); ctx.set_image(rgb, 4096, 4096, 2, 2, 1); ctx.set_exact_centroids(True)
for nb in (0, 1, 4):
    t = time.time(); ctx.train(nb); print("nbits", nb, "seconds", round(time.time() - t, 4), flush=True)
