"""GPU box: where do the CUDA-core levels (K = 32, 64) spend their time?  Filter alone vs filter with fused statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, quant_b200 as qb
rng = np.random.default_rng(0)
rgb = rng.integers(0, 256, (4096, 4096, 3), dtype=np.uint8)
ctx = qb.Context(0); ctx.set_image(rgb, 4096, 4096, 2, 2, 1)
cb, d, rep = ctx.train(10)
for K in (16, 32, 64, 128, 256, 512, 1024):
    cbk = np.ascontiguousarray(cb[:K])
    for _ in range(3): r = ctx.assign_only(cbk)
    print(f"K={K}: filter alone {r['ms_assign']:.3f} ms (+ resolve {r['ms_resolve']:.3f}); in the train: "
          f"{[ (x['ms_assign'], x['ms_accumulate']) for x in rep if x['K'] == K]}")
