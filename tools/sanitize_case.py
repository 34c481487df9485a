"""Smallest case that touches every kernel (tensor-core filter included): for compute-sanitizer runs on a GPU box."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_b200 as qb

rng = np.random.default_rng(3)
xs, ys, w, h = 96, 64, 2, 2
rgb = rng.integers(0, 256, (ys, xs, 3), dtype=np.uint8)
rgb[:16, :32] = 200
with qb.Context(0) as ctx:
    ctx.set_image(rgb, xs, ys, w, h, qb.CS_SCALED)
    cb, d, rep = ctx.train(9)                                   # K = 2 .. 512: small-K fused, CUDA-core, tensor-core
    a = ctx.get_assign()
    _, mse = ctx.decode(qb.codebook_to_bytes(cb, qb.CS_SCALED))
    ctx.train(5, eps=1e-3, mode=qb.MODE_FULL_REPAIR)
    odd = rng.integers(0, 256, (67, 101, 3), dtype=np.uint8)    # non-divisible shape: slow gather path
    ctx.set_image(odd, 101, 67, 3, 2, qb.CS_NORMAL)
    ctx.train(4)
print("sanitize case ok", d, mse, int(a.max()))
