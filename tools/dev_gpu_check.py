"""Developer smoke check on a GPU box: CUDA path vs the C oracle on synthetic images."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_b200 as qb
from oracle.pyoracle import PortLib, SCALED, NORMAL

P = PortLib()
ctx = qb.Context(0)
print(ctx.device_info())
print("fp32 peak TFLOP/s", ctx.measure_fp32_peak())

def smooth_image(xs, ys, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:ys, 0:xs]
    base = 128 + 90*np.sin(xx/37.0 + seed) * np.cos(yy/23.0) + 20*np.sin((xx+yy)/5.0)
    img = np.stack([base, base*0.8+30, 255-base], -1) + rng.normal(0, 6, (ys, xs, 3))
    img[: ys//4, : xs//3] = 200          # flat region: exact duplicates
    return np.clip(img, 0, 255).astype(np.uint8)

def check(name, rgb, xs, ys, w, h, nbits, cs):
    X = P.blocks(rgb, xs, ys, w, h, cs)
    t = time.time(); cb, a, d, cb0, lv = P.quantize(X, nbits, levels=True); t_cpu = time.time() - t
    ctx.set_image(rgb, xs, ys, w, h, cs)
    T = P.blocks_lattice(rgb, xs, ys, w, h, cs).astype(np.int64)
    L = T - 128 if cs == SCALED else T
    bad = 0
    for l in lv:
        r = ctx.assign_accumulate(l["cb_pre"])
        mism = int((r["assign"].astype(np.uint64) != l["assign"]).sum())
        K = l["K"]
        n = np.bincount(l["assign"].astype(np.int64), minlength=K).astype(np.uint64)
        S = np.zeros((K, L.shape[1]), np.int64); np.add.at(S, l["assign"].astype(np.int64), L)
        Q = np.zeros(K, np.int64); np.add.at(Q, l["assign"].astype(np.int64), (L*L).sum(1))
        ok_stats = mism == 0 and np.array_equal(n, r["count"]) and np.array_equal(S, r["sum"]) and np.array_equal(Q.astype(np.uint64), r["sqsum"])
        post, d0, d1 = qb.finalize_level(cs, X.shape[0], r["count"], r["sum"], r["sqsum"], l["cb_pre"])
        rel = np.max(np.abs(post - l["cb_post"]) / np.maximum(np.abs(l["cb_post"]), 1e-300)) if mism == 0 else -1
        print(f"  {name} K={K:5d} mism={mism} flagged={r['flagged']} stats_ok={ok_stats} cb_rel={rel:.2e} d0 {d0:.6e}/{l['d0']:.6e} d1 {d1:.6e}/{l['d1']:.6e}")
        bad += mism + (0 if ok_stats else 1)
    t = time.time(); gcb, gd, rep = ctx.train(nbits); ga = ctx.get_assign_u64(); t_gpu = time.time() - t
    print(f"  {name} train: final assign mism {(ga != a).sum()} cb bytes diff {(qb.codebook_to_bytes(gcb, cs) != P.codebook_to_bytes(cb, cs)).sum()} "
          f"cb max rel {np.max(np.abs(gcb-cb)/np.maximum(np.abs(cb),1e-300)):.2e} dist {gd:.8e} vs {d:.8e}  cpu {t_cpu:.3f}s gpu {t_gpu:.4f}s")
    for r in rep:
        print("    ", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()})
    return bad

bad = 0
rng = np.random.default_rng(1234)
noise = rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)
bad += check("noise256 2x2", noise, 256, 256, 2, 2, 8, SCALED)
sm = smooth_image(384, 256, 3)
bad += check("smooth 2x2", sm, 384, 256, 2, 2, 10, SCALED)
bad += check("smooth 2x2 NORMAL", sm, 384, 256, 2, 2, 8, NORMAL)
bad += check("smooth 4x4", sm, 384, 256, 4, 4, 6, SCALED)
bad += check("smooth 1x1", sm[:128, :128].copy(), 128, 128, 1, 1, 8, SCALED)
odd = smooth_image(101, 67, 5)
bad += check("odd 2x2", odd, 101, 67, 2, 2, 6, SCALED)
bad += check("odd 3x2 (dim18 generic)", odd, 101, 67, 3, 2, 6, SCALED)
bad += check("odd 1x3", odd, 101, 67, 1, 3, 5, NORMAL)
flat = np.full((64, 64, 3), 77, np.uint8)
bad += check("flat", flat, 64, 64, 2, 2, 4, SCALED)
print("TOTAL BAD", bad)
sys.exit(1 if bad else 0)
