"""CPU fuzz of the auto mode's census on WHOLE TRAINS (no GPU):  python tools/fuzz_census_trains.py [seconds] [seed]

Random small images of the kinds tools/fuzz_parity.py uses (noise, gradients, palettes, flat with outliers) are trained
by the C oracle (= the reference's arithmetic), which gives every split level's codebook and assignment.  Each level's
codebook is then rebuilt in the library's arithmetic (tools/fuzz_kd_census.py: our_level) and put through the three
checks of check_level: flagged-reproducible codevectors equal the reference's bits, a robust census margin implies the
same KD tree, order-safe ties are decided as the reference's walk decides them.  This is the host half of what
fuzz_parity.py checks end to end on a GPU, at a few hundred trains per minute per core."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fuzz_kd_census as F
from oracle.pyoracle import PortLib
from quant_b200 import _lib


def make_image(rng, xs, ys, kind):
    if kind == 0:
        img = rng.integers(0, 256, (ys, xs, 3))
    elif kind == 1:
        yy, xx = np.mgrid[0:ys, 0:xs]
        base = (xx * 3 + yy * 2) % 256
        img = np.stack([base, (base + 40) % 256, 255 - base], -1) + rng.integers(-2, 3, (ys, xs, 3))
    elif kind == 2:
        pal = rng.integers(0, 256, (5, 3))
        img = pal[rng.integers(0, 5, (ys, xs))]
    else:
        img = np.full((ys, xs, 3), int(rng.integers(0, 256)))
        for _ in range(10):
            img[rng.integers(0, ys), rng.integers(0, xs)] = rng.integers(0, 256, 3)
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    lib = _lib.load()
    P = PortLib()
    t_end = time.time() + budget
    trains = levels = robust = ties = bad = repeat_per_level = repeat_per_path = 0
    while time.time() < t_end:
        w, h = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        if rng.random() < 0.05:
            w, h = int(rng.integers(5, 7)), int(rng.integers(5, 7))
        xs, ys = int(rng.integers(w, 120)), int(rng.integers(h, 100))
        kind, nbits = int(rng.integers(0, 4)), int(rng.integers(2, 10))
        rgb = make_image(rng, xs, ys, kind)
        X = P.blocks(rgb, xs, ys, w, h, 1)
        L = P.blocks_lattice(rgb, xs, ys, w, h, 1).astype(np.int64)
        _, _, _, _, lv = P.quantize(X, nbits, levels=True)
        Xu = np.unique(X, axis=0)
        if len(Xu) > 400:
            Xu = Xu[rng.choice(len(Xu), 400, replace=False)]
        trains += 1
        stats = {}
        for l in range(1, len(lv)):
            ours, flags = F.our_level(L, lv[l - 1]["assign"].astype(np.int64), lv[l - 1]["K"])
            levels += 1
            try:
                r, t = F.check_level(lib, ours, flags, lv[l]["cb_pre"], Xu, stats)
                robust += r
                ties += t
            except AssertionError as e:
                bad += 1
                np.savez(f"/tmp/census_train_fail_{bad}.npz", rgb=rgb, xs=xs, ys=ys, w=w, h=h, nbits=nbits, level=l)
                print(f"VIOLATION kind={kind} xs={xs} ys={ys} w={w} h={h} nbits={nbits} level {l}: {e}", flush=True)
        repeat_per_level += stats.get("unsafe_per_level", 0) > 0
        repeat_per_path += stats.get("unsafe_per_path", 0) > 0
    print(f"trains with an order-unsafe exact tie among reproducible candidates (-> exact repeat): {repeat_per_level} under one margin "
          f"per level, {repeat_per_path} under the per-path rule")
    print(f"census train fuzz: {trains} trains, {levels} levels ({robust} robust), {ties} order-safe ties checked, {bad} violations - seed {seed}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
