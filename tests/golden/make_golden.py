"""Generates tests/golden/*.npz from the REAL reference (oracle/_ref/libquantref_strict.so, built
from the unmodified sources under /root/reference by `make -C oracle ref`).

Run in the build container only (the GPU box has neither /root/reference nor a need for this):
    python tests/golden/make_golden.py
The fixtures pin both the C restatement (oracle/lbg_oracle.c) and the CUDA path.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import NORMAL, SCALED, RefLib  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
KODIM = "/root/reference/images/kodim"


def load_png(path):
    from PIL import Image
    im = np.asarray(Image.open(path).convert("RGB"))
    return np.ascontiguousarray(im), im.shape[1], im.shape[0]  # bytes, xSize = width, ySize = height


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def make_case(R, Rrel, name, rgb, xs, ys, w, h, nbits, cs, store_levels=True):
    rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1)
    X = R.blocks(rgb, xs, ys, w, h, cs)
    cb0, lv = R.levels(X, nbits)  # raises unless the replay == quantize() bit-for-bit
    cb, a, dist = R.quantize(X, nbits)
    comp = R.compress(rgb, xs, ys, w, h, nbits, cs)
    assert np.array_equal(comp["assign"], a)
    cbb = R.codebook_to_bytes(cb, cs)
    assert np.array_equal(cbb, comp["codebook_bytes"])
    dec = R.decode(cbb, a, xs, ys, w, h)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "x.quant")
        R.compress_to_file(rgb, xs, ys, w, h, nbits, p, cs)
        qbytes = open(p, "rb").read()
        dec2, dxs, dys = R.decompress_file(p, xs * ys * 3)
        assert (dxs, dys) == (xs, ys) and np.array_equal(dec2, dec)
    rel = Rrel.compress(rgb, xs, ys, w, h, nbits, cs)
    d = dict(rgb=rgb, params=np.array([xs, ys, w, h, nbits, cs], np.int64), cb0=cb0, codebook=cb,
             assign=a.astype(np.uint32), distortion=np.float64(dist),
             report_distortion=np.float64(comp["distortion"]), bpp=np.float32(comp["bpp"]),
             codebook_bytes=cbb, decoded_sha=np.array(sha(dec)), quant_sha=np.array(sha(qbytes)),
             quant_len=np.int64(len(qbytes)),
             release_assign_mismatch=np.int64((rel["assign"] != a).sum()),
             release_codebook_bytes_mismatch=np.int64((rel["codebook_bytes"] != cbb).sum()))
    if store_levels:
        for i, l in enumerate(lv):
            adt = np.uint8 if l["K"] <= 256 else np.uint16
            d[f"L{i}_pre"] = l["cb_pre"]
            d[f"L{i}_post"] = l["cb_post"]
            d[f"L{i}_assign"] = l["assign"].astype(adt)
            d[f"L{i}_d"] = np.array([l["d0"], l["d1"]])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    used = len(np.unique(a))
    print(f"{name}: N={X.shape[0]} dim={X.shape[1]} K={1 << nbits} used={used} dist={dist:.6g} "
          f"release-build diffs: {d['release_assign_mismatch']} idx / "
          f"{d['release_codebook_bytes_mismatch']} bytes")
    return cb


def main():
    R, Rrel = RefLib("strict"), RefLib("release")
    k1, xs1, ys1 = load_png(os.path.join(KODIM, "kodim01.png"))
    k5, _, _ = load_png(os.path.join(KODIM, "kodim05.png"))
    k23, _, _ = load_png(os.path.join(KODIM, "kodim23.png"))

    # BASELINE config 1: README default on a Kodak image (final state only; image stored once)
    cb_k1 = make_case(R, Rrel, "kodim01_full_2x2_n10", k1, xs1, ys1, 2, 2, 10, SCALED,
                      store_levels=False)
    crop = np.ascontiguousarray(k1[128:384, 192:576])  # 384 x 256
    make_case(R, Rrel, "kodim01_crop_2x2_n10", crop, 384, 256, 2, 2, 10, SCALED)
    make_case(R, Rrel, "kodim01_crop_2x2_n8_normal", crop, 384, 256, 2, 2, 8, NORMAL)
    crop4 = np.ascontiguousarray(k23[100:356, 200:584])
    make_case(R, Rrel, "kodim23_crop_4x4_n8", crop4, 384, 256, 4, 4, 8, SCALED)
    small = np.ascontiguousarray(k5[40:168, 300:428])  # 128 x 128
    make_case(R, Rrel, "kodim05_small_1x1_n8", small, 128, 128, 1, 1, 8, SCALED)
    odd = np.ascontiguousarray(k5[200:267, 100:201])  # 101 wide, 67 high: nothing divides
    make_case(R, Rrel, "odd_101x67_2x2_n6", odd, 101, 67, 2, 2, 6, SCALED)
    make_case(R, Rrel, "odd_101x67_3x2_n5", odd, 101, 67, 3, 2, 5, SCALED)
    make_case(R, Rrel, "odd_101x67_1x3_n5_normal", odd, 101, 67, 1, 3, 5, NORMAL)
    make_case(R, Rrel, "odd_101x67_2x4_n4", odd, 101, 67, 2, 4, 4, SCALED)
    make_case(R, Rrel, "odd_67x101_3x3_n5", odd, 67, 101, 3, 3, 5, SCALED)
    rng = np.random.default_rng(1234)
    make_case(R, Rrel, "noise_96x64_2x2_n7", rng.integers(0, 256, (64, 96, 3), dtype=np.uint8), 96, 64,
              2, 2, 7, SCALED)
    make_case(R, Rrel, "flat_32x32_2x2_n4", np.full((32, 32, 3), 77, np.uint8), 32, 32, 2, 2, 4, SCALED)
    two = np.full((32, 32, 3), 10, np.uint8)
    two[:, 16:] = 240
    make_case(R, Rrel, "twotone_32x32_2x2_n5", two, 32, 32, 2, 2, 5, SCALED)
    make_case(R, Rrel, "tiny_8x8_2x2_n6", np.ascontiguousarray(k5[10:18, 10:18]), 8, 8, 2, 2, 6, SCALED)
    # the reference's own unit-test image (src/test.cpp:6-13): 4x4 letters, NORMAL
    letters = np.array([list(b"abc"), list(b"def"), list(b"ghi"), list(b"jkl")] * 4, np.uint8)
    lay = {}
    for (w, h) in [(1, 1), (2, 2), (1, 3), (2, 4)]:
        blk = R.blocks(letters, 4, 4, w, h, NORMAL)
        lay[f"blocks_{w}x{h}"] = blk
        by = R.codebook_to_bytes(blk, NORMAL)
        n = blk.shape[0]
        lay[f"roundtrip_{w}x{h}"] = R.decode(by, np.arange(n, dtype=np.uint64), 4, 4, w, h)
    np.savez_compressed(os.path.join(OUT, "letters_layout.npz"), rgb=letters.reshape(-1), **lay)
    # encode-only (BASELINE config 5 in miniature): fixed FP64 codebook trained on kodim01,
    # applied to crops of other Kodak images through KDTree::nearestNeighbour
    enc_rgb = np.stack([np.ascontiguousarray(k5[0:128, 0:192]), np.ascontiguousarray(k23[64:192, 300:492])])
    idx = np.stack([R.nn_rgb(cb_k1, enc_rgb[i], 192, 128, 2, 2, SCALED) for i in range(2)])
    np.savez_compressed(os.path.join(OUT, "encode_only_k1024.npz"), rgb=enc_rgb.reshape(-1),
                        params=np.array([192, 128, 2, 2, 2], np.int64), codebook=cb_k1,
                        assign=idx.astype(np.uint16))
    print("encode-only fixture:", idx.shape)


def main_fp64():
    """Fixtures of the general FP64-vector path: CIE1931 images and arbitrary double vectors.
    Written by `python tests/golden/make_golden.py fp64` without touching the other fixtures."""
    R, Rrel = RefLib("strict"), RefLib("release")
    k1, _, _ = load_png(os.path.join(KODIM, "kodim01.png"))
    k5, _, _ = load_png(os.path.join(KODIM, "kodim05.png"))
    crop = np.ascontiguousarray(k1[128:256, 192:384])  # 192 x 128
    make_case(R, Rrel, "kodim01_small_2x2_n8_cie", crop, 192, 128, 2, 2, 8, 2)
    odd = np.ascontiguousarray(k5[200:267, 100:201])
    make_case(R, Rrel, "odd_101x67_3x2_n5_cie", odd, 101, 67, 3, 2, 5, 2)
    two = np.full((32, 32, 3), 10, np.uint8)
    two[:, 16:] = 240
    make_case(R, Rrel, "twotone_32x32_2x2_n5_cie", two, 32, 32, 2, 2, 5, 2)
    rng = np.random.default_rng(77)
    sets = {"generic_gauss_d7_n6": (rng.normal(size=(6000, 7)) * np.array([1, 10, 0.1, 5, 2, 3, 1e-3]), 6),
            "generic_clusters_d12_n7": (rng.normal(size=(40, 12))[rng.integers(0, 40, 9000)] * 50
                                        + rng.normal(size=(9000, 12)) * 0.01, 7),
            "generic_dups_d5_n6": (np.round(rng.normal(size=(4000, 5)) * 2) / 3.0, 6),   # heavy duplicates, exact ties
            "generic_big_d40_n5": (rng.normal(size=(3000, 40)) * 1e6 + 3e6, 5)}
    for name, (X, nbits) in sets.items():
        X = np.ascontiguousarray(X, np.float64)
        cb0, lv = R.levels(X, nbits)
        cb, a, dist = R.quantize(X, nbits)
        d = dict(X=X, nbits=np.int64(nbits), cb0=cb0, codebook=cb, assign=a.astype(np.uint32), distortion=np.float64(dist))
        for i, l in enumerate(lv):
            d[f"L{i}_pre"] = l["cb_pre"]
            d[f"L{i}_post"] = l["cb_post"]
            d[f"L{i}_assign"] = l["assign"].astype(np.uint16)
            d[f"L{i}_d"] = np.array([l["d0"], l["d1"]])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(f"{name}: N={X.shape[0]} dim={X.shape[1]} K={1 << nbits} used={len(np.unique(a))} dist={dist:.6g}")


if __name__ == "__main__":
    main_fp64() if sys.argv[1:] == ["fp64"] else main()
