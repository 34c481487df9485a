"""Host-side mirror of the reference interface (quant_b200/*.py): layout, byte conversion, decode
and the .quant container - written the way the reference's own test is (src/test.cpp)."""
import hashlib
import os

import numpy as np
import pytest

import quant_b200 as qb
from quant_b200 import (ColorSpaces, CompressedImage, Quantizers, RGBImage, getBlocksAsVectorsFromImage,
                        getImageFromVectors, getQuantizer, vectorsToCharVectorsColorSpaced)
from conftest import GOLDEN, golden_names, load_golden


def letters_image():
    px = np.array([list(b"abc"), list(b"def"), list(b"ghi"), list(b"jkl")] * 4, np.uint8)
    return RGBImage.from_array(px, 4, 4)


@pytest.mark.parametrize("w,h", [(1, 1), (2, 2), (1, 3), (2, 4)])
def test_compressor_test_something(w, h):
    """compressor_test.something (src/test.cpp:5-62): blocks -> chars -> image is the identity."""
    img = letters_image()
    blocks = getBlocksAsVectorsFromImage(img, w, h, ColorSpaces.NORMAL)
    converted = vectorsToCharVectorsColorSpaced(blocks, ColorSpaces.NORMAL)
    expected = getImageFromVectors(converted, img.xSize, img.ySize, w, h)
    assert np.array_equal(expected.img, img.img)
    z = np.load(os.path.join(GOLDEN, "letters_layout.npz"))
    assert np.array_equal(blocks, z[f"blocks_{w}x{h}"])  # the reference's own vectors


@pytest.mark.parametrize("name", golden_names())
def test_blocks_decode_and_container_match_reference(port, name):
    g = load_golden(name)
    img = RGBImage.from_array(g.rgb, g.xs, g.ys)
    blocks = getBlocksAsVectorsFromImage(img, g.w, g.h, g.cs)
    assert np.array_equal(blocks, port.blocks(g.rgb, g.xs, g.ys, g.w, g.h, g.cs))
    ci = CompressedImage()
    ci.codeVectors = vectorsToCharVectorsColorSpaced(g.z["codebook"], g.cs)
    assert np.array_equal(ci.codeVectors, g.z["codebook_bytes"])
    ci.assignedCodeVector = g.z["assign"].astype(np.uint64)
    ci.xSize, ci.ySize, ci.blockWidth, ci.blockHeight = g.xs, g.ys, g.w, g.h
    ci.colorSpace = ColorSpaces(g.cs)
    dec = CompressedImage.decompress(ci)
    assert hashlib.sha256(dec.img.tobytes()).hexdigest() == str(g.z["decoded_sha"])
    blob = ci.to_bytes()
    assert len(blob) == int(g.z["quant_len"])
    assert hashlib.sha256(blob).hexdigest() == str(g.z["quant_sha"])
    assert np.float32(ci.sizeInBits()) / np.float32(g.xs * g.ys) == np.float32(g.z["bpp"])


def test_quant_file_round_trip(tmp_path):
    g = load_golden("odd_101x67_2x2_n6")
    ci = CompressedImage()
    ci.codeVectors = g.z["codebook_bytes"]
    ci.assignedCodeVector = g.z["assign"].astype(np.uint64)
    ci.xSize, ci.ySize, ci.blockWidth, ci.blockHeight = g.xs, g.ys, g.w, g.h
    p = str(tmp_path / "a.quant")
    ci.saveToFile(p)
    back = CompressedImage()
    back.loadFromFile(p)
    assert np.array_equal(back.codeVectors, ci.codeVectors)
    assert np.array_equal(back.assignedCodeVector, ci.assignedCodeVector)
    assert (back.xSize, back.ySize, back.blockWidth, back.blockHeight) == (g.xs, g.ys, g.w, g.h)
    assert CompressedImage.decompress(back) == CompressedImage.decompress(ci)


def test_ppm_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    img = RGBImage.from_array(rng.integers(0, 256, (7 * 5, 3), dtype=np.uint8), 7, 5)
    p = str(tmp_path / "a.ppm")
    img.saveToFile(p)
    assert open(p, "rb").read().startswith(b"P6\n7 5\n255\n")
    assert RGBImage(p) == img
    assert img.sizeInBytes() == 105


def test_get_quantizer_factory():
    assert getQuantizer(Quantizers.MEDIAN_CUT) is None      # src/Quantizer.cpp:146-155
    assert getQuantizer(Quantizers.LBG_MEDIAN_CUT) is None
    assert getQuantizer(Quantizers.ABC) is None
    assert isinstance(getQuantizer(Quantizers.LBG), qb.LBGQuantizer)  # no device touched yet


def test_vectors_to_lattice_bytes(port):
    g = load_golden("odd_101x67_2x2_n6")
    X = port.blocks(g.rgb, g.xs, g.ys, g.w, g.h, 1)
    mat, cs = qb.vectors_to_lattice_bytes(X)
    assert cs == qb.CS_SCALED
    assert np.array_equal(((mat.astype(np.int16) ^ 0x80)) / 255.0, X)
    Xn = port.blocks(g.rgb, g.xs, g.ys, g.w, g.h, 0)
    mat, cs = qb.vectors_to_lattice_bytes(Xn)
    assert cs == qb.CS_NORMAL and np.array_equal(mat.astype(np.int8).astype(np.float64), Xn)
    with pytest.raises(ValueError):
        qb.vectors_to_lattice_bytes(X + 1e-9)
    with pytest.raises(IndexError):
        qb.LBGQuantizer().quantize(np.zeros((0, 12)), 4, 1e-6)  # trainingSet.at(0) throws


def test_cie1931_block_vectors_match_the_oracle(port):
    """Cie1931::RGBtoColorSpace through the host mirror == the C port (itself pinned to the real reference)."""
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (7, 9, 3), dtype=np.uint8)
    img = qb.RGBImage.from_array(rgb, 9, 7)
    for (w, h) in [(1, 1), (2, 2), (4, 3)]:
        got = getBlocksAsVectorsFromImage(img, w, h, ColorSpaces.CIE1931)
        assert got.tobytes() == port.blocks(rgb, 9, 7, w, h, 2).tobytes()
    with pytest.raises(ValueError):
        getBlocksAsVectorsFromImage(img, 1, 1, 3)


def test_finalize_level_reproduces_the_compensated_sum_of_equal_members(port):
    """A cell whose members are all the same vector: the reference's Kahan sum of n equal terms v is fl(n*v), and
    qb200_finalize_level must return fl(fl(n*v)/n) - bit for bit what Solution::fixCodeVectors computes."""
    rng = np.random.default_rng(9)
    dim = 6
    for n in [1, 2, 3, 7, 10, 66, 255, 1000, 4099, 65537]:
        t = rng.integers(0, 256, (4, dim))                      # four cells, each n copies of one lattice vector
        X = np.repeat(t / 255.0, n, axis=0)
        assign = np.repeat(np.arange(4, dtype=np.uint64), n)
        want = port.fix(X, assign, 4)
        L = (t - 128).astype(np.int64)
        count = np.full(4, n, np.uint64)
        post, _, _ = qb.finalize_level(qb.CS_SCALED, 4 * n, count, L * n, ((L * L).sum(1) * n).astype(np.uint64))
        assert post.tobytes() == want.tobytes(), n
    # two different members: the plain integer-sum formula, within a few ulp
    X = np.array([[10, 20, 30, 40, 50, 60], [11, 20, 30, 40, 50, 61]]) / 255.0
    L = np.array([[10, 20, 30, 40, 50, 60], [11, 20, 30, 40, 50, 61]]) - 128
    post, _, _ = qb.finalize_level(qb.CS_SCALED, 2, np.array([2], np.uint64), L.sum(0, keepdims=True),
                                   np.array([(L * L).sum()], np.uint64))
    assert np.allclose(post, port.fix(X, np.zeros(2, np.uint64), 1), rtol=1e-15, atol=0)
