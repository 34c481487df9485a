"""The C++ drop-in host layer (quant_b200/host: Quantizer.hpp / Compressor.hpp / `quant` CLI over the C ABI).

CPU part: the reference's own unit test restated (layout round trips), .quant decode through the CLI against
the reference's decoded image, and a loud failure (no CPU fallback) when compressing without a GPU.
GPU part: the CLI's .quant file and decoded PPM must be byte-identical to the reference's on every fixture."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, HAVE_GPU, golden_names, load_golden

HOST = os.path.join(ROOT, "quant_b200", "host")
QUANT = os.path.join(HOST, "quant")
HOST_TEST = os.path.join(HOST, "host_test")


@pytest.fixture(scope="module")
def host_built():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "quant_b200", "csrc"), "-j4"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HOST], stdout=subprocess.DEVNULL)
    return True


def write_ppm(path, rgb, xs, ys):
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (xs, ys))
        f.write(np.ascontiguousarray(rgb, np.uint8).tobytes())


def ppm_payload(path, xs, ys):
    data = open(path, "rb").read()
    return data[len(data) - xs * ys * 3:]


def test_layout_roundtrips_like_the_reference_unit_test(host_built):
    out = subprocess.run([HOST_TEST, "layout"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.parametrize("name", ["odd_101x67_2x2_n6", "odd_101x67_1x3_n5_normal", "kodim23_crop_4x4_n8"])
def test_cli_decompress_matches_reference_decode(host_built, tmp_path, name):
    """loadFromFile + decompress + saveToFile are host code: a .quant built from the reference's codebook
    bytes and indices must decode to the reference's image."""
    import quant_b200 as qb
    g = load_golden(name)
    ci = qb.CompressedImage()
    ci.codeVectors, ci.assignedCodeVector = g.z["codebook_bytes"], g.z["assign"].astype(np.uint64)
    ci.xSize, ci.ySize, ci.blockWidth, ci.blockHeight = g.xs, g.ys, g.w, g.h
    ci.colorSpace = qb.ColorSpaces(g.cs)
    blob = ci.to_bytes()
    assert hashlib.sha256(blob).hexdigest() == str(g.z["quant_sha"])
    q, p = str(tmp_path / "a.quant"), str(tmp_path / "a.ppm")
    open(q, "wb").write(blob)
    r = subprocess.run([QUANT, q, "-o", p], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert hashlib.sha256(ppm_payload(p, g.xs, g.ys)).hexdigest() == str(g.z["decoded_sha"])


@pytest.mark.skipif(HAVE_GPU, reason="checks the behaviour on a machine WITHOUT a GPU")
def test_compress_without_gpu_fails_loudly(host_built, tmp_path):
    g = load_golden("tiny_8x8_2x2_n6")
    p = str(tmp_path / "t.ppm")
    write_ppm(p, g.rgb, g.xs, g.ys)
    r = subprocess.run([QUANT, p, "-o", str(tmp_path / "t.quant"), "-n", "4"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU" in r.stderr and not os.path.exists(str(tmp_path / "t.quant"))


def test_cli_rejects_bad_arguments(host_built, tmp_path):
    r = subprocess.run([QUANT, "x.ppm"], capture_output=True, text=True)
    assert r.returncode == 2 and "saveto" in r.stderr
    r = subprocess.run([QUANT, "x.txt", "-o", "y.bin"], capture_output=True, text=True)
    assert r.returncode == 1 and "File type not supported" in r.stderr
    assert subprocess.run([QUANT, "--help"], capture_output=True).returncode == 0


CIE_GOLDENS = ["kodim01_small_2x2_n8_cie", "odd_101x67_3x2_n5_cie"]   # --c 2: FP64 vectors on the device


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_names() + CIE_GOLDENS)
def test_cli_compress_writes_the_reference_quant_file(host_built, tmp_path, name):
    g = load_golden(name)
    p, q, d = str(tmp_path / "i.ppm"), str(tmp_path / "o.quant"), str(tmp_path / "o.ppm")
    write_ppm(p, g.rgb, g.xs, g.ys)
    args = ["-n", str(g.nbits), "-w", str(g.w), "-h", str(g.h), "--c", str(g.cs)]
    r = subprocess.run([QUANT, p, "-o", q, "-r", "1"] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert hashlib.sha256(open(q, "rb").read()).hexdigest() == str(g.z["quant_sha"])
    rep = dict(l.split("=", 1) for l in r.stdout.splitlines() if "=" in l)
    dist = float([v for k, v in rep.items() if k.startswith("Distortion")][0])
    assert dist == pytest.approx(float(g.z["report_distortion"]), abs=1e-9)
    r = subprocess.run([QUANT, p, "-o", d] + args, capture_output=True, text=True)   # "showcase": ppm -> ppm
    assert r.returncode == 0, r.stderr
    assert hashlib.sha256(ppm_payload(d, g.xs, g.ys)).hexdigest() == str(g.z["decoded_sha"])


@pytest.mark.gpu
def test_plugin_quantize_entry_agrees_with_compress(host_built, tmp_path):
    g = load_golden("odd_101x67_2x2_n6")
    p = str(tmp_path / "i.ppm")
    write_ppm(p, g.rgb, g.xs, g.ys)
    r = subprocess.run([HOST_TEST, "quantize", p, str(g.w), str(g.h), str(g.nbits)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # CIE1931: the generic entry sees vectors off the byte lattice and takes the FP64 path; same result as compress()
    r = subprocess.run([HOST_TEST, "quantize", p, str(g.w), str(g.h), str(g.nbits), "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cli_exact_centroid_mode_on_a_palette_image(host_built, port, tmp_path):
    """QB200_EXACT_CENTROIDS=1 reaches the C++ layer through the environment: on a five-colour image (exact ties
    at every split) the CLI's .quant file must be the oracle's byte for byte."""
    xs, ys, w, h, nbits = 141, 138, 2, 2, 8
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (5, 3))[rng.integers(0, 5, (ys, xs))].astype(np.uint8)
    X = port.blocks(rgb, xs, ys, w, h, 1)
    cb_o, a_o, _ = port.quantize(X, nbits)
    want = port.quant_serialize(port.codebook_to_bytes(cb_o, 1), a_o, xs, ys, w, h, 1)
    p, q = str(tmp_path / "i.ppm"), str(tmp_path / "o.quant")
    write_ppm(p, rgb, xs, ys)
    env = dict(os.environ, QB200_EXACT_CENTROIDS="1")
    r = subprocess.run([QUANT, p, "-o", q, "-n", str(nbits), "-w", str(w), "-h", str(h), "--c", "1"],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert open(q, "rb").read() == want


def test_packed_container_roundtrip_and_cross_reading(host_built, tmp_path):
    """Extension: the bit-packed .quant container.  The Python mirror and the C++ CLI must read each
    other's packed files, decode them to the same image as the byte-aligned container, and the payload must have the
    size CompressedImage::sizeInBits() has always reported."""
    import quant_b200 as qb
    g = load_golden("kodim01_crop_2x2_n10")       # 10 bits per index: 2 bytes each in the reference's container
    ci = qb.CompressedImage()
    ci.codeVectors, ci.assignedCodeVector = g.z["codebook_bytes"], g.z["assign"].astype(np.uint64)
    ci.xSize, ci.ySize, ci.blockWidth, ci.blockHeight, ci.colorSpace = g.xs, g.ys, g.w, g.h, qb.ColorSpaces.SCALED
    plain, packed = ci.to_bytes(), ci.to_bytes_packed()
    assert hashlib.sha256(plain).hexdigest() == str(g.z["quant_sha"])          # the reference's container is untouched
    hdr = packed.index(b"\n") + 1
    assert packed.startswith(b"QP1 10 ") and (len(packed) - hdr) * 8 == ci.sizeInBits()
    assert len(packed) < 0.75 * len(plain)
    pq, out = str(tmp_path / "p.quant"), str(tmp_path / "p.ppm")
    ci.saveToFilePacked(pq)
    back = qb.CompressedImage()
    back.loadFromFile(pq)
    assert np.array_equal(back.assignedCodeVector, ci.assignedCodeVector) and np.array_equal(back.codeVectors, ci.codeVectors)
    r = subprocess.run([QUANT, pq, "-o", out], capture_output=True, text=True)     # C++ reads the Python-written file
    assert r.returncode == 0, r.stderr
    assert hashlib.sha256(ppm_payload(out, g.xs, g.ys)).hexdigest() == str(g.z["decoded_sha"])
    for bits in (1, 3, 8, 13, 24):
        rng = np.random.default_rng(bits)
        a = rng.integers(0, 1 << bits, 1001).astype(np.uint64)
        s = qb.pack_indices(a, bits)
        assert s.size == (1001 * bits + 7) // 8 and np.array_equal(qb.unpack_indices(s, 1001, bits), a)


def test_entropy_container_code_length_limit_in_cpp(host_built, tmp_path):
    """Fibonacci index counts (optimal Huffman depth 34 > the 32-bit limit), a single used codevector, an empty image."""
    out = subprocess.run([HOST_TEST, "entropy", str(tmp_path / "e.quant")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_entropy_container_roundtrip_and_cross_reading(host_built, tmp_path):
    """Extension (second half): the Huffman-coded .quant container "QH1".  Python writes / C++ reads
    and decodes, C++ re-writes a plain file as QH1 (`quant a.quant -o b.quant --entropy`, host code only) / Python
    reads; both decode to the reference's image; the index payload is within 1 % + 1 byte of the indices' empirical
    entropy and smaller than the bit-packed container on a natural image."""
    import quant_b200 as qb
    g = load_golden("kodim01_crop_2x2_n10")
    ci = qb.CompressedImage()
    ci.codeVectors, ci.assignedCodeVector = g.z["codebook_bytes"], g.z["assign"].astype(np.uint64)
    ci.xSize, ci.ySize, ci.blockWidth, ci.blockHeight, ci.colorSpace = g.xs, g.ys, g.w, g.h, qb.ColorSpaces.SCALED
    plain, packed, coded = ci.to_bytes(), ci.to_bytes_packed(), ci.to_bytes_entropy()
    assert coded.startswith(b"QH1 10 ") and len(coded) < len(packed) < len(plain)
    a = ci.assignedCodeVector.astype(np.int64)
    p = np.bincount(a, minlength=1024) / a.size
    entropy_bits = float(-(p[p > 0] * np.log2(p[p > 0])).sum()) * a.size
    payload = int(coded[:coded.index(b"\n")].split()[-1])
    assert entropy_bits / 8 <= payload <= 1.01 * entropy_bits / 8 + 1 + a.size / 8     # Huffman: < 1 bit per index above
    f_plain, f_py, f_cpp, out = (str(tmp_path / n) for n in ("a.quant", "py.quant", "cpp.quant", "o.ppm"))
    open(f_plain, "wb").write(plain)
    ci.saveToFileEntropy(f_py)
    back = qb.CompressedImage()
    back.loadFromFile(f_py)
    assert np.array_equal(back.assignedCodeVector, ci.assignedCodeVector) and np.array_equal(back.codeVectors, ci.codeVectors)
    r = subprocess.run([QUANT, f_py, "-o", out], capture_output=True, text=True)            # C++ decodes the Python-written file
    assert r.returncode == 0, r.stderr
    assert hashlib.sha256(ppm_payload(out, g.xs, g.ys)).hexdigest() == str(g.z["decoded_sha"])
    r = subprocess.run([QUANT, f_plain, "-o", f_cpp, "--entropy"], capture_output=True, text=True)   # C++ writes QH1
    assert r.returncode == 0, r.stderr
    assert open(f_cpp, "rb").read(4) == b"QH1 " and os.path.getsize(f_cpp) < len(packed)
    back2 = qb.CompressedImage()
    back2.loadFromFile(f_cpp)                                                              # Python reads the C++-written file
    assert np.array_equal(back2.assignedCodeVector, ci.assignedCodeVector) and np.array_equal(back2.codeVectors, ci.codeVectors)
    f_back = str(tmp_path / "back.quant")
    r = subprocess.run([QUANT, f_cpp, "-o", f_back], capture_output=True, text=True)       # ... and C++ turns it back into
    assert r.returncode == 0, r.stderr                                                     # the reference's container
    assert hashlib.sha256(open(f_back, "rb").read()).hexdigest() == str(g.z["quant_sha"])
    # degenerate histograms: one symbol only, two symbols, a long geometric tail (length limit), empty
    for counts in ([0, 7, 0, 0], [3, 0, 9, 0], [2 ** i for i in range(40)], [0, 0]):
        L = qb.huffman_lengths(counts)
        assert L.max(initial=0) <= 32 and sum(2.0 ** -int(l) for l in L if l) <= 1.0
        rng = np.random.default_rng(len(counts))
        used = np.nonzero(counts)[0]
        idx = rng.choice(used, 500).astype(np.uint64) if used.size else np.zeros(0, np.uint64)
        assert np.array_equal(qb.huffman_decode(qb.huffman_encode(idx, L), idx.size, L), idx)
    # a corrupt stream is an error, not garbage
    bad = bytearray(coded)
    del bad[-200:]
    open(f_py, "wb").write(bytes(bad))
    assert subprocess.run([QUANT, f_py, "-o", out], capture_output=True).returncode != 0


@pytest.mark.gpu
def test_cli_pack_flag_and_device_side_packing(host_built, tmp_path):
    import quant_b200 as qb
    g = load_golden("odd_101x67_3x2_n5")
    p, q, q2, d = str(tmp_path / "i.ppm"), str(tmp_path / "o.quant"), str(tmp_path / "o2.quant"), str(tmp_path / "o.ppm")
    write_ppm(p, g.rgb, g.xs, g.ys)
    args = ["-n", str(g.nbits), "-w", str(g.w), "-h", str(g.h), "--c", str(g.cs)]
    assert subprocess.run([QUANT, p, "-o", q, "--pack"] + args, capture_output=True).returncode == 0
    back = qb.CompressedImage()
    back.loadFromFile(q)                                   # Python reads the C++-written packed file
    assert np.array_equal(back.assignedCodeVector, g.z["assign"].astype(np.uint64))
    assert np.array_equal(back.codeVectors, g.z["codebook_bytes"])
    back.saveToFile(q2)                                    # ... and rewrites the reference's container from it
    assert hashlib.sha256(open(q2, "rb").read()).hexdigest() == str(g.z["quant_sha"])
    assert subprocess.run([QUANT, q, "-o", d], capture_output=True).returncode == 0
    assert hashlib.sha256(ppm_payload(d, g.xs, g.ys)).hexdigest() == str(g.z["decoded_sha"])
    ctx = qb.Context(0)
    ctx.set_image(g.rgb, g.xs, g.ys, g.w, g.h, g.cs)
    ctx.train(g.nbits)
    for bits in (g.nbits, g.nbits + 3, 17, 32):            # packed on the device == packed on the host
        assert np.array_equal(ctx.get_assign_packed(bits), qb.pack_indices(ctx.get_assign(), bits))
    with pytest.raises(qb.Qb200Error):
        ctx.get_assign_packed(0)
    ctx.close()
