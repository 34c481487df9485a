"""The C-ABI shared library: loads, exports every symbol include/qb200.h declares, and its
host-only entry points agree with the oracle.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import quant_b200 as qb
from quant_b200 import _lib
from conftest import HAVE_GPU, ROOT, load_golden


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "qb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qb200_[a-z0-9_]+)\s*\(", src)) - {"qb200_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    lib = qb.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/qb200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "python binding and header disagree"
    assert lib.qb200_version() == 200


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libqb200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


@pytest.mark.skipif(HAVE_GPU, reason="this box has a GPU")
def test_no_device_is_an_error_not_a_fallback():
    with pytest.raises(qb.Qb200Error) as e:
        qb.Context(0)
    assert e.value.code == _lib.ERR_NODEV


def test_null_arguments_are_rejected():
    lib = qb.load()
    assert lib.qb200_create(0, None) == _lib.ERR_ARG
    assert lib.qb200_codebook_to_bytes(None, 1, 3, 1, None) == _lib.ERR_ARG
    assert lib.qb200_get_assign(None, None) == _lib.ERR_ARG
    assert lib.qb200_num_vectors(None) == 0


@pytest.mark.parametrize("cs", [0, 1])
def test_codebook_to_bytes_matches_oracle(port, cs):
    rng = np.random.default_rng(3)
    if cs == 1:
        cb = rng.random((300, 12)) * 1.3 - 0.1
        cb[:40] = (rng.integers(0, 255, (40, 12)) + 0.5) / 255.0     # exact halves
        cb[40] = 0.0                                                 # dead cell
    else:
        cb = rng.random((300, 12)) * 255 - 128
        cb[:40] = rng.integers(-128, 127, (40, 12)) + 0.5
    assert np.array_equal(qb.codebook_to_bytes(cb, cs), port.codebook_to_bytes(cb, cs))


def test_kd_build_matches_oracle_tree(port):
    g = load_golden("kodim01_crop_2x2_n10")
    for i in (3, 6, 9):
        cb = g.level(i)["cb_pre"]
        K = cb.shape[0]
        order = np.empty(K, np.uint32)
        nn, dp = C.c_int(), C.c_int()
        rc = qb.load().qb200_debug_kd_build(cb.ctypes.data_as(C.c_void_p), K, cb.shape[1],
                                            order.ctypes.data_as(C.c_void_p), C.byref(nn), C.byref(dp))
        assert rc == 0
        depth, nodes, oorder = port.kd_order(cb)
        assert (dp.value, nn.value) == (depth, nodes)
        assert np.array_equal(order, oorder)
    # degenerate: all points identical -> balanced count/2 splits
    cb = np.zeros((64, 12))
    order = np.empty(64, np.uint32)
    nn, dp = C.c_int(), C.c_int()
    assert qb.load().qb200_debug_kd_build(cb.ctypes.data_as(C.c_void_p), 64, 12,
                                          order.ctypes.data_as(C.c_void_p), C.byref(nn), C.byref(dp)) == 0
    depth, nodes, oorder = port.kd_order(cb)
    assert (dp.value, nn.value) == (depth, nodes) and np.array_equal(order, oorder)


def test_kd_tree_robustness_margin():
    """The census behind the auto centroid mode's trust in the tree's visiting order (qb200_debug_kd_margin): generic
    points give a comfortable margin, duplicated non-zero points none, and the same duplicates count as harmless when
    they are flagged bit-reproducible or are dead cells' zero vectors."""
    import ctypes as C
    import numpy as np
    from quant_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)

    def margin(pts, flags=None):
        pts = np.ascontiguousarray(pts, np.float64)
        m = C.c_double()
        f = None if flags is None else np.ascontiguousarray(flags, np.uint8).ctypes.data_as(C.c_void_p)
        assert lib.qb200_debug_kd_margin(pts.ctypes.data_as(C.c_void_p), pts.shape[0], pts.shape[1], f, C.byref(m)) == 0
        return m.value

    pts = rng.random((500, 12))
    assert margin(pts) > 1e-9
    dup = pts.copy()
    dup[100:140] = dup[100]                       # 40 copies of one non-zero point
    assert margin(dup) == 0.0
    flags = np.zeros(500, np.uint8)
    flags[100:140] = 1
    assert margin(dup, flags) > 1e-9
    dead = pts.copy()
    dead[200:260] = 0.0                           # dead cells
    assert margin(dead) > 1e-9


def test_kd_census_compares_the_winning_spread_with_every_eligible_dimension():
    """Fuzz find of round 2: three dimensions share the root's largest spread exactly (attained by bit-reproducible
    points); a point that is NOT bit-reproducible sits one ulp below the maximum of the last of them.  In the reference's
    codebook that point may be one ulp above and make that dimension the root cut, so the census must report no margin -
    also when an earlier runner-up with the same spread is entirely bit-reproducible."""
    import ctypes as C
    import numpy as np
    from quant_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(11)
    pts = rng.random((64, 3)) * 0.8 + 0.1          # everything inside (0.1, 0.9)
    pts[0] = [0.0, 0.0, 0.0]
    pts[1] = [1.0, 1.0, 1.0]                      # spreads of all three dimensions: exactly 1.0
    flags = np.ones(64, np.uint8)

    def margin(p, f):
        p = np.ascontiguousarray(p, np.float64)
        m = C.c_double()
        assert lib.qb200_debug_kd_margin(p.ctypes.data_as(C.c_void_p), p.shape[0], p.shape[1],
                                         np.ascontiguousarray(f, np.uint8).ctypes.data_as(C.c_void_p), C.byref(m)) == 0
        return m.value

    assert margin(pts, flags) > 1e-9               # all three ties are between numbers both codebooks share
    near = pts.copy()
    near[5, 2] = np.nextafter(1.0, 0.0)            # one ulp below the maximum of the LAST dimension ...
    f2 = flags.copy()
    f2[5] = 0                                      # ... and not bit-reproducible
    assert margin(near, f2) < 1e-12
    f2[5] = 1
    assert margin(near, f2) > 1e-9                 # the same point, bit-reproducible: harmless
