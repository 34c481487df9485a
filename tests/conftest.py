import glob
import os
import subprocess
import sys

import numpy as np
import pytest

# Several ranks of the multi-device tests share ONE GPU on a single-GPU box, and a rank waits inside a kernel for its
# peers' kernels: their streams must not be folded onto the same hardware queue (default: 8 queues per context).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu():
    try:
        import ctypes
        lib = ctypes.CDLL("libcuda.so.1")
        if lib.cuInit(0) != 0:
            return False
        n = ctypes.c_int()
        return lib.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def port():
    """The plain-C oracle (built on demand: compiling the checker is not using it)."""
    so = os.path.join(ROOT, "oracle", "liblbg_oracle.so")
    src = os.path.join(ROOT, "oracle", "lbg_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    from oracle.pyoracle import PortLib
    return PortLib()


@pytest.fixture(scope="session")
def reflib():
    """The real reference (strict build) - only where oracle/_ref was built (the build container)."""
    from oracle.pyoracle import RefLib, have_ref
    if not have_ref("strict"):
        pytest.skip("oracle/_ref not built here")
    return RefLib("strict")


def golden_names():
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    # *_cie / generic_*: FP64-vector fixtures (no byte lattice), covered by their own tests
    return [n for n in names if n not in ("letters_layout", "encode_only_k1024") and not n.endswith("_cie")
            and not n.startswith("generic_") and not n.startswith("regress_")]


class Golden:
    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.z = z
        self.xs, self.ys, self.w, self.h, self.nbits, self.cs = (int(v) for v in z["params"])
        self.rgb = z["rgb"]
        self.dim = 3 * self.w * self.h
        self.N = ((self.xs + self.w - 1) // self.w) * ((self.ys + self.h - 1) // self.h)
        self.has_levels = "L0_pre" in z.files

    def level(self, i):
        z = self.z
        d = z[f"L{i}_d"]
        return dict(K=2 << i, cb_pre=z[f"L{i}_pre"], cb_post=z[f"L{i}_post"],
                    assign=z[f"L{i}_assign"].astype(np.uint64), d0=float(d[0]), d1=float(d[1]))


def load_golden(name):
    return Golden(name)


@pytest.fixture(scope="session")
def gpu_ctx():
    import quant_b200 as qb
    ctx = qb.Context(0)
    yield ctx
    ctx.close()
