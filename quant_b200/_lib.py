"""ctypes binding of libqb200.so (C ABI: include/qb200.h).

There is no fallback of any kind: if the shared library has not been built
(``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C quant_b200/csrc``) importing
the product fails loudly, and so does creating a context on a machine without a B200.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqb200.so")

OK, ERR_ARG, ERR_NODEV, ERR_CUDA, ERR_OOM, ERR_STATE, ERR_COMM = 0, -1, -2, -3, -4, -5, -6
CS_NORMAL, CS_SCALED, CS_CIE1931 = 0, 1, 2
MODE_PARITY, MODE_FULL, MODE_FULL_REPAIR = 0, 1, 2


class LevelReport(C.Structure):
    _fields_ = [("K", C.c_uint32), ("flagged", C.c_uint32), ("changed", C.c_uint32),
                ("ties", C.c_uint32), ("dead_cells", C.c_uint32), ("kd_depth", C.c_uint32), ("iterations", C.c_uint32), ("repaired", C.c_uint32),
                ("ms_assign", C.c_float),
                ("ms_resolve", C.c_float), ("ms_accumulate", C.c_float), ("refiltered", C.c_uint32),
                ("distortion_pre", C.c_double), ("distortion_post", C.c_double),
                ("sensitive", C.c_uint32), ("reserved", C.c_uint32)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)

# name -> (restype, argtypes); every symbol include/qb200.h declares
SIGNATURES = {
    "qb200_version": (C.c_int, []),
    "qb200_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "qb200_create_multi": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "qb200_comm_export": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "qb200_comm_attach": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "qb200_allreduce_u64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "qb200_destroy": (None, [C.c_void_p]),
    "qb200_last_error": (C.c_char_p, [C.c_void_p]),
    "qb200_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "qb200_set_exact_centroids": (C.c_int, [C.c_void_p, C.c_int]),
    "qb200_last_train_exact": (C.c_int, [C.c_void_p]),
    "qb200_get_assign_packed": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "qb200_set_seed": (C.c_int, [C.c_void_p, C.c_uint64]),
    "qb200_set_rank": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "qb200_set_tensor_cores": (C.c_int, [C.c_void_p, C.c_int]),
    "qb200_device_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "qb200_set_image": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int]),
    "qb200_set_image_shard": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_size_t, C.c_size_t]),
    "qb200_set_image_band": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t]),
    "qb200_set_vectors_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "qb200_set_vectors_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                       C.c_int]),
    "qb200_num_vectors": (C.c_size_t, [C.c_void_p]),
    "qb200_dim": (C.c_int, [C.c_void_p]),
    "qb200_train": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_uint64, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "qb200_get_assign": (C.c_int, [C.c_void_p, C.c_void_p]),
    "qb200_get_assign_u64": (C.c_int, [C.c_void_p, C.c_void_p]),
    "qb200_assign_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "qb200_assign_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_uint32)]),
    "qb200_assign_only": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "qb200_finalize_level": (C.c_int, [C.c_int, C.c_uint32, C.c_int, C.c_uint64, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "qb200_codebook_to_bytes": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]),
    "qb200_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                               C.POINTER(C.c_double)]),
    "qb200_measure_fp32_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "qb200_debug_kd_build": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                       C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "qb200_debug_kd_tree": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int),
                                      C.c_void_p, C.POINTER(C.c_double)]),
    "qb200_debug_level_codebook": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "qb200_debug_kd_margin": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(C.c_double)]),
    "qb200_debug_filter_records": (C.c_int, [C.c_void_p, C.c_void_p]),
    "qb200_launch_count": (C.c_int, [C.c_int]),
}

_lib = None


class Qb200Error(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"libqb200 error {code}: {text}")
        self.code = code


def load() -> C.CDLL:
    """Loads libqb200.so; raises if it is missing (there is no CPU or PyTorch fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C quant_b200/csrc` "
            "(or __graft_entry__.build()). quant_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
