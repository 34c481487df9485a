"""Thin object wrapper over the C ABI: one ``Context`` per GPU (include/qb200.h)."""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np

from . import _lib
from ._lib import (ALLREDUCE_FN, CS_SCALED, LevelReport, MODE_PARITY,
                   Qb200Error)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """Owns one ``qb200_ctx``. ``device`` is the CUDA ordinal; ``devices`` (a list of ordinals, or "all") makes a
    multi-device context instead (qb200_create_multi: one process driving several GPUs)."""

    def __init__(self, device: int = 0, devices=None):
        self.lib = _lib.load()
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.qb200_create(device, C.byref(h))
        elif isinstance(devices, str) and devices == "all":
            rc = self.lib.qb200_create_multi(0, None, C.byref(h))
        else:
            ids = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self.lib.qb200_create_multi(len(devices), ids, C.byref(h))
        if rc != 0:
            raise Qb200Error(rc, (self.lib.qb200_last_error(None) or b"").decode())
        self.h = h
        self.device = device
        self._keep = None  # keeps host/device buffers alive while the context borrows them

    # -- plumbing ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.qb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise Qb200Error(rc, (self.lib.qb200_last_error(self.h) or b"").decode())

    def set_stream(self, cuda_stream: int | None):
        self._check(self.lib.qb200_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def set_tensor_cores(self, enable: bool):
        """Filter engine for K >= 16: tcgen05 tensor-core kernel (default) or the FP32 CUDA-core kernel."""
        self._check(self.lib.qb200_set_tensor_cores(self.h, 1 if enable else 0))

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self._check(self.lib.qb200_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi),
                                               C.byref(mem)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), total_mem=mem.value)

    # -- training set -----------------------------------------------------------------------
    def set_image(self, rgb: np.ndarray, xSize: int, ySize: int, w: int, h: int,
                  colorspace: int = CS_SCALED, n_images: int = 1):
        rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1)
        if rgb.size != n_images * xSize * ySize * 3:
            raise ValueError("rgb has the wrong number of bytes")
        self._keep = rgb
        self._check(self.lib.qb200_set_image(self.h, _ptr(rgb), xSize, ySize, w, h, colorspace,
                                             n_images, 0))

    def set_image_device(self, dev_ptr: int, xSize: int, ySize: int, w: int, h: int,
                         colorspace: int = CS_SCALED, n_images: int = 1, keep=None):
        self._keep = keep
        self._check(self.lib.qb200_set_image(self.h, C.c_void_p(dev_ptr), xSize, ySize, w, h,
                                             colorspace, n_images, 1))

    def set_image_shard(self, rgb: np.ndarray, xSize: int, ySize: int, w: int, h: int,
                        colorspace: int, row_begin: int, row_end: int):
        rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1)
        self._keep = rgb
        self._check(self.lib.qb200_set_image_shard(self.h, _ptr(rgb), xSize, ySize, w, h,
                                                   colorspace, row_begin, row_end))

    def set_image_band(self, band, xSize: int, ySize: int, w: int, h: int, colorspace: int,
                       row_begin: int, row_end: int, device_ptr: int | None = None,
                       nbytes: int | None = None, keep=None):
        """This rank's own bytes of a sharded image: a host array, or (device_ptr, nbytes)."""
        if device_ptr is not None:
            self._keep = keep
            self._check(self.lib.qb200_set_image_band(self.h, C.c_void_p(device_ptr), nbytes, 1, xSize,
                                                      ySize, w, h, colorspace, row_begin, row_end))
            return
        band = np.ascontiguousarray(band, np.uint8).reshape(-1)
        self._keep = band
        self._check(self.lib.qb200_set_image_band(self.h, _ptr(band), band.size, 0, xSize, ySize, w, h,
                                                  colorspace, row_begin, row_end))

    def set_vectors_u8(self, mat: np.ndarray, colorspace: int = CS_SCALED):
        mat = np.ascontiguousarray(mat, np.uint8)
        n, dim = mat.shape
        self._keep = mat
        self._check(self.lib.qb200_set_vectors_u8(self.h, _ptr(mat), n, dim, colorspace, 0))

    def set_vectors_f64(self, mat: np.ndarray):
        """General FP64 training vectors (no byte lattice): qb200_set_vectors_f64."""
        mat = np.ascontiguousarray(mat, np.float64)
        n, dim = mat.shape
        self._keep = mat
        self._check(self.lib.qb200_set_vectors_f64(self.h, _ptr(mat), n, dim, 0))

    @property
    def num_vectors(self) -> int:
        return int(self.lib.qb200_num_vectors(self.h))

    @property
    def dim(self) -> int:
        return int(self.lib.qb200_dim(self.h))

    # -- hot path ---------------------------------------------------------------------------
    def set_exact_centroids(self, enable):
        """Bit-exact centroids: run the reference's compensated member sums (src/Quantizer.cpp:59-70) instead of
        deriving the centroid from integer sums; see include/qb200.h.  False/0 integer sums, True/1 compensated sums
        (parallel evaluation), 2 the same as a literal sequential chain (for comparison), 3 or "auto" (the default):
        integer sums unless the train had tie-sensitive decisions, then repeated with the compensated sums."""
        self._check(self.lib.qb200_set_exact_centroids(self.h, 3 if enable == "auto" else int(enable)))

    @property
    def last_train_exact(self) -> bool:
        return bool(self.lib.qb200_last_train_exact(self.h))

    def debug_level_codebook(self, level: int) -> np.ndarray:
        """Diagnostics: the 2^(level+1) codevectors split level `level` of the last train started from."""
        out = np.empty((2 << level, self.dim), np.float64)
        self._check(self.lib.qb200_debug_level_codebook(self.h, level, out.ctypes.data_as(C.c_void_p)))
        return out

    def comm_export(self, max_words: int = 0) -> bytes:
        """This context's all-reduce exchange block as a CUDA IPC handle (64 bytes) - see qb200_comm_export."""
        buf = C.create_string_buffer(64)
        self._check(self.lib.qb200_comm_export(self.h, max_words, buf))
        return buf.raw

    def comm_attach(self, world: int, rank: int, handles: bytes):
        """Joins the group: ``handles`` = every rank's 64-byte handle, concatenated in rank order."""
        if len(handles) != 64 * world:
            raise ValueError("handles must hold 64 bytes per rank")
        self._check(self.lib.qb200_comm_attach(self.h, world, rank, C.c_char_p(handles)))

    def allreduce_u64(self, dev_ptr: int, count: int):
        self._check(self.lib.qb200_allreduce_u64(self.h, C.c_void_p(dev_ptr), count))

    def set_seed(self, seed: int):
        self._check(self.lib.qb200_set_seed(self.h, seed))

    def set_rank(self, rank: int, world: int):
        self._check(self.lib.qb200_set_rank(self.h, rank, world))

    def train(self, nbits: int, eps: float = float(np.float32(1e-6)), n_total: int = 0,
              allreduce: Optional[Callable[[int, int, int], int]] = None, reports: bool = True,
              mode: int = MODE_PARITY):
        """LBGQuantizer::quantize. Returns (codebook[K,dim] f64, distortion, [LevelReport...])."""
        K, dim = 1 << nbits, self.dim
        cb = np.empty((K, dim), np.float64)
        dist = C.c_double()
        rep = (LevelReport * max(nbits, 1))() if reports else None
        cb_fn = None
        if allreduce is not None:
            def tramp(dev, count, stream, user):
                try:
                    return int(allreduce(dev, count, stream or 0) or 0)
                except Exception:  # never let an exception cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return 1
            cb_fn = ALLREDUCE_FN(tramp)
        self._check(self.lib.qb200_train(self.h, nbits, eps, mode, n_total,
                                         C.cast(cb_fn, C.c_void_p) if cb_fn else None, None,
                                         _ptr(cb), C.byref(dist),
                                         C.cast(rep, C.c_void_p) if rep is not None else None))
        out = []
        if rep is not None:
            for i in range(nbits):
                r = rep[i]
                out.append({f: getattr(r, f) for f, _ in LevelReport._fields_})
        return cb, dist.value, out

    def get_assign(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Indices as uint32; ``out`` may be a caller-owned (e.g. pinned) uint32 array."""
        a = np.empty(self.num_vectors, np.uint32) if out is None else out
        if a.dtype != np.uint32 or a.size < self.num_vectors or not a.flags.c_contiguous:
            raise ValueError("out must be a contiguous uint32 array of num_vectors elements")
        self._check(self.lib.qb200_get_assign(self.h, _ptr(a)))
        return a

    def get_assign_u64(self) -> np.ndarray:
        a = np.empty(self.num_vectors, np.uint64)
        self._check(self.lib.qb200_get_assign_u64(self.h, _ptr(a)))
        return a

    def get_assign_packed(self, bits: int) -> np.ndarray:
        """The assignment as an LSB-first bit stream, `bits` per index, packed on the device (qb200_get_assign_packed)."""
        out = np.empty((self.num_vectors * bits + 7) // 8, np.uint8)
        self._check(self.lib.qb200_get_assign_packed(self.h, bits, _ptr(out), out.size))
        return out

    def assign_device_ptr(self) -> int:
        p = C.c_void_p()
        self._check(self.lib.qb200_assign_device_ptr(self.h, C.byref(p)))
        return p.value or 0

    def assign_accumulate(self, codebook: np.ndarray, want_assign=True, want_stats=True):
        cb = np.ascontiguousarray(codebook, np.float64)
        K, dim = cb.shape
        if dim != self.dim:
            raise ValueError("codebook dimension mismatch")
        a = np.empty(self.num_vectors, np.uint32) if want_assign else None
        n = np.empty(K, np.uint64) if want_stats else None
        S = np.empty((K, dim), np.int64) if want_stats else None
        Q = np.empty(K, np.uint64) if want_stats else None
        fl = C.c_uint32()
        self._check(self.lib.qb200_assign_accumulate(self.h, _ptr(cb), K, _ptr(a), _ptr(n), _ptr(S),
                                                     _ptr(Q), C.byref(fl)))
        return dict(assign=a, count=n, sum=S, sqsum=Q, flagged=fl.value)

    def assign_only(self, codebook: np.ndarray):
        cb = np.ascontiguousarray(codebook, np.float64)
        fl, ma, mr = C.c_uint32(), C.c_float(), C.c_float()
        self._check(self.lib.qb200_assign_only(self.h, _ptr(cb), cb.shape[0], C.byref(fl),
                                               C.byref(ma), C.byref(mr)))
        return dict(flagged=fl.value, ms_assign=ma.value, ms_resolve=mr.value)

    def decode(self, codebook_bytes: np.ndarray, want_image=True):
        cbb = np.ascontiguousarray(codebook_bytes, np.uint8)
        if cbb.size % self.dim:
            raise ValueError(f"codebook bytes ({cbb.size}) are not a multiple of the dimension {self.dim}")
        out = np.empty(len(self._keep) if self._keep is not None else 0, np.uint8) if want_image else None
        mse = C.c_double()
        self._check(self.lib.qb200_decode(self.h, _ptr(cbb), cbb.size // self.dim, _ptr(out), C.byref(mse)))
        return out, mse.value

    def filter_records(self) -> np.ndarray:
        """(num_vectors, 4) float32 records of the last tensor-core filter pass: best, second, chunk bits, 0."""
        out = np.empty((self.num_vectors, 4), np.float32)
        self._check(self.lib.qb200_debug_filter_records(self.h, _ptr(out)))
        return out

    def measure_fp32_peak(self) -> float:
        v = C.c_double()
        self._check(self.lib.qb200_measure_fp32_peak(self.h, C.byref(v)))
        return v.value


def finalize_level(colorspace, n_total, count, sums, sqsum, codebook_pre=None):
    lib = _lib.load()
    count = np.ascontiguousarray(count, np.uint64)
    sums = np.ascontiguousarray(sums, np.int64)
    sqsum = np.ascontiguousarray(sqsum, np.uint64)
    K, dim = sums.shape
    post = np.empty((K, dim), np.float64)
    pre = None if codebook_pre is None else np.ascontiguousarray(codebook_pre, np.float64)
    d0, d1 = C.c_double(float("nan")), C.c_double()
    rc = lib.qb200_finalize_level(colorspace, K, dim, n_total, _ptr(count), _ptr(sums), _ptr(sqsum),
                                  _ptr(pre), _ptr(post), C.byref(d0), C.byref(d1))
    if rc != 0:
        raise Qb200Error(rc, "qb200_finalize_level")
    return post, d0.value, d1.value


def codebook_to_bytes(codebook, colorspace=CS_SCALED):
    lib = _lib.load()
    cb = np.ascontiguousarray(codebook, np.float64)
    out = np.empty(cb.shape, np.uint8)
    rc = lib.qb200_codebook_to_bytes(_ptr(cb), cb.shape[0], cb.shape[1], colorspace, _ptr(out))
    if rc != 0:
        raise Qb200Error(rc, "qb200_codebook_to_bytes")
    return out


def launch_count(reset: bool = False) -> int:
    return int(_lib.load().qb200_launch_count(1 if reset else 0))
