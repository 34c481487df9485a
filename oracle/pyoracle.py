"""TEST INFRASTRUCTURE ONLY - ctypes loaders for the two CPU oracles.

* ``RefLib``   : oracle/_ref/libquantref_{strict,release}.so - the UNMODIFIED reference sources
                 (built by ``make -C oracle ref`` in the container that has /root/reference).
* ``PortLib``  : oracle/liblbg_oracle.so - the plain-C restatement (oracle/lbg_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  The product (quant_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")

NORMAL, SCALED, CIE1931 = 0, 1, 2


def n_vectors(xs: int, ys: int, w: int, h: int) -> int:
    return ((xs + w - 1) // w) * ((ys + h - 1) // h)


def level_offsets(nbits: int, dim: int):
    """Offsets (in doubles) of level l=1..nbits inside the concatenated per-level codebooks."""
    off, out = 0, []
    for l in range(1, nbits + 1):
        out.append(off)
        off += (1 << l) * dim
    return out, off


class RefLib:
    """The real reference, compiled from its own sources."""

    def __init__(self, flavour: str = "strict"):
        path = os.path.join(HERE, "_ref", f"libquantref_{flavour}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.flavour = flavour
        self.lib = L = C.CDLL(path)
        L.ref_blocks.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p]
        L.ref_codebook_to_bytes.argtypes = [_f64p, C.c_size_t, C.c_int, C.c_int, _u8p]
        L.ref_decode.argtypes = [_u8p, C.c_size_t, _u64p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                 C.c_int, _u8p]
        L.ref_quantize.argtypes = [_f64p, C.c_size_t, C.c_int, C.c_int, C.c_double, _f64p, _u64p,
                                   C.POINTER(C.c_double)]
        L.ref_nn.argtypes = [_f64p, C.c_size_t, C.c_int, _f64p, C.c_size_t, _u64p]
        L.ref_nn_rgb.argtypes = [_f64p, C.c_size_t, _u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, _u64p]
        L.ref_levels.argtypes = [_f64p, C.c_size_t, C.c_int, C.c_int, _f64p, _f64p, _u64p, _f64p,
                                 _f64p, _f64p]
        L.ref_compress.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                   C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                                   C.POINTER(C.c_float), C.POINTER(C.c_double)]
        L.ref_compress_to_file.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_double, C.c_int, C.c_char_p]
        L.ref_decompress_file.argtypes = [C.c_char_p, _u8p, C.c_size_t, C.POINTER(C.c_int),
                                          C.POINTER(C.c_int)]
        L.ref_set_threads.argtypes = [C.c_int]

    def max_threads(self) -> int:
        return self.lib.ref_max_threads()

    def set_threads(self, n: int) -> None:
        self.lib.ref_set_threads(n)

    def blocks(self, rgb, xs, ys, w, h, cs=SCALED):
        n, dim = n_vectors(xs, ys, w, h), 3 * w * h
        out = np.empty((n, dim), np.float64)
        got = self.lib.ref_blocks(np.ascontiguousarray(rgb, np.uint8).ravel(), xs, ys, w, h, cs,
                                  out.reshape(-1))
        assert got == n
        return out

    def codebook_to_bytes(self, cb, cs=SCALED):
        cb = np.ascontiguousarray(cb, np.float64)
        out = np.empty(cb.shape, np.uint8)
        self.lib.ref_codebook_to_bytes(cb.reshape(-1), cb.shape[0], cb.shape[1], cs, out.reshape(-1))
        return out

    def decode(self, cb_bytes, assign, xs, ys, w, h):
        out = np.empty(xs * ys * 3, np.uint8)
        cb_bytes = np.ascontiguousarray(cb_bytes, np.uint8)
        a = np.ascontiguousarray(assign, np.uint64)
        self.lib.ref_decode(cb_bytes.reshape(-1), cb_bytes.shape[0], a, a.size, xs, ys, w, h, out)
        return out

    def quantize(self, X, nbits, eps=float(np.float32(1e-6))):
        X = np.ascontiguousarray(X, np.float64)
        n, dim = X.shape
        cb = np.empty(((1 << nbits), dim), np.float64)
        a = np.empty(n, np.uint64)
        d = C.c_double()
        self.lib.ref_quantize(X.reshape(-1), n, dim, nbits, eps, cb.reshape(-1), a, C.byref(d))
        return cb, a, d.value

    def nn(self, cb, Q):
        cb = np.ascontiguousarray(cb, np.float64)
        Q = np.ascontiguousarray(Q, np.float64)
        out = np.empty(Q.shape[0], np.uint64)
        self.lib.ref_nn(cb.reshape(-1), cb.shape[0], cb.shape[1], Q.reshape(-1), Q.shape[0], out)
        return out

    def nn_rgb(self, cb, rgb, xs, ys, w, h, cs=SCALED):
        cb = np.ascontiguousarray(cb, np.float64)
        out = np.empty(n_vectors(xs, ys, w, h), np.uint64)
        self.lib.ref_nn_rgb(cb.reshape(-1), cb.shape[0],
                            np.ascontiguousarray(rgb, np.uint8).ravel(), xs, ys, w, h, cs, out)
        return out

    def levels(self, X, nbits):
        """Per-level replay; raises if it does not reproduce quantize() bit-for-bit."""
        X = np.ascontiguousarray(X, np.float64)
        n, dim = X.shape
        offs, tot = level_offsets(nbits, dim)
        cb0 = np.empty(dim, np.float64)
        pre = np.empty(tot, np.float64)
        post = np.empty(tot, np.float64)
        a = np.empty((nbits, n), np.uint64)
        d0 = np.empty(nbits, np.float64)
        d1 = np.empty(nbits, np.float64)
        rc = self.lib.ref_levels(X.reshape(-1), n, dim, nbits, cb0, pre, a.reshape(-1), post, d0, d1)
        if rc != 0:
            raise RuntimeError("per-level replay diverged from the reference's quantize()")
        lv = []
        for l in range(1, nbits + 1):
            K, o = 1 << l, offs[l - 1]
            lv.append(dict(K=K, cb_pre=pre[o:o + K * dim].reshape(K, dim).copy(),
                           cb_post=post[o:o + K * dim].reshape(K, dim).copy(),
                           assign=a[l - 1].copy(), d0=d0[l - 1], d1=d1[l - 1]))
        return cb0, lv

    def compress(self, rgb, xs, ys, w, h, nbits, cs=SCALED, eps=float(np.float32(1e-6)),
                 want_outputs=True):
        n, dim, K = n_vectors(xs, ys, w, h), 3 * w * h, 1 << nbits
        rgb = np.ascontiguousarray(rgb, np.uint8).ravel()
        cbb = np.empty((K, dim), np.uint8) if want_outputs else None
        a = np.empty(n, np.uint64) if want_outputs else None
        d, bpp, sec = C.c_double(), C.c_float(), C.c_double()
        self.lib.ref_compress(rgb, xs, ys, cs, w, h, eps, nbits,
                              cbb.ctypes.data if want_outputs else None,
                              a.ctypes.data if want_outputs else None,
                              C.byref(d), C.byref(bpp), C.byref(sec))
        return dict(codebook_bytes=cbb, assign=a, distortion=d.value, bpp=bpp.value,
                    seconds=sec.value)

    def compress_to_file(self, rgb, xs, ys, w, h, nbits, path, cs=SCALED,
                         eps=float(np.float32(1e-6))):
        self.lib.ref_compress_to_file(np.ascontiguousarray(rgb, np.uint8).ravel(), xs, ys, cs, w, h,
                                      eps, nbits, path.encode())

    def decompress_file(self, path, cap):
        out = np.empty(cap, np.uint8)
        xs, ys = C.c_int(), C.c_int()
        rc = self.lib.ref_decompress_file(path.encode(), out, cap, C.byref(xs), C.byref(ys))
        assert rc == 0
        return out[: xs.value * ys.value * 3], xs.value, ys.value


def have_ref(flavour: str = "strict") -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", f"libquantref_{flavour}.so"))


class PortLib:
    """The plain-C restatement (oracle/lbg_oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "liblbg_oracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `make -C oracle port` (or __graft_entry__.build())")
        self.lib = L = C.CDLL(path)
        _i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
        L.orc_num_vectors.restype = C.c_size_t
        L.orc_num_vectors.argtypes = [C.c_int] * 4
        L.orc_blocks.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p]
        L.orc_blocks_lattice.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i16p]
        L.orc_codebook_to_bytes.argtypes = [_f64p, C.c_size_t, C.c_int, C.c_int, _u8p]
        L.orc_decode.argtypes = [_u8p, _u64p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.orc_assign.argtypes = [_f64p, C.c_size_t, C.c_int, _f64p, C.c_size_t, _u64p]
        L.orc_assign_bruteforce.argtypes = L.orc_assign.argtypes
        L.orc_distortion.restype = C.c_double
        L.orc_distortion.argtypes = [_f64p, C.c_size_t, C.c_int, _f64p, _u64p]
        L.orc_training_sum.argtypes = [_f64p, C.c_size_t, C.c_int, _f64p]
        L.orc_fix.argtypes = [_f64p, C.c_size_t, C.c_int, _u64p, C.c_size_t, _f64p]
        L.orc_split.argtypes = [_f64p, C.c_size_t, C.c_int]
        L.orc_quantize.restype = C.c_size_t
        L.orc_quantize.argtypes = [_f64p, C.c_size_t, C.c_int, C.c_int, C.c_double, _f64p, _u64p,
                                   C.POINTER(C.c_double)] + [C.c_void_p] * 7
        L.orc_stats.argtypes = [_i16p, C.c_size_t, C.c_int, _u64p, C.c_size_t, _u64p, _i64p, _u64p]
        L.orc_centroids_from_stats.argtypes = [_u64p, _i64p, C.c_size_t, C.c_int, C.c_int, _f64p]
        L.orc_quant_serialize.restype = C.c_size_t
        L.orc_quant_serialize.argtypes = [_u8p, C.c_size_t, _u64p, C.c_size_t, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_kd_build.restype = C.c_void_p
        L.orc_kd_build.argtypes = [_f64p, C.c_size_t, C.c_int]
        L.orc_kd_free.argtypes = [C.c_void_p]
        L.orc_kd_depth.argtypes = [C.c_void_p]
        L.orc_kd_num_nodes.restype = C.c_size_t
        L.orc_kd_num_nodes.argtypes = [C.c_void_p]

    def blocks(self, rgb, xs, ys, w, h, cs=SCALED):
        out = np.empty((n_vectors(xs, ys, w, h), 3 * w * h), np.float64)
        self.lib.orc_blocks(np.ascontiguousarray(rgb, np.uint8).ravel(), xs, ys, w, h, cs,
                            out.reshape(-1))
        return out

    def blocks_lattice(self, rgb, xs, ys, w, h, cs=SCALED):
        out = np.empty((n_vectors(xs, ys, w, h), 3 * w * h), np.int16)
        self.lib.orc_blocks_lattice(np.ascontiguousarray(rgb, np.uint8).ravel(), xs, ys, w, h, cs,
                                    out.reshape(-1))
        return out

    def codebook_to_bytes(self, cb, cs=SCALED):
        cb = np.ascontiguousarray(cb, np.float64)
        out = np.empty(cb.shape, np.uint8)
        self.lib.orc_codebook_to_bytes(cb.reshape(-1), cb.shape[0], cb.shape[1], cs, out.reshape(-1))
        return out

    def decode(self, cb_bytes, assign, xs, ys, w, h):
        out = np.empty(xs * ys * 3, np.uint8)
        self.lib.orc_decode(np.ascontiguousarray(cb_bytes, np.uint8).reshape(-1),
                            np.ascontiguousarray(assign, np.uint64), xs, ys, w, h, out)
        return out

    def assign(self, X, cb, bruteforce=False):
        X = np.ascontiguousarray(X, np.float64)
        cb = np.ascontiguousarray(cb, np.float64)
        out = np.empty(X.shape[0], np.uint64)
        f = self.lib.orc_assign_bruteforce if bruteforce else self.lib.orc_assign
        f(X.reshape(-1), X.shape[0], X.shape[1], cb.reshape(-1), cb.shape[0], out)
        return out

    def distortion(self, X, cb, assign):
        X = np.ascontiguousarray(X, np.float64)
        return self.lib.orc_distortion(X.reshape(-1), X.shape[0], X.shape[1],
                                       np.ascontiguousarray(cb, np.float64).reshape(-1),
                                       np.ascontiguousarray(assign, np.uint64))

    def fix(self, X, assign, K):
        X = np.ascontiguousarray(X, np.float64)
        cb = np.empty((K, X.shape[1]), np.float64)
        self.lib.orc_fix(X.reshape(-1), X.shape[0], X.shape[1],
                         np.ascontiguousarray(assign, np.uint64), K, cb.reshape(-1))
        return cb

    def split(self, cb):
        cb = np.ascontiguousarray(cb, np.float64)
        K, dim = cb.shape
        out = np.empty((2 * K, dim), np.float64)
        out[:K] = cb
        self.lib.orc_split(out.reshape(-1), K, dim)
        return out

    def quantize(self, X, nbits, eps=float(np.float32(1e-6)), levels=False):
        X = np.ascontiguousarray(X, np.float64)
        n, dim = X.shape
        K = 1 << nbits
        cb = np.empty((K, dim), np.float64)
        a = np.empty(n, np.uint64)
        d = C.c_double()
        if not levels:
            self.lib.orc_quantize(X.reshape(-1), n, dim, nbits, eps, cb.reshape(-1), a, C.byref(d),
                                  *([None] * 7))
            return cb, a, d.value
        offs, tot = level_offsets(nbits, dim)
        cb0 = np.empty(dim, np.float64)
        pre = np.empty(tot, np.float64)
        post = np.empty(tot, np.float64)
        al = np.empty((nbits, n), np.uint64)
        d0 = np.empty(nbits, np.float64)
        d1 = np.empty(nbits, np.float64)
        its = np.empty(nbits, np.int32)
        self.lib.orc_quantize(X.reshape(-1), n, dim, nbits, eps, cb.reshape(-1), a, C.byref(d),
                              cb0.ctypes.data, pre.ctypes.data, al.ctypes.data, post.ctypes.data,
                              d0.ctypes.data, d1.ctypes.data, its.ctypes.data)
        lv = []
        for l in range(1, nbits + 1):
            Kl, o = 1 << l, offs[l - 1]
            lv.append(dict(K=Kl, cb_pre=pre[o:o + Kl * dim].reshape(Kl, dim).copy(),
                           cb_post=post[o:o + Kl * dim].reshape(Kl, dim).copy(),
                           assign=al[l - 1].copy(), d0=d0[l - 1], d1=d1[l - 1], iters=int(its[l - 1])))
        return cb, a, d.value, cb0, lv

    def stats(self, T, assign, K):
        T = np.ascontiguousarray(T, np.int16)
        n = np.empty(K, np.uint64)
        S = np.empty((K, T.shape[1]), np.int64)
        Q = np.empty(K, np.uint64)
        self.lib.orc_stats(T.reshape(-1), T.shape[0], T.shape[1],
                           np.ascontiguousarray(assign, np.uint64), K, n, S.reshape(-1), Q)
        return n, S, Q

    def centroids_from_stats(self, n, S, cs=SCALED):
        K, dim = S.shape
        cb = np.empty((K, dim), np.float64)
        self.lib.orc_centroids_from_stats(np.ascontiguousarray(n, np.uint64),
                                          np.ascontiguousarray(S, np.int64).reshape(-1), K, dim, cs,
                                          cb.reshape(-1))
        return cb

    def quant_serialize(self, cb_bytes, assign, xs, ys, w, h, cs_field=SCALED):
        cb_bytes = np.ascontiguousarray(cb_bytes, np.uint8)
        a = np.ascontiguousarray(assign, np.uint64)
        K = cb_bytes.shape[0]
        tot = self.lib.orc_quant_serialize(cb_bytes.reshape(-1), K, a, a.size, xs, ys, w, h, cs_field,
                                           None, 0)
        out = np.empty(tot, np.uint8)
        got = self.lib.orc_quant_serialize(cb_bytes.reshape(-1), K, a, a.size, xs, ys, w, h, cs_field,
                                           out.ctypes.data, tot)
        assert got == tot
        return out.tobytes()

    def kd_shape(self, cb):
        cb = np.ascontiguousarray(cb, np.float64)
        t = self.lib.orc_kd_build(cb.reshape(-1), cb.shape[0], cb.shape[1])
        r = (self.lib.orc_kd_depth(t), self.lib.orc_kd_num_nodes(t))
        self.lib.orc_kd_free(t)
        return r

    def kd_order(self, cb):
        cb = np.ascontiguousarray(cb, np.float64)
        t = self.lib.orc_kd_build(cb.reshape(-1), cb.shape[0], cb.shape[1])
        out = np.empty(cb.shape[0], np.uint32)
        self.lib.orc_kd_order.argtypes = [C.c_void_p, _u32p]
        self.lib.orc_kd_order(t, out)
        r = (self.lib.orc_kd_depth(t), self.lib.orc_kd_num_nodes(t), out)
        self.lib.orc_kd_free(t)
        return r
