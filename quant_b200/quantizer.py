"""Python mirror of the reference's quantiser plugin interface
(/root/reference/include/Quantizer.hpp:8-20), backed by libqb200 - no CPU path.

    q = getQuantizer(Quantizers.LBG)
    codeVectors, assignedCodeVector, distortion = q.quantize(trainingSet, n, eps)

``trainingSet`` is an (N, dim) float64 array: what getBlocksAsVectorsFromImage returns.  The fast GPU
path works on the integer lattice the image bytes live on, so the generic entry first tries to map the
doubles back to bytes (exactly); inputs that are not on a supported lattice take the library's general
FP64-vector path (qb200_set_vectors_f64).  ``CompressedImage.compress`` skips this
and hands the raw image bytes to the library.
"""
from __future__ import annotations

import enum
from typing import Optional, Tuple

import numpy as np

from ._lib import CS_NORMAL, CS_SCALED
from .context import Context


class Quantizers(enum.IntEnum):
    LBG = 0
    MEDIAN_CUT = 1
    LBG_MEDIAN_CUT = 2
    ABC = 3


def vectors_to_lattice_bytes(X: np.ndarray) -> Tuple[np.ndarray, int]:
    """(N, dim) float64 -> ((N, dim) uint8 image bytes, colour space) or ValueError.

    NORMAL lattice: x is an integer in [-128, 127]          (byte = x mod 256)
    SCALED lattice: x == (t/255.0) for an integer t in [0,255] (byte = t xor 0x80)
    Integer-valued inputs take the NORMAL lattice (their sums are exact in FP64, as in the
    reference); everything else must sit on the SCALED lattice bit-for-bit.
    """
    X = np.ascontiguousarray(X, np.float64)
    r = np.rint(X)
    if np.array_equal(r, X) and X.size and X.min() >= -128 and X.max() <= 127:
        return (r.astype(np.int64) & 0xFF).astype(np.uint8), CS_NORMAL
    t = np.rint(X * 255.0)
    ok = (t >= 0) & (t <= 255)
    if ok.all() and np.array_equal(t / 255.0, X):
        return (t.astype(np.int64) ^ 0x80).astype(np.uint8), CS_SCALED
    raise ValueError("trainingSet is not on the NORMAL or SCALED byte lattice")


class AbstractQuantizer:
    def quantize(self, trainingSet, n: int, eps: float):
        raise NotImplementedError


class LBGQuantizer(AbstractQuantizer):
    """LBG with codebook splitting, HEAD schedule (/root/reference/src/Quantizer.cpp:121-143)."""

    def __init__(self, device: int = 0, context: Optional[Context] = None):
        self._ctx = context
        self._device = device
        self.last_reports = None

    @property
    def context(self) -> Context:
        if self._ctx is None:
            self._ctx = Context(self._device)
        return self._ctx

    def quantize(self, trainingSet, n: int, eps: float):
        X = np.asarray(trainingSet, np.float64)
        if X.ndim != 2 or X.shape[0] == 0:
            # Solution's constructor calls trainingSet.at(0) (src/Quantizer.cpp:91)
            raise IndexError("trainingSet is empty")
        ctx = self.context
        try:
            mat, cs = vectors_to_lattice_bytes(X)
            ctx.set_vectors_u8(mat, cs)
        except ValueError:          # arbitrary doubles: the general FP64-vector path of the library
            if X.shape[1] > 192:
                raise
            ctx.set_vectors_f64(X)
        cb, dist, self.last_reports = ctx.train(int(n), float(eps))
        return cb, ctx.get_assign_u64(), dist

    def quantize_image(self, rgb, xSize, ySize, w, h, colorspace, n, eps):
        """Fast path used by CompressedImage.compress: raw bytes in, no N x dim doubles."""
        ctx = self.context
        ctx.set_image(rgb, xSize, ySize, w, h, colorspace)
        cb, dist, self.last_reports = ctx.train(int(n), float(eps))
        return cb, ctx.get_assign_u64(), dist


def getQuantizer(q, device: int = 0, context: Optional[Context] = None):
    """src/Quantizer.cpp:146-155: only LBG exists; every other enum value yields None (nullptr)."""
    if int(q) == int(Quantizers.LBG):
        return LBGQuantizer(device, context)
    return None
