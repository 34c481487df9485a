"""Python mirror of the codec layer around the hot path
(/root/reference/include/Compressor.hpp:9-48, src/Compressor.cpp, include/ColorSpace.hpp).

Only ``compress`` touches the GPU (through the C ABI); block extraction, byte conversion, decode
and the ``.quant`` container are restated here as host logic with the reference's exact layout
rules so that the parity tests read like the reference's own test (src/test.cpp).
"""
from __future__ import annotations

import enum
import time
from dataclasses import dataclass
from typing import Tuple

import numpy as np

from ._lib import CS_CIE1931, CS_NORMAL, CS_SCALED
from .context import codebook_to_bytes
from .quantizer import Quantizers, getQuantizer
from .rgbimage import RGBImage


class ColorSpaces(enum.IntEnum):
    NORMAL = 0
    SCALED = 1
    CIE1931 = 2


def _check_cs(cs) -> int:
    cs = int(cs)
    if cs not in (CS_NORMAL, CS_SCALED, CS_CIE1931):
        raise ValueError(f"unknown colour space {cs}")
    return cs


def pack_indices(indices, bits: int) -> np.ndarray:
    """Index i -> stream bits [i*bits, (i+1)*bits), bit p of the stream = bit p%8 of byte p/8 (LSB first)."""
    a = np.asarray(indices, np.uint64)
    b = ((a[:, None] >> np.arange(bits, dtype=np.uint64)[None, :]) & 1).astype(np.uint8).reshape(-1)
    return np.packbits(b, bitorder="little")


def unpack_indices(stream, n: int, bits: int) -> np.ndarray:
    b = np.unpackbits(np.asarray(stream, np.uint8), bitorder="little")[:n * bits].reshape(n, bits).astype(np.uint64)
    return (b << np.arange(bits, dtype=np.uint64)[None, :]).sum(1).astype(np.uint64)

MAX_CODE_LEN = 32


def huffman_lengths(counts) -> np.ndarray:
    """Code lengths (0 for unused symbols, at most MAX_CODE_LEN) of a Huffman code for `counts`."""
    import heapq
    counts = np.asarray(counts, np.uint64)
    length = np.zeros(counts.size, np.uint8)
    used = np.nonzero(counts)[0]
    if used.size == 0:
        return length
    if used.size == 1:
        length[used[0]] = 1
        return length
    heap = [(int(counts[k]), i, None) for i, k in enumerate(used)]   # (weight, tie-break, children)
    heapq.heapify(heap)
    parent, nxt = {}, used.size
    while len(heap) > 1:
        a = heapq.heappop(heap)
        b = heapq.heappop(heap)
        parent[a[1]] = parent[b[1]] = nxt
        heapq.heappush(heap, (a[0] + b[0], nxt, None))
        nxt += 1
    depth = {}
    for node in range(nxt - 1, -1, -1):          # parents carry larger numbers than their children
        depth[node] = 0 if node not in parent else depth[parent[node]] + 1
    for i, k in enumerate(used):
        length[k] = min(depth[i], MAX_CODE_LEN)
    kraft = sum(1 << (MAX_CODE_LEN - int(l)) for l in length[used])
    while kraft > (1 << MAX_CODE_LEN):           # lengthen the longest codes still short of the limit
        cand = used[length[used] < MAX_CODE_LEN]
        k = cand[np.argmax(length[cand])]
        kraft -= 1 << (MAX_CODE_LEN - int(length[k]) - 1)
        length[k] += 1
    return length


def canonical_codes(length):
    """(code per symbol, symbols sorted by (length, index), first_code / first_pos / count per length)."""
    length = np.asarray(length, np.int64)
    if length.size and length.max() > MAX_CODE_LEN:
        raise ValueError("code length out of range")
    n_of = np.bincount(length, minlength=MAX_CODE_LEN + 2)
    n_of[0] = 0
    first_code = np.zeros(MAX_CODE_LEN + 2, np.int64)
    first_pos = np.zeros(MAX_CODE_LEN + 2, np.int64)
    code = pos = 0
    for l in range(1, MAX_CODE_LEN + 1):
        code <<= 1
        first_code[l], first_pos[l] = code, pos
        code += int(n_of[l])
        pos += int(n_of[l])
        if code > (1 << l):
            raise ValueError("code lengths oversubscribed")
    used = np.nonzero(length)[0]
    order = used[np.lexsort((used, length[used]))]
    codes = np.zeros(length.size, np.int64)
    rank = np.arange(order.size) - first_pos[length[order]]
    codes[order] = first_code[length[order]] + rank
    return codes, order, first_code, first_pos, n_of


def huffman_encode(indices, length) -> np.ndarray:
    """The indices as canonical Huffman codes, MSB first, zero padded to a whole byte."""
    a = np.asarray(indices, np.int64)
    length = np.asarray(length, np.int64)
    codes, _, _, _, _ = canonical_codes(length)
    ls, cs = length[a], codes[a]
    if a.size and ls.min() == 0:
        raise ValueError("an index without a code")
    start = np.concatenate([[0], np.cumsum(ls)[:-1]]) if a.size else np.zeros(0, np.int64)
    total = int(ls.sum())
    bits = np.zeros(((total + 7) // 8) * 8, np.uint8)
    for b in range(int(ls.max()) if a.size else 0):
        m = ls > b
        bits[start[m] + b] = (cs[m] >> (ls[m] - 1 - b)) & 1
    return np.packbits(bits)


def huffman_decode(stream, n: int, length) -> np.ndarray:
    length = np.asarray(length, np.int64)
    _, order, first_code, first_pos, n_of = canonical_codes(length)
    bits = np.unpackbits(np.asarray(stream, np.uint8)).tolist()
    fc, fp, cnt, out, at = first_code.tolist(), first_pos.tolist(), n_of.tolist(), np.zeros(n, np.uint64), 0
    order = order.tolist()
    for i in range(n):
        code = l = 0
        while True:
            if at >= len(bits) or l >= MAX_CODE_LEN:
                raise ValueError("corrupt entropy-coded index stream")
            code = (code << 1) | bits[at]
            at += 1
            l += 1
            if cnt[l] and 0 <= code - fc[l] < cnt[l]:
                break
        out[i] = order[fp[l] + code - fc[l]]
    return out



def _block_index_map(xSize: int, ySize: int, w: int, h: int):
    """img index of every (vector, pixel-in-block) pair: (N, w*h) int64, per src/Compressor.cpp:44-50."""
    wB, hB = (xSize + w - 1) // w, (ySize + h - 1) // h
    i = np.arange(wB, dtype=np.int64)[:, None, None, None]
    j = np.arange(hB, dtype=np.int64)[None, :, None, None]
    dx = np.arange(w, dtype=np.int64)[None, None, :, None]
    dy = np.arange(h, dtype=np.int64)[None, None, None, :]
    idx = (i * w + dx) * ySize + (j * h + dy)
    return idx.reshape(wB * hB, w * h)


def getBlocksAsVectorsFromImage(image: RGBImage, w: int, h: int, cs) -> np.ndarray:
    """src/Compressor.cpp:31-62 -> (N, 3*w*h) float64."""
    cs = _check_cs(cs)
    idx = _block_index_map(image.xSize, image.ySize, w, h)
    npix = image.xSize * image.ySize
    valid = idx < npix
    px = image.img[np.where(valid, idx, 0)].astype(np.int8).astype(np.float64)  # signed char
    if cs == CS_SCALED:
        px = (px + 128.0) / 255
    elif cs == CS_CIE1931:   # Cie1931::RGBtoColorSpace (src/ColorSpace.cpp:35-39), sums left to right
        c0, c1, c2 = px[..., 0], px[..., 1], px[..., 2]
        px = np.stack([(c0 * 0.490 + c1 * 0.310 + c2 * 0.200) / 0.17697,
                       (c0 * 0.17697 + c1 * 0.81240 + c2 * 0.01063) / 0.17697,
                       (0.0 + c1 * 0.01 + c2 * 0.99) / 0.17697], -1)
    px = np.where(valid[..., None], px, 0.0)
    return px.reshape(idx.shape[0], -1)


def vectorsToCharVectorsColorSpaced(vectors: np.ndarray, cs) -> np.ndarray:
    """src/Compressor.cpp:12-29 -> (K, dim) uint8 (the reference's chars, reinterpreted)."""
    return codebook_to_bytes(vectors, _check_cs(cs))


def getImageFromVectors(blocks: np.ndarray, xSize: int, ySize: int, w: int, h: int) -> RGBImage:
    """src/Compressor.cpp:64-92; writes happen in (i, j, x, y) order, later writes win."""
    blocks = np.ascontiguousarray(blocks, np.uint8)
    idx = _block_index_map(xSize, ySize, w, h)
    npix = xSize * ySize
    out = np.zeros((npix, 3), np.uint8)
    flat_idx = idx.reshape(-1)
    vals = blocks.reshape(-1, 3)
    keep = flat_idx < npix
    # numpy fancy assignment applies duplicates in order, i.e. the last one wins, like the loops
    out[flat_idx[keep]] = vals[keep]
    return RGBImage.from_array(out, xSize, ySize)


@dataclass
class CompressionRaport:
    distortion: float
    bitsPerPixel: float
    uncompressedSize: int
    compressedSize: int
    compressionTime: float

    def __str__(self):  # src/Compressor.cpp:290-305
        return ("Compression raport: \n"
                f"Distortion        = {self.distortion:.10f}\n"
                f"Bits per pixel    = {self.bitsPerPixel:.10f}\n"
                f"Uncompressed size = {_pretty(self.uncompressedSize)}\n"
                f"Compressed size   = {_pretty(self.compressedSize)}\n"
                f"Compression ratio = {self.compressedSize / self.uncompressedSize:.3f}\n"
                f"Compression time  = {self.compressionTime:.3f}s\n")


def _pretty(b: int) -> str:  # src/Compressor.cpp:269-288 (including its odd Mb remainder)
    if b < 1024:
        return f"{b}b"
    if b < 1024 * 1024:
        return f"{b // 1024},{b % 1024}Kb"
    return f"{b // (1024 * 1024)},{b % (1024 * 1024)}Mb"


def _smallest_pow2(n: int) -> int:  # src/Compressor.cpp:167-172
    p = 0
    while n // 2:
        n //= 2
        p += 1
    return p


class CompressedImage:
    def __init__(self):
        self.codeVectors = np.zeros((0, 0), np.uint8)       # (K, dim) bytes
        self.assignedCodeVector = np.zeros(0, np.uint64)     # N indices (size_t)
        self.xSize = self.ySize = 0
        self.blockWidth = self.blockHeight = 0
        self.colorSpace = ColorSpaces.SCALED
        self.quantizer = Quantizers.LBG

    # -- the hot path --------------------------------------------------------------------------
    @staticmethod
    def compress(image: RGBImage, quantizer, colorSpace, blockWidth: int, blockHeight: int,
                 eps: float, N: int, device: int = 0, context=None
                 ) -> Tuple["CompressedImage", CompressionRaport]:
        """src/Compressor.cpp:107-154.  The timed region is the reference's: block extraction +
        quantize, here = host bytes -> GPU -> codebook and indices back on the host."""
        cs = _check_cs(colorSpace)
        q = getQuantizer(quantizer, device, context)
        if q is None:
            raise ValueError("quantizer not implemented (the reference returns nullptr and crashes)")
        t0 = time.perf_counter()
        codeVectors, assigned, _ = q.quantize_image(image.img, image.xSize, image.ySize, blockWidth,
                                                    blockHeight, cs, N, eps)
        dt = time.perf_counter() - t0
        res = CompressedImage()
        res.codeVectors = vectorsToCharVectorsColorSpaced(codeVectors, cs)
        res.assignedCodeVector = assigned
        res.xSize, res.ySize = image.xSize, image.ySize
        res.blockWidth, res.blockHeight = blockWidth, blockHeight
        res.colorSpace = ColorSpaces(cs)  # the reference leaves this member uninitialised
        res.quantizer = Quantizers(int(quantizer))
        res.last_reports = q.last_reports
        bpp = np.float32(res.sizeInBits()) / np.float32(image.xSize * image.ySize)
        # report distortion: decode on the GPU and compare bytes as signed chars (:137-146)
        _, mse = q.context.decode(res.codeVectors, want_image=False)
        rap = CompressionRaport(mse, float(bpp), image.sizeInBytes(), res.sizeInBits() // 8, dt)
        return res, rap

    @staticmethod
    def decompress(cImg: "CompressedImage") -> RGBImage:
        """src/Compressor.cpp:156-165."""
        blocks = cImg.codeVectors[np.asarray(cImg.assignedCodeVector, np.int64)]
        return getImageFromVectors(blocks, cImg.xSize, cImg.ySize, cImg.blockWidth, cImg.blockHeight)

    def sizeInBits(self) -> int:
        """src/Compressor.cpp:174-182 (approximate, bit-packed size)."""
        bits = (_smallest_pow2(len(self.codeVectors)) * len(self.assignedCodeVector)
                + self.blockWidth * self.blockHeight * len(self.codeVectors) * 8 * 3)
        return ((bits + 7) // 8) * 8

    # -- .quant container (src/Compressor.cpp:190-267) ------------------------------------------
    def to_bytes(self) -> bytes:
        bits = _smallest_pow2(len(self.codeVectors))
        assert bits <= 24
        hdr = b"%d %d %d %d %d %d %d\n" % (bits, int(self.colorSpace), len(self.assignedCodeVector),
                                          self.xSize, self.ySize, self.blockWidth, self.blockHeight)
        bpi = (bits + 7) // 8
        a = np.asarray(self.assignedCodeVector, "<u8")
        idx = a.view(np.uint8).reshape(-1, 8)[:, :bpi]  # low bytes of a little-endian size_t
        return hdr + np.ascontiguousarray(self.codeVectors, np.uint8).tobytes() + idx.tobytes()

    def to_bytes_packed(self) -> bytes:
        """Extension (README "possible improvements"; `quant --pack`): header "QP1 ..." and the indices as one
        LSB-first bit stream of `bits` bits each - the size sizeInBits() has always reported."""
        bits = _smallest_pow2(len(self.codeVectors))
        hdr = b"QP1 %d %d %d %d %d %d %d\n" % (bits, int(self.colorSpace), len(self.assignedCodeVector),
                                               self.xSize, self.ySize, self.blockWidth, self.blockHeight)
        return hdr + np.ascontiguousarray(self.codeVectors, np.uint8).tobytes() + pack_indices(self.assignedCodeVector, bits).tobytes()

    def to_bytes_entropy(self) -> bytes:
        """Extension (`quant --entropy`): header "QH1 ... <stream bytes>", codebook, one code length per codevector, and
        the indices as canonical Huffman codes of their own histogram (quant_b200/host/src/Compressor.cpp)."""
        K = len(self.codeVectors)
        bits = _smallest_pow2(K)
        a = np.asarray(self.assignedCodeVector, np.int64)
        length = huffman_lengths(np.bincount(a, minlength=K))
        stream = huffman_encode(a, length)
        hdr = b"QH1 %d %d %d %d %d %d %d %d\n" % (bits, int(self.colorSpace), a.size, self.xSize, self.ySize,
                                                  self.blockWidth, self.blockHeight, stream.size)
        return hdr + np.ascontiguousarray(self.codeVectors, np.uint8).tobytes() + length.tobytes() + stream.tobytes()

    def saveToFile(self, path: str):
        with open(path, "wb") as f:
            f.write(self.to_bytes())

    def saveToFileEntropy(self, path: str):
        with open(path, "wb") as f:
            f.write(self.to_bytes_entropy())

    def saveToFilePacked(self, path: str):
        with open(path, "wb") as f:
            f.write(self.to_bytes_packed())

    def loadFromFile(self, path: str):
        with open(path, "rb") as f:
            data = f.read()
        nl = data.index(b"\n")
        fields = data[:nl].split()
        packed, entropy = fields[0] == b"QP1", fields[0] == b"QH1"
        if packed or entropy:
            fields = fields[1:]
        bits, cs, n, xs, ys, bw, bh = (int(t) for t in fields[:7])
        self.xSize, self.ySize, self.blockWidth, self.blockHeight = xs, ys, bw, bh
        try:
            self.colorSpace = ColorSpaces(cs)
        except ValueError:  # files written by the reference carry an uninitialised value here
            self.colorSpace = cs
        dim, K = bw * bh * 3, 1 << bits
        pos = nl + 1
        self.codeVectors = np.frombuffer(data, np.uint8, K * dim, pos).reshape(K, dim).copy()
        pos += K * dim
        if entropy:
            length = np.frombuffer(data, np.uint8, K, pos)
            stream = np.frombuffer(data, np.uint8, int(fields[7]), pos + K)
            self.assignedCodeVector = huffman_decode(stream, n, length)
            return
        if packed:
            stream = np.frombuffer(data, np.uint8, (n * bits + 7) // 8, pos)
            self.assignedCodeVector = unpack_indices(stream, n, bits)
            return
        bpi = (bits + 7) // 8
        raw = np.frombuffer(data, np.uint8, n * bpi, pos).reshape(n, bpi)
        a = np.zeros((n, 8), np.uint8)
        a[:, :bpi] = raw
        self.assignedCodeVector = a.view("<u8").reshape(-1).astype(np.uint64)
