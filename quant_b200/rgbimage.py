"""RGBImage: binary PPM (P6) container, mirroring /root/reference/include/RGBImage.hpp:15-23.

``img`` is an (xSize*ySize, 3) uint8 array in file order.  Note the reference's quirk that the
rest of the path depends on: ``xSize`` is the PPM *width*, yet pixels are later
addressed as ``x*ySize + y``.
"""
from __future__ import annotations

import numpy as np

MAX_COL = 256


class RGBImage:
    def __init__(self, path: str | None = None):
        self.xSize = 0
        self.ySize = 0
        self.img = np.zeros((0, 3), np.uint8)
        if path is not None:
            self._load(path)

    @classmethod
    def from_array(cls, rgb: np.ndarray, xSize: int, ySize: int) -> "RGBImage":
        im = cls()
        im.xSize, im.ySize = int(xSize), int(ySize)
        im.img = np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
        if im.img.shape[0] != im.xSize * im.ySize:
            raise ValueError("pixel count does not match xSize*ySize")
        return im

    def _load(self, path: str):
        # src/RGBImage.cpp:6-24: "P6", x, y, maxval (255), ONE whitespace byte, then raw bytes
        with open(path, "rb") as f:
            data = f.read()
        toks, pos = [], 0
        while len(toks) < 4:
            while data[pos:pos + 1].isspace():
                pos += 1
            start = pos
            while not data[pos:pos + 1].isspace():
                pos += 1
            toks.append(data[start:pos])
        pos += 1
        if toks[0] != b"P6":
            raise ValueError("not a binary PPM (P6)")
        self.xSize, self.ySize, maxc = int(toks[1]), int(toks[2]), int(toks[3])
        if maxc != MAX_COL - 1:
            raise ValueError("PPM maxval must be 255")
        n = self.xSize * self.ySize * 3
        buf = np.frombuffer(data, np.uint8, count=min(n, len(data) - pos), offset=pos)
        img = np.zeros(n, np.uint8)
        img[: buf.size] = buf
        self.img = img.reshape(-1, 3)

    def saveToFile(self, path: str):
        # src/RGBImage.cpp:26-33
        with open(path, "wb") as f:
            f.write(b"P6\n%d %d\n%d\n" % (self.xSize, self.ySize, MAX_COL - 1))
            f.write(self.img.tobytes())

    def sizeInBytes(self) -> int:
        return int(self.img.shape[0]) * 3

    def __eq__(self, o):
        return (isinstance(o, RGBImage) and self.xSize == o.xSize and self.ySize == o.ySize
                and np.array_equal(self.img, o.img))
