"""quant_b200 - B200-native LBG codebook training / index assignment behind the reference's
Quantizer / Compressor interface.  The compute path is libqb200.so (hand-written sm_100a CUDA,
C ABI in include/qb200.h); this package is the host-side mirror of the reference's API.
"""
from ._lib import (CS_CIE1931, CS_NORMAL, CS_SCALED, LIB_PATH, MODE_FULL, MODE_FULL_REPAIR, MODE_PARITY, Qb200Error, load)
from . import distributed
from .context import Context, codebook_to_bytes, finalize_level, launch_count
from .rgbimage import RGBImage
from .quantizer import (AbstractQuantizer, LBGQuantizer, Quantizers, getQuantizer,
                        vectors_to_lattice_bytes)
from .compressor import (ColorSpaces, CompressedImage, CompressionRaport,
                         getBlocksAsVectorsFromImage, getImageFromVectors,
                         pack_indices, unpack_indices, huffman_lengths, huffman_encode, huffman_decode, vectorsToCharVectorsColorSpaced)

__all__ = [n for n in dir() if not n.startswith("_")]
