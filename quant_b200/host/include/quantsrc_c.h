/* quantsrc_c.h - a plain-C doorway into libquantsrc.so's C++ codec API (CompressedImage::compress), for callers
 * that cannot link C++ (bench.py's second end-to-end leg, ctypes tests).  It is NOT the drop-in boundary of the hot
 * path - that is include/qb200.h; this only lets a C caller drive the reference-facing C++ call
 * (/root/reference/src/Compressor.cpp:107-154) with pageable host memory and size_t indices, exactly as
 * `quant in.ppm -o out.quant` does. */
#ifndef QUANTSRC_C_H
#define QUANTSRC_C_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* CompressedImage::compress on xSize*ySize RGB pixels (3 bytes each, file order).  Any output pointer may be NULL.
 *   cb_bytes_out   K*dim codebook bytes        assign_out   one index per block (size_t widened to uint64)
 *   distortion_out the report's pixel MSE      bpp_out      the report's bits per pixel
 *   seconds_out    the report's "Compression time": host bytes in -> codebook + indices back on the host
 * Returns the number of codevectors, or a negative value when the C++ layer threw (quantsrc_last_error). */
int quantsrc_compress(const uint8_t *rgb, int xSize, int ySize, int colorspace, int blockWidth, int blockHeight,
                      double eps, int nbits, uint8_t *cb_bytes_out, uint64_t *assign_out, double *distortion_out,
                      float *bpp_out, double *seconds_out);
const char *quantsrc_last_error(void);
#ifdef __cplusplus
}
#endif
#endif
