// Plain-C doorway into the C++ codec API (include/quantsrc_c.h): lets bench.py and ctypes tests drive
// CompressedImage::compress - the call the reference's main() makes (/root/reference/src/main.cpp:79-80) - with
// pageable host memory, std::vector<RGB> pixels and std::vector<size_t> indices.
#include "quantsrc_c.h"

#include <cstring>
#include <exception>
#include <string>

#include "Compressor.hpp"
#include "RGBImage.hpp"

namespace {
thread_local std::string g_error;
}

extern "C" {

int quantsrc_compress(const uint8_t *rgb, int xSize, int ySize, int colorspace, int blockWidth, int blockHeight,
                      double eps, int nbits, uint8_t *cb_bytes_out, uint64_t *assign_out, double *distortion_out,
                      float *bpp_out, double *seconds_out) {
  try {
    if (!rgb || xSize <= 0 || ySize <= 0) throw std::runtime_error("quantsrc_compress: empty image");
    RGBImage im;
    im.xSize = xSize;
    im.ySize = ySize;
    im.img.resize((size_t)xSize * ySize);
    std::memcpy(im.img.data(), rgb, im.img.size() * 3);
    auto res = CompressedImage::compress(im, Quantizers::LBG, (ColorSpaces)colorspace, blockWidth, blockHeight,
                                         (VectorType)eps, nbits);
    CompressedImage &ci = res.first;
    const size_t dim = (size_t)3 * blockWidth * blockHeight;
    if (cb_bytes_out)
      for (size_t k = 0; k < ci.codeVectors.size(); k++)
        for (size_t d = 0; d < dim; d++) cb_bytes_out[k * dim + d] = (uint8_t)ci.codeVectors[k][d];
    if (assign_out)
      for (size_t i = 0; i < ci.assignedCodeVector.size(); i++) assign_out[i] = ci.assignedCodeVector[i];
    if (distortion_out) *distortion_out = res.second.distortion;
    if (bpp_out) *bpp_out = res.second.bitsPerPixel;
    if (seconds_out) *seconds_out = res.second.compressionTime.count();
    return (int)ci.codeVectors.size();
  } catch (const std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

const char *quantsrc_last_error(void) { return g_error.c_str(); }

}  // extern "C"
