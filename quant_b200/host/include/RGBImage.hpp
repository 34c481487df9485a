// Binary PPM (P6) image container with the reference's public surface
// (/root/reference/include/RGBImage.hpp:8-23).  Pixels are SIGNED chars, as in the reference: that signedness
// is what the colour spaces are defined on.
//
// Layout quirk the whole path depends on: xSize is the PPM *width*, the pixel buffer is in
// file (row-major) order, yet block extraction addresses pixel (x, y) as img[x * ySize + y].
#pragma once
#include <array>
#include <cstddef>
#include <string>
#include <vector>

// 8-bit channels: MAX_COL - 1 = 255 is the only maxval the reader accepts
const static int MAX_COL_BITS = 8;
const static int MAX_COL = 1 << MAX_COL_BITS;

typedef std::array<char, 3> RGB;          // one pixel as stored
typedef std::array<double, 3> RGBDouble;  // one pixel in a colour space

class RGBImage {
 public:
  int xSize = 0, ySize = 0;  // PPM width and height
  std::vector<RGB> img;      // xSize * ySize pixels in file order

  RGBImage() = default;
  explicit RGBImage(const std::string &path);  // reads a P6 file; throws std::runtime_error when malformed
                                               // (the reference only asserts, which Release builds compile out)
  void saveToFile(const std::string &path);    // writes "P6\n<x> <y>\n255\n" + the pixel bytes
  size_t sizeInBytes() const;                  // 3 * pixels: the "uncompressed size" of the report
};
