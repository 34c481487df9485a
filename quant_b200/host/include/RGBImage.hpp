// Binary PPM (P6) image container with the reference's public surface
// (/root/reference/include/RGBImage.hpp:8-23).  Pixels are SIGNED chars, as in the reference:
// that signedness is what the colour spaces are defined on.
#pragma once
#include <array>
#include <cstddef>
#include <string>
#include <vector>

const static int MAX_COL_BITS = 8;
const static int MAX_COL = 1 << MAX_COL_BITS;

typedef std::array<char, 3> RGB;
typedef std::array<double, 3> RGBDouble;

class RGBImage {
 public:
  RGBImage() = default;
  explicit RGBImage(const std::string &path);  // throws std::runtime_error on a malformed file
  void saveToFile(const std::string &path);
  size_t sizeInBytes() const;
  std::vector<RGB> img;  // file order; pixel (x, y) is later addressed as img[x * ySize + y]
  int xSize = 0, ySize = 0;
};
