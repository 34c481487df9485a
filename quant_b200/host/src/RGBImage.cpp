// P6 reader / writer (/root/reference/src/RGBImage.cpp:6-33).  The reference validates with
// assert() only (compiled out in Release); here malformed input throws.
#include "RGBImage.hpp"

#include <cstring>
#include <fstream>
#include <stdexcept>

RGBImage::RGBImage(const std::string &path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("cannot open " + path);
  std::string magic;
  int maxval = 0;
  in >> magic >> xSize >> ySize >> maxval;
  if (!in || magic != "P6") throw std::runtime_error(path + ": not a binary PPM (P6)");
  if (maxval != MAX_COL - 1) throw std::runtime_error(path + ": maxval must be 255");
  if (xSize <= 0 || ySize <= 0) throw std::runtime_error(path + ": bad dimensions");
  in.get();  // the single whitespace byte that ends the header
  img.assign((size_t)xSize * ySize, RGB{{0, 0, 0}});
  in.read(reinterpret_cast<char *>(img.data()), (std::streamsize)img.size() * 3);
}

void RGBImage::saveToFile(const std::string &path) {
  std::ofstream out(path, std::ios::binary);
  if (!out) throw std::runtime_error("cannot write " + path);
  out << "P6\n" << xSize << " " << ySize << "\n" << MAX_COL - 1 << "\n";
  out.write(reinterpret_cast<const char *>(img.data()), (std::streamsize)img.size() * 3);
}

size_t RGBImage::sizeInBytes() const { return img.size() * 3; }
