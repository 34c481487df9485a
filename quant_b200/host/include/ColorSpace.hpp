// Colour spaces of the codec (/root/reference/include/ColorSpace.hpp:6-17).
//   NORMAL : value = (double)(signed char)byte              src/ColorSpace.cpp:4-11
//   SCALED : value = ((signed char)byte + 128.0) / 255      src/ColorSpace.cpp:16-28   (CLI default)
//   CIE1931: a 3x3 matrix on top of SCALED - NOT on the B200 path (non-lattice inputs); asking the
//            factory for it throws, compress() rejects it (out of scope).
#pragma once
#include <memory>

#include "RGBImage.hpp"

enum class ColorSpaces { NORMAL, SCALED, CIE1931 };

class ColorSpace {
 public:
  virtual RGBDouble RGBtoColorSpace(const RGB &);
  virtual RGB colorSpaceToRGB(const RGBDouble &);
  virtual ~ColorSpace() = default;
};

typedef std::unique_ptr<ColorSpace> ColorSpacePtr;

ColorSpacePtr getColorSpace(ColorSpaces);
