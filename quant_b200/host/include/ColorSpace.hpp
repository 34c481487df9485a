// Colour spaces of the codec (/root/reference/include/ColorSpace.hpp:6-17): a pixel (three signed bytes) is
// mapped to three doubles before quantisation and back afterwards.
//
//   ColorSpaces::NORMAL   value = (double)(signed char)byte                 src/ColorSpace.cpp:4-11
//   ColorSpaces::SCALED   value = ((signed char)byte + 128.0) / 255         src/ColorSpace.cpp:16-28  (CLI default)
//   ColorSpaces::CIE1931  a 3x3 matrix on top of SCALED - NOT on the B200 path: its values leave the byte
//                         lattice the integer statistics rely on, so getColorSpace() throws for it and
//                         compress() rejects it (out of scope)
//
// The base class IS the NORMAL colour space (as in the reference); SCALED derives from it in ColorSpace.cpp.
#pragma once
#include <memory>

#include "RGBImage.hpp"

enum class ColorSpaces { NORMAL, SCALED, CIE1931 };

class ColorSpace {
 public:
  virtual ~ColorSpace() = default;
  virtual RGBDouble RGBtoColorSpace(const RGB &pixel);   // forward map (identity on the signed bytes)
  virtual RGB colorSpaceToRGB(const RGBDouble &value);  // inverse map: (char)std::round(component)
};

typedef std::unique_ptr<ColorSpace> ColorSpacePtr;

// Factory: throws std::runtime_error for CIE1931 (the reference builds a matrix colour space there).
ColorSpacePtr getColorSpace(ColorSpaces which);
