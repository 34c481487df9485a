// NORMAL and SCALED colour maps (/root/reference/src/ColorSpace.cpp:4-28, 49-62).
#include "ColorSpace.hpp"

#include <cmath>
#include <stdexcept>

// NORMAL: the signed byte itself.
RGBDouble ColorSpace::RGBtoColorSpace(const RGB &c) {
  return RGBDouble{{(double)c[0], (double)c[1], (double)c[2]}};
}
RGB ColorSpace::colorSpaceToRGB(const RGBDouble &c) {
  return RGB{{(char)std::round(c[0]), (char)std::round(c[1]), (char)std::round(c[2])}};
}

namespace {
// SCALED: (signed byte + 128) / 255.  The inverse is the reference's literal expression
// (char)round((c - 128.0) * 255): it lands on the right byte only through 8-bit wrap-around
// (-32640 == 128 mod 256), which is kept on purpose - the codebook bytes of a .quant file depend on it.
class ScaledColor : public ColorSpace {
 public:
  RGBDouble RGBtoColorSpace(const RGB &c) override {
    RGBDouble r;
    for (int i = 0; i < 3; i++) r[i] = ((double)c[i] + 128.0) / 255;
    return r;
  }
  RGB colorSpaceToRGB(const RGBDouble &c) override {
    RGB r;
    for (int i = 0; i < 3; i++) r[i] = (char)(long long)std::round((c[i] - 128.0) * 255);
    return r;
  }
};
}  // namespace

ColorSpacePtr getColorSpace(ColorSpaces cs) {
  switch (cs) {
    case ColorSpaces::NORMAL: return ColorSpacePtr(new ColorSpace());
    case ColorSpaces::SCALED: return ColorSpacePtr(new ScaledColor());
    default: break;
  }
  throw std::runtime_error("colour space CIE1931 is outside the B200 path (byte-lattice inputs only)");
}
