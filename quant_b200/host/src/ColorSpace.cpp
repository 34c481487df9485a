// NORMAL, SCALED and CIE1931 colour maps (/root/reference/src/ColorSpace.cpp:4-62).
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
// CIE1931: XYZ = M * rgb / 0.17697 on the SIGNED bytes, back through the reference's rounded inverse matrix.
// The values are not on a byte lattice: the library trains on them as FP64 vectors (qb200_set_vectors_f64 /
// QB200_CS_CIE1931).  Every row is summed left to right, as the reference writes it; the build forbids contraction.
class Cie1931Color : public ColorSpace {
 public:
  RGBDouble RGBtoColorSpace(const RGB &c) override {
    static const double M[3][3] = {{0.490, 0.310, 0.200}, {0.17697, 0.81240, 0.01063}, {0.0, 0.01, 0.99}};
    const double p[3] = {(double)c[0], (double)c[1], (double)c[2]};
    RGBDouble r;
    for (int i = 0; i < 3; i++) {
      const double first = i == 2 ? 0.0 : p[0] * M[i][0];  // the reference's third row starts with the integer 0
      r[i] = (first + p[1] * M[i][1] + p[2] * M[i][2]) / 0.17697;
    }
    return r;
  }
  RGB colorSpaceToRGB(const RGBDouble &c) override {
    static const double W[3][3] = {{0.418, -0.15866, -0.082835}, {-0.091169, 0.25243, 0.015708}, {0.0009209, -0.0025498, 0.17860}};
    RGB r;
    for (int i = 0; i < 3; i++) r[i] = (char)(long long)std::round(c[0] * W[i][0] + c[1] * W[i][1] + c[2] * W[i][2]);
    return r;
  }
};
}  // namespace

ColorSpacePtr getColorSpace(ColorSpaces cs) {
  switch (cs) {
    case ColorSpaces::NORMAL: return ColorSpacePtr(new ColorSpace());
    case ColorSpaces::SCALED: return ColorSpacePtr(new ScaledColor());
    case ColorSpaces::CIE1931: return ColorSpacePtr(new Cie1931Color());
  }
  return nullptr;
}
