// Codec layer around the hot path, with the reference's public surface
// (/root/reference/include/Compressor.hpp:9-48).  compress() hands the raw image bytes to libqb200
// (no N x dim doubles are materialised); everything else is host code that keeps the reference's
// layout rules and the .quant container byte for byte.
#pragma once
#include <chrono>
#include <memory>
#include <ostream>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "ColorSpace.hpp"
#include "Quantizer.hpp"
#include "VectorOperations.hpp"

class CompressionRaport {
 public:
  VectorType distortion;  // pixel-domain mean squared error of decode(compress(image)), bytes as signed chars
  float bitsPerPixel;
  size_t uncompressedSize;
  size_t compressedSize;
  std::chrono::duration<double> compressionTime;  // host bytes in -> codebook + indices on the host
  friend std::ostream &operator<<(std::ostream &stream, const CompressionRaport &raport);
};

class CompressedImage {
 public:
  CompressedImage() = default;
  void saveToFile(const std::string &path);
  void loadFromFile(const std::string &path);
  size_t sizeInBits();

  static std::pair<CompressedImage, CompressionRaport> compress(const RGBImage &image, Quantizers quantizer,
                                                                ColorSpaces colorSpace, int blockWidth,
                                                                int blockHeight, VectorType eps, int N);
  static RGBImage decompress(const CompressedImage &);

  std::vector<CharVector> codeVectors;
  std::vector<size_t> assignedCodeVector;
  size_t xSize = 0, ySize = 0;
  size_t blockWidth = 0, blockHeight = 0;
  ColorSpaces colorSpace = ColorSpaces::SCALED;  // the reference leaves this member uninitialised in compress()
  Quantizers quantizer = Quantizers::LBG;
};

std::vector<CharVector> vectorsToCharVectorsColorSpaced(const std::vector<Vector> &vectors, const ColorSpacePtr &cs);
std::vector<Vector> getBlocksAsVectorsFromImage(const RGBImage &image, int w, int h, const ColorSpacePtr &);
RGBImage getImageFromVectors(const std::vector<CharVector> &blocks, int xSize, int ySize, int w, int h);
