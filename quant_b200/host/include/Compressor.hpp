// Codec layer around the hot path, with the reference's public surface
// (/root/reference/include/Compressor.hpp:9-48).  compress() hands the raw image bytes to libqb200 (no N x dim
// doubles are materialised); everything else is host code that keeps the reference's layout rules and the
// .quant container byte for byte.
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

// What `quant -r` prints (operator<<, same text as the reference's).
class CompressionRaport {
 public:
  VectorType distortion;                           // pixel-domain MSE of decode(compress(image)), bytes compared as
                                                   // signed chars - computed on the GPU (qb200_decode)
  float bitsPerPixel;                              // sizeInBits() / pixels
  size_t uncompressedSize;                         // 3 bytes per pixel
  size_t compressedSize;                           // sizeInBits() / 8
  std::chrono::duration<double> compressionTime;  // host bytes in -> codebook + indices back on the host
  friend std::ostream &operator<<(std::ostream &stream, const CompressionRaport &raport);
};

class CompressedImage {
 public:
  // ---- state (public in the reference too: main() and the tests touch it directly) ----
  std::vector<CharVector> codeVectors;     // K codevectors as bytes, dim = 3 * blockWidth * blockHeight each
  std::vector<size_t> assignedCodeVector;  // one codebook index per block, block i*hBlocks + j
  size_t xSize = 0, ySize = 0;             // image size in pixels
  size_t blockWidth = 0, blockHeight = 0;  // block size in pixels
  ColorSpaces colorSpace = ColorSpaces::SCALED;  // set by compress(); the reference leaves it uninitialised there
  Quantizers quantizer = Quantizers::LBG;

  CompressedImage() = default;

  // ---- the hot path: block extraction + LBG training + index assignment on the B200 ----
  static std::pair<CompressedImage, CompressionRaport> compress(const RGBImage &image, Quantizers quantizer,
                                                                ColorSpaces colorSpace, int blockWidth,
                                                                int blockHeight, VectorType eps, int N);
  // ---- the rest: host code ----
  static RGBImage decompress(const CompressedImage &);
  void saveToFile(const std::string &path);    // .quant: ASCII header line, codebook bytes, byte-aligned indices
  void loadFromFile(const std::string &path);  // reads both containers
  // Extension (the reference's README lists it under "possible improvements"; `quant --pack`): the same container
  // with the indices bit-packed, ceil(N * bits / 8) bytes instead of N * ceil(bits / 8), and a header line that
  // starts with "QP1 " so that the two cannot be confused.  Index i occupies stream bits [i*bits, (i+1)*bits),
  // LSB first - exactly the size sizeInBits() has always reported.
  void saveToFilePacked(const std::string &path);
  // Extension (`quant --entropy`): header "QH1 ..." and the indices as canonical Huffman codes of their own histogram
  // (code lengths stored per codevector) - smaller than the packed container wherever some cells are much more
  // populated than others (flat areas of natural images); see Compressor.cpp for the layout.
  void saveToFileEntropy(const std::string &path);
  size_t sizeInBits();                         // bit-packed size estimate the report uses
};

// Free helpers of the reference's codec, same signatures (src/Compressor.cpp:12-92): block <-> vector layout
// and codebook doubles -> bytes.  Host restatements; the GPU path never materialises these vectors.
std::vector<Vector> getBlocksAsVectorsFromImage(const RGBImage &image, int w, int h, const ColorSpacePtr &);
std::vector<CharVector> vectorsToCharVectorsColorSpaced(const std::vector<Vector> &vectors, const ColorSpacePtr &cs);
RGBImage getImageFromVectors(const std::vector<CharVector> &blocks, int xSize, int ySize, int w, int h);
