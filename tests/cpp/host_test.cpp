// C++ checks of the drop-in host layer, written like the reference's only unit test
// (/root/reference/src/test.cpp:5-62, compressor_test.something) plus the plugin entry on a GPU.
//   host_test layout                CPU only: block -> vector -> char vector -> image round trips
//   host_test quantize <in.ppm> w h n [cs]  needs a B200: getQuantizer(LBG)->quantize(vectors) must agree with
//                                   CompressedImage::compress(image) (codebook bytes, indices)
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "Compressor.hpp"
#include "KDTree.hpp"

static int failures = 0;
#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);           \
      failures++;                                                           \
    }                                                                       \
  } while (0)

static int test_layout() {
  // 4x4 image of letters, ColorSpaces::NORMAL; 1x3 exercises the y-overflow wrap
  RGBImage img;
  img.xSize = 4;
  img.ySize = 4;
  for (int p = 0; p < 16; p++) img.img.push_back(RGB{{(char)('a' + p), (char)('A' + p), (char)('0' + p % 10)}});
  const int shapes[][2] = {{1, 1}, {2, 2}, {1, 3}, {2, 4}, {3, 3}, {4, 1}};
  for (auto &s : shapes) {
    for (ColorSpaces c : {ColorSpaces::NORMAL, ColorSpaces::SCALED}) {
      ColorSpacePtr cs = getColorSpace(c);
      auto vecs = getBlocksAsVectorsFromImage(img, s[0], s[1], cs);
      EXPECT(vecs.size() == (size_t)((4 + s[0] - 1) / s[0]) * ((4 + s[1] - 1) / s[1]));
      EXPECT(vecs[0].size() == (size_t)3 * s[0] * s[1]);
      auto chars = vectorsToCharVectorsColorSpaced(vecs, cs);
      RGBImage back = getImageFromVectors(chars, 4, 4, s[0], s[1]);
      EXPECT(back.img == img.img);
    }
  }
  // .quant round trip through a file
  CompressedImage c;
  c.xSize = 4; c.ySize = 4; c.blockWidth = 2; c.blockHeight = 2;
  c.codeVectors.assign(4, CharVector(12));
  for (int k = 0; k < 4; k++) for (int d = 0; d < 12; d++) c.codeVectors[k][d] = (char)(k * 40 - 100 + d);
  c.assignedCodeVector = {3, 0, 2, 1};
  const std::string path = "/tmp/qb200_host_test.quant";
  c.saveToFile(path);
  CompressedImage d;
  d.loadFromFile(path);
  EXPECT(d.codeVectors == c.codeVectors && d.assignedCodeVector == c.assignedCodeVector);
  EXPECT(d.xSize == 4 && d.ySize == 4 && d.blockWidth == 2 && d.blockHeight == 2);
  EXPECT(CompressedImage::decompress(d).img == CompressedImage::decompress(c).img);
  EXPECT(c.sizeInBits() == ((2 * 4 + 2 * 2 * 4 * 8 * 3 + 7) / 8) * 8);
  EXPECT(getQuantizer(Quantizers::MEDIAN_CUT) == nullptr && getQuantizer(Quantizers::ABC) == nullptr);
  EXPECT(norm(Vector{1.0, 2.0, 2.0}) == 9.0);
  EXPECT((Vector{1.0, 2.0} * 1.2)[1] == 2.0 * 1.2);
  return failures;
}

static int test_quantize(const std::string &ppm, int w, int h, int n, ColorSpaces space) {
  RGBImage img(ppm);
  auto res = CompressedImage::compress(img, Quantizers::LBG, space, w, h, 1e-6f, n);
  ColorSpacePtr cs = getColorSpace(space);  // CIE1931: the vectors below are not on a byte lattice (FP64 path)
  auto vecs = getBlocksAsVectorsFromImage(img, w, h, cs);
  auto q = getQuantizer(Quantizers::LBG);
  auto out = q->quantize(vecs, n, 1e-6f);
  auto chars = vectorsToCharVectorsColorSpaced(std::get<0>(out), cs);
  EXPECT(chars == res.first.codeVectors);
  EXPECT(std::get<1>(out) == res.first.assignedCodeVector);
  EXPECT(std::get<2>(out) >= 0);
  // the public nearest-neighbour class against the returned codebook: one more LBG assignment step
  {
    KDTree tree(vecs[0].size(), std::get<0>(out));
    std::vector<size_t> nn = tree.nearestNeighbours(vecs);
    EXPECT(nn.size() == vecs.size());
    size_t worse = 0;
    for (size_t i = 0; i < vecs.size(); i += 97) {
      const double d_nn = norm(vecs[i] - std::get<0>(out)[nn[i]]);
      for (size_t k = 0; k < std::get<0>(out).size(); k += 7) worse += norm(vecs[i] - std::get<0>(out)[k]) < d_nn;
      EXPECT(tree.nearestNeighbour(vecs[i]) == nn[i]);
      if (i > 97 * 20) break;
    }
    EXPECT(worse == 0);
  }
  bool threw = false;
  try {
    q->quantize(std::vector<Vector>(), n, 1e-6f);
  } catch (const std::out_of_range &) {
    threw = true;
  }
  EXPECT(threw);
  std::cout << res.second;
  return failures;
}

int main(int argc, char **argv) {
  try {
    if (argc >= 2 && std::strcmp(argv[1], "layout") == 0) return test_layout() ? 1 : (std::puts("layout ok"), 0);
    if (argc >= 6 && std::strcmp(argv[1], "quantize") == 0)
      return test_quantize(argv[2], std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]),
                           argc >= 7 ? (ColorSpaces)std::atoi(argv[6]) : ColorSpaces::SCALED) ? 1
                                                                                              : (std::puts("quantize ok"), 0);
  } catch (const std::exception &e) {
    std::printf("exception: %s\n", e.what());
    return 3;
  }
  std::puts("usage: host_test layout | quantize <in.ppm> w h n [colourspace]");
  return 2;
}
