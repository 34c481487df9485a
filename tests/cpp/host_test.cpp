// C++ checks of the drop-in host layer, written like the reference's only unit test
// (/root/reference/src/test.cpp:5-62, compressor_test.something) plus the plugin entry on a GPU.
//   host_test layout                CPU only: block -> vector -> char vector -> image round trips
//   host_test entropy <tmp.quant>   CPU only: the Huffman-coded container on an index histogram whose optimal code is
//                                   deeper than the 32-bit limit (Fibonacci counts), on a one-symbol and an empty image
//   host_test quantize <in.ppm> w h n [cs]  needs a B200: getQuantizer(LBG)->quantize(vectors) must agree with
//                                   CompressedImage::compress(image) (codebook bytes, indices)
//   host_test multi xs ys w h n ndev [exact]  needs ndev B200s: one image trained on ONE device and on a
//                                   qb200_create_multi context over ndev devices - codebook bits, indices and
//                                   distortion must be identical (ndev = 0: every visible device)
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include <chrono>
#include <vector>

#include "../../include/qb200.h"
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

// saveToFileEntropy / loadFromFile round trips where the code-length limit matters.
static int test_entropy(const char *path) {
  auto roundtrip = [&](CompressedImage &c) {
    c.saveToFileEntropy(path);
    CompressedImage b;
    b.loadFromFile(path);
    EXPECT(b.assignedCodeVector == c.assignedCodeVector);
    EXPECT(b.codeVectors == c.codeVectors);
    EXPECT(b.xSize == c.xSize && b.ySize == c.ySize && b.blockWidth == c.blockWidth && b.blockHeight == c.blockHeight);
  };
  CompressedImage c;
  c.blockWidth = c.blockHeight = 1;
  c.colorSpace = ColorSpaces::SCALED;
  c.quantizer = Quantizers::LBG;
  c.codeVectors.assign(64, CharVector(3));
  for (size_t k = 0; k < 64; k++) c.codeVectors[k][0] = (char)k;
  // Fibonacci counts for 35 of the 64 codevectors: the unrestricted Huffman tree is a 34-deep comb
  size_t fa = 1, fb = 1;
  for (size_t k = 0; k < 35; k++) {
    c.assignedCodeVector.insert(c.assignedCodeVector.end(), fa, k + 3);
    const size_t next = fa + fb;
    fa = fb;
    fb = next;
  }
  c.xSize = c.assignedCodeVector.size();
  c.ySize = 1;
  // interleave a little so that long and short codes alternate in the stream
  for (size_t i = 0; i + 1 < c.assignedCodeVector.size(); i += 97) std::swap(c.assignedCodeVector[i], c.assignedCodeVector[c.assignedCodeVector.size() - 1 - i]);
  roundtrip(c);
  CompressedImage one = c;   // a single used codevector: one-bit codes
  one.assignedCodeVector.assign(1000, 17);
  one.xSize = 1000;
  roundtrip(one);
  CompressedImage none = c;  // no blocks at all
  none.assignedCodeVector.clear();
  none.xSize = 0;
  roundtrip(none);
  return failures;
}

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

static int test_multi(int xs, int ys, int w, int h, int n, int ndev, int exact) {
  std::vector<uint8_t> rgb((size_t)xs * ys * 3);
  uint64_t st = 0x9E3779B97F4A7C15ull;
  for (auto &b : rgb) {  // xorshift noise with a flat patch of duplicates
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    b = (uint8_t)(st >> 24);
  }
  for (size_t i = 0; i < rgb.size() / 50; i++) rgb[i] = 200;
  const size_t dim = (size_t)3 * w * h, K = (size_t)1 << n;
  auto run = [&](qb200_ctx *ctx, std::vector<double> &cb, std::vector<uint64_t> &a, double &dist, double &ms) -> int {
    if (qb200_set_exact_centroids(ctx, exact)) return 1;
    cb.resize(K * dim);
    for (int rep = 0; rep < 2; rep++) {  // second run timed: allocations and module loading are behind it
      const auto t0 = std::chrono::steady_clock::now();
      int rc = qb200_set_image(ctx, rgb.data(), xs, ys, w, h, QB200_CS_SCALED, 1, 0);
      if (!rc) rc = qb200_train(ctx, n, 1e-6, QB200_MODE_PARITY, 0, nullptr, nullptr, cb.data(), &dist, nullptr);
      a.resize(qb200_num_vectors(ctx));
      if (!rc) rc = qb200_get_assign_u64(ctx, a.data());
      if (rc) {
        std::printf("libqb200 error %d: %s\n", rc, qb200_last_error(ctx));
        return 1;
      }
      ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return 0;
  };
  qb200_ctx *one = nullptr, *many = nullptr;
  if (qb200_create(0, &one) || qb200_create_multi(ndev, nullptr, &many)) {
    std::printf("context creation failed: %s\n", qb200_last_error(nullptr));
    return 1;
  }
  std::vector<double> cb1, cbm;
  std::vector<uint64_t> a1, am;
  double d1 = 0, dm = 0, ms1 = 0, msm = 0;
  if (run(one, cb1, a1, d1, ms1) || run(many, cbm, am, dm, msm)) return 1;
  EXPECT(cb1.size() == cbm.size() && std::memcmp(cb1.data(), cbm.data(), cb1.size() * 8) == 0);
  EXPECT(a1 == am);
  EXPECT(std::memcmp(&d1, &dm, 8) == 0);
  std::vector<uint8_t> cbb(K * dim);
  double mse1 = -1, msem = -2;
  qb200_codebook_to_bytes(cb1.data(), K, (int)dim, QB200_CS_SCALED, cbb.data());
  EXPECT(qb200_decode(one, cbb.data(), (uint32_t)K, nullptr, &mse1) == 0);
  EXPECT(qb200_decode(many, cbb.data(), (uint32_t)K, nullptr, &msem) == 0);
  EXPECT(mse1 == msem);
  std::printf("multi: %dx%d %dx%d K=%zu%s: 1 device %.2f ms, %d devices %.2f ms per host-bytes-in -> indices-out train; distortion %.9g\n",
              xs, ys, w, h, K, exact ? " (exact centroids)" : "", ms1, ndev, msm, d1);
  qb200_destroy(one);
  qb200_destroy(many);
  return failures;
}

int main(int argc, char **argv) {
  try {
    if (argc >= 8 && std::strcmp(argv[1], "multi") == 0)
      return test_multi(std::atoi(argv[2]), std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]), std::atoi(argv[6]),
                        std::atoi(argv[7]), argc >= 9 ? std::atoi(argv[8]) : 0) ? 1 : (std::puts("multi ok"), 0);
    if (argc >= 2 && std::strcmp(argv[1], "layout") == 0) return test_layout() ? 1 : (std::puts("layout ok"), 0);
    if (argc >= 3 && std::strcmp(argv[1], "entropy") == 0) return test_entropy(argv[2]) ? 1 : (std::puts("entropy ok"), 0);
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
