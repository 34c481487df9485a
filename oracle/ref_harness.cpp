// TEST INFRASTRUCTURE ONLY.  C-ABI harness around the UNMODIFIED reference sources
// (compiled where they lie under /root/reference by oracle/Makefile; outputs go to
// oracle/_ref/, which is git-ignored).  Nothing here re-implements the reference's
// arithmetic except `replay_levels`, which re-walks LBGQuantizer::quantize
// (/root/reference/src/Quantizer.cpp:121-143) with the reference's own public KDTree
// (/root/reference/include/KDTree.hpp:7-16) so that per-level state can be dumped; the replay
// is accepted only if its final codebook and assignment are bit-identical to what the
// reference's quantize() returns on the same input (checked on every call).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load the resulting library.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "Compressor.hpp"
#include "KDTree.hpp"
#include "Quantizer.hpp"
#include "RGBImage.hpp"
#include "ColorSpace.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

RGBImage make_image(const uint8_t *rgb, int xSize, int ySize) {
  RGBImage im;
  im.xSize = xSize;
  im.ySize = ySize;
  im.img.resize((size_t)xSize * ySize);
  if (!im.img.empty()) std::memcpy(im.img.data(), rgb, im.img.size() * 3);
  return im;
}

std::vector<Vector> to_vectors(const double *X, size_t N, int dim) {
  std::vector<Vector> v(N);
  for (size_t i = 0; i < N; i++) {
    Vector t(dim);
    for (int d = 0; d < dim; d++) t[d] = X[i * dim + d];
    v[i] = std::move(t);
  }
  return v;
}

}  // namespace

extern "C" {

// 1 when built with -ffast-math (the reference's Release flags), 0 for the strict-IEEE build.
int ref_is_fast_math(void) {
#ifdef __FAST_MATH__
  return 1;
#else
  return 0;
#endif
}

int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void ref_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// getBlocksAsVectorsFromImage (src/Compressor.cpp:31-62). out: N*dim doubles.
int ref_blocks(const uint8_t *rgb, int xSize, int ySize, int w, int h, int cs, double *out) {
  RGBImage im = make_image(rgb, xSize, ySize);
  ColorSpacePtr csp = getColorSpace((ColorSpaces)cs);
  std::vector<Vector> v = getBlocksAsVectorsFromImage(im, w, h, csp);
  size_t dim = 3 * (size_t)w * h;
  for (size_t i = 0; i < v.size(); i++)
    for (size_t d = 0; d < dim; d++) out[i * dim + d] = v[i][d];
  return (int)v.size();
}

// vectorsToCharVectorsColorSpaced (src/Compressor.cpp:12-29).
int ref_codebook_to_bytes(const double *cb, size_t K, int dim, int cs, uint8_t *out) {
  ColorSpacePtr csp = getColorSpace((ColorSpaces)cs);
  std::vector<CharVector> c = vectorsToCharVectorsColorSpaced(to_vectors(cb, K, dim), csp);
  for (size_t k = 0; k < K; k++)
    for (int d = 0; d < dim; d++) out[k * dim + d] = (uint8_t)c[k][d];
  return 0;
}

// getImageFromVectors(decompress) (src/Compressor.cpp:64-92,156-165).
int ref_decode(const uint8_t *cb_bytes, size_t K, const uint64_t *assign, size_t N, int xSize,
               int ySize, int w, int h, uint8_t *rgb_out) {
  CompressedImage ci;
  int dim = 3 * w * h;
  ci.codeVectors.resize(K);
  for (size_t k = 0; k < K; k++) {
    ci.codeVectors[k].resize(dim);
    for (int d = 0; d < dim; d++) ci.codeVectors[k][d] = (char)cb_bytes[k * dim + d];
  }
  ci.assignedCodeVector.assign(assign, assign + N);
  ci.xSize = xSize;
  ci.ySize = ySize;
  ci.blockWidth = w;
  ci.blockHeight = h;
  RGBImage im = CompressedImage::decompress(ci);
  std::memcpy(rgb_out, im.img.data(), im.img.size() * 3);
  return 0;
}

// getQuantizer(LBG)->quantize (src/Quantizer.cpp:121-143).
int ref_quantize(const double *X, size_t N, int dim, int nbits, double eps, double *cb_out,
                 uint64_t *assign_out, double *dist_out) {
  std::vector<Vector> ts = to_vectors(X, N, dim);
  QuantizerPtr q = getQuantizer(Quantizers::LBG);
  std::vector<Vector> cb;
  std::vector<size_t> as;
  VectorType dist;
  std::tie(cb, as, dist) = q->quantize(ts, (size_t)nbits, (VectorType)eps);
  for (size_t k = 0; k < cb.size(); k++)
    for (int d = 0; d < dim; d++) cb_out[k * dim + d] = cb[k][d];
  for (size_t i = 0; i < N; i++) assign_out[i] = as[i];
  *dist_out = dist;
  return (int)cb.size();
}

// KDTree(dim, cb) + nearestNeighbour (src/KDTree.cpp:16-29) over a batch of queries.
int ref_nn(const double *cb, size_t K, int dim, const double *Q, size_t n, uint64_t *out) {
  std::vector<Vector> cbv = to_vectors(cb, K, dim);
  const KDTree tree(dim, cbv);
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) {
    Vector q(dim);
    for (int d = 0; d < dim; d++) q[d] = Q[i * dim + d];
    out[i] = tree.nearestNeighbour(q);
  }
  return 0;
}

// Same, but queries come straight from an RGB buffer through the reference's own block
// extraction; avoids materialising an N*dim double array on the Python side.
int ref_nn_rgb(const double *cb, size_t K, const uint8_t *rgb, int xSize, int ySize, int w, int h,
               int cs, uint64_t *out) {
  RGBImage im = make_image(rgb, xSize, ySize);
  ColorSpacePtr csp = getColorSpace((ColorSpaces)cs);
  std::vector<Vector> v = getBlocksAsVectorsFromImage(im, w, h, csp);
  int dim = 3 * w * h;
  std::vector<Vector> cbv = to_vectors(cb, K, dim);
  const KDTree tree(dim, cbv);
#pragma omp parallel for
  for (size_t i = 0; i < v.size(); i++) out[i] = tree.nearestNeighbour(v[i]);
  return (int)v.size();
}

// Per-level replay of quantize() with the reference's public pieces.
//   cb_pre   : sum_{l=1..nbits} 2^l * dim doubles  (codebook entering level l, after the split)
//   assign   : nbits * N                           (assignment of level l, w.r.t. cb_pre)
//   cb_post  : same shape as cb_pre                (codebook after fixCodeVectors)
//   d0, d1   : nbits each                          (distortion before / after the fix)
// Returns 0 when the replay's final state is bit-identical to quantize()'s, 1 otherwise.
int ref_levels(const double *X, size_t N, int dim, int nbits, double *cb0, double *cb_pre,
               uint64_t *assign, double *cb_post, double *d0, double *d1) {
  std::vector<Vector> ts = to_vectors(X, N, dim);
  // trainingSetSum (src/Quantizer.cpp:46-57) followed by /= N (:129-130)
  Vector sum(dim), c(dim);
  for (size_t i = 0; i < N; i++) {
    Vector y = ts[i] - c;
    Vector t = sum + y;
    c = (t - sum) - y;
    sum = t;
  }
  std::vector<Vector> cb(1);
  cb[0] = sum;
  cb[0] /= (VectorType)N;
  for (int d = 0; d < dim; d++) cb0[d] = cb[0][d];

  std::vector<size_t> as(N);
  double dist = 0;
  size_t off = 0;
  for (int l = 1; l <= nbits; l++) {
    concat(cb, cb);
    for (size_t i = 0; i < cb.size() / 2; i++) {
      cb[i] *= (VectorType)(1 + 0.2);
      cb[i + cb.size() / 2] *= (VectorType)(1 - 0.2);
    }
    size_t K = cb.size();
    for (size_t k = 0; k < K; k++)
      for (int d = 0; d < dim; d++) cb_pre[off + k * dim + d] = cb[k][d];
    {
      const KDTree tree(dim, cb);
#pragma omp parallel for
      for (size_t i = 0; i < N; i++) as[i] = tree.nearestNeighbour(ts[i]);
    }
    for (size_t i = 0; i < N; i++) assign[(size_t)(l - 1) * N + i] = as[i];
    auto distortion = [&]() {
      double r = 0;
      for (size_t i = 0; i < N; i++) r += norm(ts[i] - cb[as[i]]);
      return r / ((VectorType)(N * dim));
    };
    d0[l - 1] = distortion();
    std::vector<std::vector<size_t>> area(K);
    for (size_t i = 0; i < N; i++) area[as[i]].push_back(i);
#pragma omp parallel for
    for (size_t k = 0; k < K; k++) {
      Vector s(dim), cc(dim);
      for (size_t x : area[k]) {
        Vector y = ts[x] - cc;
        Vector t = s + y;
        cc = (t - s) - y;
        s = t;
      }
      cb[k] = s;
      if (area[k].size()) cb[k] /= (VectorType)area[k].size();
    }
    d1[l - 1] = dist = distortion();
    for (size_t k = 0; k < K; k++)
      for (int d = 0; d < dim; d++) cb_post[off + k * dim + d] = cb[k][d];
    off += K * dim;
  }
  (void)dist;
  // Cross-check against the real thing.
  QuantizerPtr q = getQuantizer(Quantizers::LBG);
  std::vector<Vector> rcb;
  std::vector<size_t> ras;
  VectorType rdist;
  std::tie(rcb, ras, rdist) = q->quantize(ts, (size_t)nbits, (VectorType)1e-6f);
  if (rcb.size() != cb.size()) return 1;
  for (size_t k = 0; k < cb.size(); k++)
    for (int d = 0; d < dim; d++)
      if (std::memcmp(&rcb[k][d], &cb[k][d], sizeof(double)) != 0) return 1;
  for (size_t i = 0; i < N; i++)
    if (ras[i] != as[i]) return 1;
  return 0;
}

// CompressedImage::compress (src/Compressor.cpp:107-154) - the reference's whole timed path.
// seconds_out is the reference's own measureExecutionTime value (block extraction + quantize).
int ref_compress(const uint8_t *rgb, int xSize, int ySize, int cs, int w, int h, double eps,
                 int nbits, uint8_t *cb_bytes_out, uint64_t *assign_out, double *distortion_out,
                 float *bpp_out, double *seconds_out) {
  RGBImage im = make_image(rgb, xSize, ySize);
  auto res = CompressedImage::compress(im, Quantizers::LBG, (ColorSpaces)cs, w, h,
                                       (VectorType)eps, nbits);
  CompressedImage &ci = res.first;
  int dim = 3 * w * h;
  if (cb_bytes_out)
    for (size_t k = 0; k < ci.codeVectors.size(); k++)
      for (int d = 0; d < dim; d++) cb_bytes_out[k * dim + d] = (uint8_t)ci.codeVectors[k][d];
  if (assign_out)
    for (size_t i = 0; i < ci.assignedCodeVector.size(); i++)
      assign_out[i] = ci.assignedCodeVector[i];
  if (distortion_out) *distortion_out = res.second.distortion;
  if (bpp_out) *bpp_out = res.second.bitsPerPixel;
  if (seconds_out) *seconds_out = res.second.compressionTime.count();
  return (int)ci.codeVectors.size();
}

// compress + saveToFile (src/Compressor.cpp:190-227): writes a real .quant file.
int ref_compress_to_file(const uint8_t *rgb, int xSize, int ySize, int cs, int w, int h,
                         double eps, int nbits, const char *path) {
  RGBImage im = make_image(rgb, xSize, ySize);
  auto res = CompressedImage::compress(im, Quantizers::LBG, (ColorSpaces)cs, w, h,
                                       (VectorType)eps, nbits);
  res.first.colorSpace = (ColorSpaces)cs;  // the reference leaves this field uninitialised
  res.first.saveToFile(path);
  return 0;
}

// loadFromFile + decompress (src/Compressor.cpp:229-267,156-165). rgb_out must hold
// 3*xSize*ySize bytes; dims are returned through xs/ys.
int ref_decompress_file(const char *path, uint8_t *rgb_out, size_t cap, int *xs, int *ys) {
  CompressedImage ci;
  ci.loadFromFile(path);
  RGBImage im = CompressedImage::decompress(ci);
  *xs = im.xSize;
  *ys = im.ySize;
  if (im.img.size() * 3 > cap) return 1;
  std::memcpy(rgb_out, im.img.data(), im.img.size() * 3);
  return 0;
}

}  // extern "C"
