// getQuantizer / LBG quantiser backed by libqb200 - replaces /root/reference/src/Quantizer.cpp.
//
// The reference's LBGQuantizer::quantize (src/Quantizer.cpp:121-143) and everything under it
// (assignCodeVectors :24-32, updateDistortion :9-22, fixCodeVectors :72-87, the split :134-138)
// run on the GPU behind qb200_train.  This file only marshals the generic vector<Vector>
// signature: the fast GPU path works on the byte lattice image data lives on, so the doubles are mapped
// back to bytes when that is exact; any other input goes to the library as FP64 vectors
// (qb200_set_vectors_f64).
#include "Quantizer.hpp"

#include <cmath>
#include <cstdint>
#include <stdexcept>

#include "../../../include/qb200.h"
#include "B200Context.hpp"

namespace {

// NORMAL lattice: every element is an integer in [-128, 127] (byte = value mod 256).
// SCALED lattice: every element equals t/255.0 for an integer t in [0, 255] (byte = t xor 0x80).
// Returns the colour space of the lattice, or -1 when the vectors are on neither.
int to_lattice_bytes(const std::vector<Vector> &set, size_t dim, std::vector<uint8_t> &bytes) {
  bytes.resize(set.size() * dim);
  bool normal = true, scaled = true;
  for (size_t i = 0; i < set.size() && (normal || scaled); i++) {
    if (set[i].size() != dim) throw std::runtime_error("quantize: vectors of unequal dimension");
    for (size_t d = 0; d < dim; d++) {
      const double x = set[i][d];
      if (normal && !(x == std::nearbyint(x) && x >= -128 && x <= 127)) normal = false;
      if (scaled) {
        const double t = std::nearbyint(x * 255.0);
        if (!(t >= 0 && t <= 255 && t / 255.0 == x)) scaled = false;
      }
    }
  }
  if (!normal && !scaled) return -1;
  const int cs = normal ? QB200_CS_NORMAL : QB200_CS_SCALED;
  for (size_t i = 0; i < set.size(); i++)
    for (size_t d = 0; d < dim; d++) {
      const double x = set[i][d];
      bytes[i * dim + d] = normal ? (uint8_t)(int8_t)(int)x : (uint8_t)((int)std::nearbyint(x * 255.0) ^ 0x80);
    }
  return cs;
}

class LBGQuantizer : public AbstractQuantizer {
 public:
  std::tuple<std::vector<Vector>, std::vector<size_t>, VectorType> quantize(const std::vector<Vector> &trainingSet,
                                                                            size_t n, VectorType eps) override {
    const size_t dim = trainingSet.at(0).size();  // empty set: std::out_of_range, as in the reference
    std::vector<uint8_t> bytes;
    const int cs = to_lattice_bytes(trainingSet, dim, bytes);
    qb200_ctx *ctx = qbhost::context();
    if (cs >= 0) {
      qbhost::check(qb200_set_vectors_u8(ctx, bytes.data(), trainingSet.size(), (int)dim, cs, 0), "qb200_set_vectors_u8");
    } else {  // arbitrary doubles
      std::vector<double> flat(trainingSet.size() * dim);
      for (size_t i = 0; i < trainingSet.size(); i++) {
        if (trainingSet[i].size() != dim) throw std::runtime_error("quantize: vectors of unequal dimension");
        for (size_t d = 0; d < dim; d++) flat[i * dim + d] = trainingSet[i][d];
      }
      qbhost::check(qb200_set_vectors_f64(ctx, flat.data(), trainingSet.size(), (int)dim, 0), "qb200_set_vectors_f64");
    }
    const size_t K = (size_t)1 << n;
    std::vector<double> cb(K * dim);
    double distortion = 0;
    qbhost::check(qb200_train(ctx, (int)n, eps, QB200_MODE_PARITY, 0, nullptr, nullptr, cb.data(), &distortion, nullptr),
                  "qb200_train");
    static_assert(sizeof(size_t) == sizeof(uint64_t), "size_t must be 64-bit");
    std::vector<size_t> assign(trainingSet.size());
    qbhost::check(qb200_get_assign_u64(ctx, reinterpret_cast<uint64_t *>(assign.data())), "qb200_get_assign_u64");
    std::vector<Vector> codebook(K);
    for (size_t k = 0; k < K; k++) codebook[k] = Vector(cb.begin() + k * dim, cb.begin() + (k + 1) * dim);
    return std::make_tuple(std::move(codebook), std::move(assign), distortion);
  }
};

}  // namespace

QuantizerPtr getQuantizer(Quantizers q) {
  if (q == Quantizers::LBG) return QuantizerPtr(new LBGQuantizer());
  return nullptr;  // MEDIAN_CUT, LBG_MEDIAN_CUT, ABC: not implemented in the reference either
}
