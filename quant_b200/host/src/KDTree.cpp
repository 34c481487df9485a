// KDTree on top of libqb200 - replaces /root/reference/src/KDTree.cpp (nanoflann wrapper).
#include "KDTree.hpp"

#include <cmath>
#include <cstdint>
#include <stdexcept>

#include "../../../include/qb200.h"
#include "B200Context.hpp"

class KDTree::KDTreeImpl {
 public:
  size_t dim = 0;
  std::vector<double> codebook;  // K x dim, as given
};

namespace {
// same lattice rule as Quantizer.cpp: NORMAL = integers in [-128, 127], SCALED = t/255.0 with t in [0, 255];
// -1: neither (the queries then go to the library as FP64 vectors)
int lattice_of(const std::vector<Vector> &pts, size_t dim, std::vector<uint8_t> &bytes) {
  bool normal = true, scaled = true;
  for (const Vector &p : pts) {
    if (p.size() != dim) throw std::runtime_error("KDTree: query of the wrong dimension");
    for (double x : p) {
      if (normal && !(x == std::nearbyint(x) && x >= -128 && x <= 127)) normal = false;
      const double t = std::nearbyint(x * 255.0);
      if (scaled && !(t >= 0 && t <= 255 && t / 255.0 == x)) scaled = false;
    }
  }
  if (!normal && !scaled) return -1;
  bytes.resize(pts.size() * dim);
  for (size_t i = 0; i < pts.size(); i++)
    for (size_t d = 0; d < dim; d++)
      bytes[i * dim + d] = normal ? (uint8_t)(int8_t)(int)pts[i][d] : (uint8_t)((int)std::nearbyint(pts[i][d] * 255.0) ^ 0x80);
  return normal ? QB200_CS_NORMAL : QB200_CS_SCALED;
}
}  // namespace

KDTree::KDTree(size_t dim, const std::vector<Vector> &points) : impl(new KDTreeImpl()) {
  impl->dim = dim;
  impl->codebook.reserve(points.size() * dim);
  for (const Vector &p : points) {
    if (p.size() != dim) throw std::runtime_error("KDTree: point of the wrong dimension");
    impl->codebook.insert(impl->codebook.end(), p.begin(), p.end());
  }
}

KDTree::~KDTree() = default;

std::vector<size_t> KDTree::nearestNeighbours(const std::vector<Vector> &pts) const {
  std::vector<size_t> out(pts.size());
  if (pts.empty()) return out;
  const size_t K = impl->codebook.size() / impl->dim;
  if (K == 0) throw std::runtime_error("KDTree: no points");
  std::vector<uint8_t> bytes;
  const int cs = lattice_of(pts, impl->dim, bytes);
  qb200_ctx *ctx = qbhost::context();
  if (cs >= 0) {
    qbhost::check(qb200_set_vectors_u8(ctx, bytes.data(), pts.size(), (int)impl->dim, cs, 0), "qb200_set_vectors_u8");
  } else {
    std::vector<double> flat;
    flat.reserve(pts.size() * impl->dim);
    for (const Vector &p : pts) flat.insert(flat.end(), p.begin(), p.end());
    qbhost::check(qb200_set_vectors_f64(ctx, flat.data(), pts.size(), (int)impl->dim, 0), "qb200_set_vectors_f64");
  }
  std::vector<uint32_t> idx(pts.size());
  qbhost::check(qb200_assign_accumulate(ctx, impl->codebook.data(), (uint32_t)K, idx.data(), nullptr, nullptr, nullptr, nullptr),
                "qb200_assign_accumulate");
  for (size_t i = 0; i < idx.size(); i++) out[i] = idx[i];
  return out;
}

size_t KDTree::nearestNeighbour(const Vector &pt) const { return nearestNeighbours(std::vector<Vector>(1, pt))[0]; }
