// Nearest-codevector search with the reference's public surface (/root/reference/include/KDTree.hpp:7-16,
// src/KDTree.cpp:16-29): KDTree(dim, points) + nearestNeighbour(pt).  The reference wraps nanoflann; here the
// search runs on the B200 through qb200_assign_accumulate (exact: same index the reference's tree returns,
// ties included).  A single query per call wastes the GPU - nearestNeighbours() takes a batch, which is what
// encode-only use (a fixed trained codebook, BASELINE config 5) should call.
// Points and queries must lie on the NORMAL or SCALED byte lattice; otherwise std::runtime_error.
#pragma once
#include <memory>
#include <vector>

#include "VectorOperations.hpp"

class KDTree {
 public:
  KDTree(size_t dim, const std::vector<Vector> &points);
  size_t nearestNeighbour(const Vector &pt) const;
  std::vector<size_t> nearestNeighbours(const std::vector<Vector> &pts) const;  // extension: one GPU pass
  ~KDTree();

 private:
  class KDTreeImpl;
  std::unique_ptr<KDTreeImpl> impl;
};
