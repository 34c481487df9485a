// Nearest-codevector search with the reference's public surface (/root/reference/include/KDTree.hpp:7-16,
// src/KDTree.cpp:16-29): construct from K points of one dimension, ask for the nearest point of a query.
// The reference wraps nanoflann; here the search runs on the B200 through qb200_assign_accumulate and is
// exact - the index the reference's tree returns, its tie order included.
//
// One query per call wastes the GPU: nearestNeighbours() (an extension) takes a batch, which is what
// encode-only use against a fixed trained codebook (BASELINE config 5) should call.  Queries on the NORMAL or
// SCALED byte lattice (everything that came from an image) take the fast path; any other doubles are searched
// as FP64 vectors, with the same exact result.
#pragma once
#include <memory>
#include <vector>

#include "VectorOperations.hpp"

class KDTree {
  class KDTreeImpl;                  // keeps a copy of the points (the codebook); no device state of its own
  std::unique_ptr<KDTreeImpl> impl;

 public:
  KDTree(size_t dim, const std::vector<Vector> &points);
  ~KDTree();

  size_t nearestNeighbour(const Vector &pt) const;                               // the reference's entry point
  std::vector<size_t> nearestNeighbours(const std::vector<Vector> &pts) const;  // extension: one GPU pass
};
