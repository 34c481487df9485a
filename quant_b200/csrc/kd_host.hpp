// Host-side construction of the KD tree whose traversal order defines which codevector the
// reference returns on exact distance ties.
//
// The reference finds nearest codevectors with nanoflann 1.2.3 (vendored at
// /root/reference/include/external/nanoflann.hpp, wrapped by src/KDTree.cpp:16-29, leaf size 10).
// The GPU path is a brute-force FP32 filter; only queries whose two best candidates are closer
// than the filter's error bound are re-solved exactly, by walking THIS tree on the device with the
// reference's FP64 arithmetic (resolve kernel).  So the tree has to have the same shape, the same
// split planes and the same point order inside leaves as nanoflann's:
//   bounding box   nanoflann.hpp:1021-1043   computeBoundingBox
//   recursion      nanoflann.hpp:1046-1094   divideTree
//   split choice   nanoflann.hpp:1108-1147   middleSplit_
//   partition      nanoflann.hpp:1159-1186   planeSplit
// K <= 65536 points, so the build is micro- to milliseconds on one host core and overlaps the
// assignment kernel of the same level.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "qb200_device.cuh"

namespace qb {

struct KdHostTree {
  std::vector<KdNode> nodes;        // node 0 is the root
  std::vector<unsigned int> order;  // nanoflann's `vind` after the build
  std::vector<double> box_low, box_high;  // root bounding box as left behind by the build
  int depth = 0;
  // Smallest relative distance from flipping of any comparison that shaped the tree (span eligibility, choice of the
  // cut dimension, points against the cutting plane, the clamp of the plane).  Above ~1e-9 the tree keeps its SHAPE and
  // point order for any codebook within a few ulps of this one - what the auto centroid mode needs to know about
  // decisions that hinge on the visiting order (qb200_api.cu, resolve_bruteforce_kernel).  0: degenerate (duplicates).
  double min_margin = 0.0;
};

// points: K x dim, row-major FP64 (colour-space domain, exactly the doubles the reference holds).
// exact (may be null): per point, 1 when its coordinates are the same numbers with either centroid arithmetic (dead
// cells, children of one-vector cells) - only used for the robustness census (min_margin).
void build_kd_tree(const double *points, size_t K, int dim, int leaf_max, KdHostTree &out, const unsigned char *exact = nullptr);

}  // namespace qb
