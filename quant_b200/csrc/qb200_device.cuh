// Device-side data layout shared by the kernels of libqb200 (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qb {

constexpr int kMaxDim = 192;  // 8x8 blocks

// How a block vector is gathered from the raw byte buffer.  Restates the indexing of
// getBlocksAsVectorsFromImage (/root/reference/src/Compressor.cpp:31-62): pixel index
// x*ySize + y (x is the slow axis), vector index i*hBlocks + j, element ((dx*h)+dy)*3 + ch.
// A flat N x dim byte matrix (qb200_set_vectors_u8) is the special case hB = 1.
struct VecSource {
  const uint8_t *buf;             // device bytes; buf[0] is image byte `origin` of image 0
  unsigned long long origin;      // first image byte present in buf (shards); 0 otherwise
  unsigned long long img_bytes;   // xSize*ySize*3: bytes of ONE image; elements at or past it read as pad
  unsigned long long per_image;   // vectors per image (wBlocks*hBlocks)
  unsigned long long first_vec;   // in-image index of this context's first vector (shards)
  unsigned long long n_local;     // vectors owned by this context
  unsigned long long row_stride;  // bytes between block rows i and i+1  (w*ySize*3)
  unsigned int col_stride;        // bytes between blocks j and j+1      (h*3)
  unsigned int hB;                // blocks per block row
  int dim;                        // 3*w*h
  int pad_lattice;                // lattice value of the colour-space value 0.0 (SCALED -128, NORMAL 0)
  // Byte offset of element e relative to the block's first byte:
  // ((e/3)/h)*ySize*3 + ((e/3)%h)*3 + e%3.  Lives in the kernel parameter (constant) bank.
  unsigned int elem_off[kMaxDim];
};

// Base byte offset (inside its image) and image number of local vector v.
__device__ __forceinline__ void vec_base(const VecSource &s, unsigned long long v_local,
                                         unsigned long long &base, unsigned long long &img) {
  unsigned long long v = s.first_vec + v_local;
  img = 0;
  if (v >= s.per_image) {  // multi-image batches only
    img = v / s.per_image;
    v -= img * s.per_image;
  }
  unsigned long long i, j;
  if (v < 0xffffffffull) {
    unsigned int vi = (unsigned int)v;
    i = vi / s.hB;
    j = vi - (unsigned int)i * s.hB;
  } else {
    i = v / s.hB;
    j = v - i * s.hB;
  }
  base = i * s.row_stride + j * (unsigned long long)s.col_stride;
}

// Lattice value L = (int8)byte of element e of the vector at (img, base); pad past the image end.
__device__ __forceinline__ int load_lattice(const VecSource &s, unsigned long long img,
                                            unsigned long long base, int e) {
  unsigned long long o = base + s.elem_off[e];
  if (o >= s.img_bytes) return s.pad_lattice;
  return (int)(signed char)__ldg(s.buf + (img * s.img_bytes + o - s.origin));
}

// Flattened KD tree in nanoflann's shape (see kd_host.hpp); 32 bytes per node.
struct KdNode {
  int child1, child2;  // -1/-1: leaf
  int a;               // inner: divfeat; leaf: left (first position in vind)
  int b;               // leaf: right (one past the last position in vind)
  double divlow, divhigh;
};

struct KdDevice {
  const KdNode *nodes;
  const unsigned int *vind;
  const double *bbox_low, *bbox_high;  // root bounding box, dim entries each
  int n_nodes;
  int depth;
};

// Per-cell integer statistics: row k = { n_k, S_k[0..dim), Q_k } as 64-bit words
// (S two's complement).  This is also the all-reduce payload: K*(dim+2) words.
__host__ __device__ __forceinline__ size_t stats_words(unsigned int K, int dim) {
  return (size_t)K * (size_t)(dim + 2);
}

}  // namespace qb
