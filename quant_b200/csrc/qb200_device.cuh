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
  // Fast path (set by the host when no element of any vector falls outside its image - block sizes divide
  // the image - and every byte offset fits 31 bits): 32-bit address arithmetic, no per-element checks,
  // divisions by hB / per_image through precomputed multipliers.
  unsigned int fast;
  unsigned int multi;                  // more than one image in the batch
  unsigned int hb_mul, hb_shift;       // v / hB        (fastdiv)
  unsigned int pi_mul, pi_shift;       // v / per_image (fastdiv)
  unsigned int row_stride32, img_bytes32, first_vec32, per_image32, origin32;
  // Dense copy of the training set made once per set_image/set_vectors (pack_vectors_kernel): local vector v is
  // the dense_stride bytes at dense + v * dense_stride (dim rounded up to a multiple of 4, padding bytes 0).
  // When present, the hot kernels read whole words from it instead of gathering bytes from the image.
  const uint8_t *dense;
  unsigned int dense_stride;
  // General FP64 training vectors (qb200_set_vectors_f64, CIE1931 images): n_local x dim doubles, row-major.
  // When set, the byte fields above are not used by the training kernels: the filter runs on the values
  // rounded to FP32, the resolver and the centroid sums on the doubles themselves (qb200_generic.cu).
  const double *f64;
};

// Division of a 32-bit n by an invariant d (Granlund-Montgomery): q = (t + ((n - t) >> 1)) >> (l - 1),
// t = umulhi(m, n), l = ceil(log2 d), m = floor(2^32 * (2^l - d) / d) + 1.  d == 1: mul = 0, shift = 0.
__host__ inline void fastdiv_make(unsigned int d, unsigned int &mul, unsigned int &shift) {
  unsigned int l = 0;
  while ((1ull << l) < d) l++;
  mul = (unsigned int)((((1ull << l) - d) << 32) / d + 1);
  shift = l;
}
__device__ __forceinline__ unsigned int fastdiv(unsigned int n, unsigned int mul, unsigned int shift) {
  if (shift == 0) return n;  // d == 1
  const unsigned int t = __umulhi(mul, n);
  return (t + ((n - t) >> 1)) >> (shift - 1);
}

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

// Fast path only (s.fast != 0): address of the first byte of local vector v's block.
__device__ __forceinline__ const signed char *fast_vec_ptr(const VecSource &s, unsigned long long v_local) {
  unsigned int v = s.first_vec32 + (unsigned int)v_local, img_off = 0;
  if (s.multi) {
    const unsigned int img = fastdiv(v, s.pi_mul, s.pi_shift);
    v -= img * s.per_image32;
    img_off = img * s.img_bytes32;
  }
  const unsigned int i = fastdiv(v, s.hb_mul, s.hb_shift), j = v - i * s.hB;
  return reinterpret_cast<const signed char *>(s.buf) + (img_off + i * s.row_stride32 + j * s.col_stride - s.origin32);
}

// All DIM lattice values of local vector v (template DIM: fully unrolled, values stay in registers).
template <int DIM, typename T>
__device__ __forceinline__ void gather_lattice(const VecSource &s, unsigned long long v_local, T *out) {
  if (s.dense) {
    constexpr int WORDS = (DIM + 3) / 4;
    const unsigned int *p = reinterpret_cast<const unsigned int *>(s.dense + v_local * (unsigned long long)s.dense_stride);
    unsigned int w[WORDS];
    if constexpr (WORDS % 4 == 0) {  // 16-byte rows (dim 48: 3 x 128-bit loads)
#pragma unroll
      for (int i = 0; i < WORDS / 4; i++) {
        const uint4 t = __ldg(reinterpret_cast<const uint4 *>(p) + i);
        w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < WORDS; i++) w[i] = __ldg(p + i);
    }
#pragma unroll
    for (int e = 0; e < DIM; e++) out[e] = (T)(int)(signed char)(w[e >> 2] >> (8 * (e & 3)));
  } else if (s.fast) {
    const signed char *p = fast_vec_ptr(s, v_local);
#pragma unroll
    for (int e = 0; e < DIM; e++) out[e] = (T)(int)__ldg(p + s.elem_off[e]);
  } else {
    unsigned long long base, img;
    vec_base(s, v_local, base, img);
#pragma unroll
    for (int e = 0; e < DIM; e++) out[e] = (T)load_lattice(s, img, base, e);
  }
}

// Flattened KD tree in nanoflann's shape (see kd_host.hpp); 32 bytes per node.
constexpr int kKdFeatMask = 0xffff;      // inner node: a & kKdFeatMask = divfeat
constexpr int kKdDivLowExact = 1 << 16;  // inner node, census of the auto centroid mode only: divlow / divhigh is a number
constexpr int kKdDivHighExact = 1 << 17; // the integer-sum and the compensated-sum codebooks share bit for bit
constexpr int kKdNodeFragile = 1 << 18;  // inner node, census only: a comparison that shaped this node's split is within
                                         // kKdRobustMargin of flipping - everything UNDER the node may be arranged differently
                                         // in the reference's tree (nothing outside it is affected)
constexpr double kKdRobustMargin = 1e-9;
struct KdNode {
  int child1, child2;  // -1/-1: leaf
  int a;               // inner: divfeat (+ the two census bits above); leaf: left (first position in vind)
  int b;               // leaf: right (one past the last position in vind); inner: first position of child2's range
  double divlow, divhigh;
};

struct KdDevice {
  const KdNode *nodes;
  const unsigned int *vind;
  const unsigned int *inv;   // inverse of vind: position of codevector k in traversal (leaf) order
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
