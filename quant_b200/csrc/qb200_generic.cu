// Kernels of the general FP64-vector path: training vectors that are not on the byte lattice -
// AbstractQuantizer::quantize on arbitrary doubles (include/Quantizer.hpp:12-14) and the CIE1931 colour space
// (/root/reference/src/ColorSpace.cpp:31-48).  Integer per-cell statistics do not exist for such data, so this
// path keeps the filter + exact resolver for the assignment (the filter sees the vectors rounded to FP32), sums
// the members with the reference's own compensated loop (qb200_exact.cu) and evaluates updateDistortion in FP64.
#include "qb200_launch.hpp"

namespace qb {

namespace {

// Cie1931::RGBtoColorSpace: c[i] are signed chars; every sum is evaluated left to right, then divided by 0.17697.
// Spelled with round-to-nearest intrinsics: the reference's x86-64 build has no fused multiply-add.
__global__ void cie_vectors_kernel(const VecSource src, double *__restrict__ out) {
  const int ppb = src.dim / 3;  // pixels per block
  const unsigned long long total = src.n_local * (unsigned long long)ppb;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long v = i / ppb;
    const int e = (int)(i - v * ppb) * 3;
    unsigned long long base, img;
    vec_base(src, v, base, img);
    const unsigned long long o = base + src.elem_off[e];
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;  // a pixel past the end of the image is the literal 0.0 (src/Compressor.cpp:53-57)
    if (o < src.img_bytes) {
      const uint8_t *p = src.buf + (img * src.img_bytes + o - src.origin);
      const double c0 = (double)(int)(signed char)__ldg(p), c1 = (double)(int)(signed char)__ldg(p + 1),
                   c2 = (double)(int)(signed char)__ldg(p + 2);
      r0 = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(c0, 0.490), __dmul_rn(c1, 0.310)), __dmul_rn(c2, 0.200)), 0.17697);
      r1 = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(c0, 0.17697), __dmul_rn(c1, 0.81240)), __dmul_rn(c2, 0.01063)), 0.17697);
      r2 = __ddiv_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(c1, 0.01)), __dmul_rn(c2, 0.99)), 0.17697);
    }
    double *dst = out + v * (unsigned long long)src.dim + e;
    dst[0] = r0;
    dst[1] = r1;
    dst[2] = r2;
  }
}

// One vector per thread (grid-stride), norm(x - c) as the reference spells it (x - c, then the squares added left
// to right, include/VectorOperations.hpp:107-111); fixed grid + fixed shared-memory tree: deterministic.
__global__ void __launch_bounds__(256)
    distortion_f64_kernel(const VecSource src, const uint32_t *__restrict__ assign, const double *__restrict__ cb,
                          double *__restrict__ partials) {
  __shared__ double s_acc[256];
  const int dim = src.dim;
  double acc = 0.0;
  for (unsigned long long v = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; v < src.n_local;
       v += (unsigned long long)gridDim.x * blockDim.x) {
    const double *x = src.f64 + v * dim, *c = cb + (size_t)assign[v] * dim;
    double r = 0.0;
    for (int e = 0; e < dim; e++) {
      const double d = __dsub_rn(x[e], c[e]);
      r = __dadd_rn(r, __dmul_rn(d, d));
    }
    acc = __dadd_rn(acc, r);
  }
  s_acc[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_acc[threadIdx.x] = __dadd_rn(s_acc[threadIdx.x], s_acc[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = s_acc[0];
}

__global__ void sum_partials_kernel(const double *__restrict__ partials, const int n, double *__restrict__ slots,
                                    const int rank, const int world) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double acc = 0.0;
  for (int i = 0; i < n; i++) acc = __dadd_rn(acc, partials[i]);
  for (int r = 0; r < world; r++) slots[r] = r == rank ? acc : 0.0;
}

}  // namespace

cudaError_t launch_sum_partials(const double *partials, int n_partials, double *slots, int rank, int world,
                                cudaStream_t stream) {
  sum_partials_kernel<<<1, 32, 0, stream>>>(partials, n_partials, slots, rank, world);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_cie_vectors(const VecSource &src, double *out, int sm_count, cudaStream_t stream) {
  const unsigned long long total = src.n_local * (unsigned long long)(src.dim / 3);
  if (total == 0) return cudaSuccess;
  unsigned long long blocks = (total + 255) / 256;
  if (blocks > (unsigned long long)sm_count * 16) blocks = (unsigned long long)sm_count * 16;
  cie_vectors_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(src, out);
  count_launch();
  return cudaGetLastError();
}

int distortion_blocks(int sm_count) { return sm_count * 4; }

cudaError_t launch_distortion_f64(const VecSource &src, const uint32_t *assign, const double *cb, double *partials,
                                  int sm_count, cudaStream_t stream) {
  distortion_f64_kernel<<<distortion_blocks(sm_count), 256, 0, stream>>>(src, assign, cb, partials);
  count_launch();
  return cudaGetLastError();
}

}  // namespace qb
