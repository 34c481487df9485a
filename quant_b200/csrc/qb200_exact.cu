// Opt-in bit-exact centroid sums (qb200_set_exact_centroids / QB200_EXACT_CENTROIDS=1).
//
// The reference sums a cell's members with a compensated (Kahan) loop over FP64 values t/255.0 in ascending
// vector order (Solution::sumInArea / trainingSetSum, /root/reference/src/Quantizer.cpp:46-70) and divides by
// the member count (:81-85, :129-130).  The default path of this library derives the centroid from the integer
// sum S_t instead, ((double)S_t / 255) / n, which agrees to <= 4e-16 relative but not always in the last bit; on
// inputs full of duplicated vectors (palettes, flat areas) that last bit decides exact ties of the NEXT split
// level (x against 1.2c / 0.8c with c == x), so end-to-end results can leave the reference's there.  A rounded
// compensated sum has no closed form in the integer statistics (its second-order error depends on the order
// of the addends), so this mode executes the reference's own operation sequence:
//   1. stable radix sort of (cell, local vector index)  -> members of every cell in ascending index order
//   2. kahan_sums_kernel: one warp per cell, one lane per dimension, the four dependent FP64 operations of the
//      reference's loop per member (the chain is latency bound: ~N steps at K = 1, ~N/K at level K)
//   3. finalize_split_kernel divides the sums by n.
// The NORMAL colour space never needs it: its addends are integers, every partial sum is exact.
// Ranks of a sharded run continue each other's chains in rank order (see exact_centroid_sums in qb200_api.cu).
#include "qb200_launch.hpp"

#include <cub/device/device_radix_sort.cuh>

namespace qb {

namespace {

__global__ void iota_kernel(uint32_t *__restrict__ out, const unsigned long long n) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x)
    out[i] = (uint32_t)i;
}

__device__ __forceinline__ void cp_async_4(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned int)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned int)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

// One warp per cell.  The members' rows (dense bytes, or doubles for general FP64 vectors) are copied to shared
// memory with cp.async one batch ahead and their indices are loaded two batches ahead, so the only thing the warp
// ever waits for is its own FP64 chain.  DCH = dimensions per lane (lane handles e = lane + 32 m); `batch` members
// (<= 128, up to four per lane) are staged at a time.  The code is kept small on purpose: the loop streams through the
// instruction cache once per batch.
template <int DCH, bool F64>
__global__ void __launch_bounds__(128)
    kahan_sums_kernel(const VecSource src, const uint32_t *__restrict__ keys_sorted, const uint32_t *__restrict__ order,
                      const int K, const int scaled, const int batch, double *__restrict__ state,
                      unsigned long long *__restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *s_val = reinterpret_cast<double *>(smem_raw);  // colour-space value of a raw byte (src/ColorSpace.cpp:16-21)
  if (!F64) {
    for (int u = threadIdx.x; u < 256; u += blockDim.x) {
      const double L = (double)(int)(signed char)u;
      s_val[u] = scaled ? __ddiv_rn(__dadd_rn(L, 128.0), 255.0) : L;
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * (blockDim.x >> 5) + warp;
  if (k >= K) return;
  const int dim = src.dim;
  const unsigned int stride = F64 ? (unsigned int)dim * 8u : src.dense_stride;  // bytes per staged row
  const unsigned int words = F64 ? (unsigned int)dim : stride >> 2;             // cp.async transfers per row
  unsigned char *buf = smem_raw + 2048 + (size_t)warp * 2 * batch * stride;     // [2][batch][stride]
  const unsigned int n = (unsigned int)src.n_local;
  // members of cell k: positions [beg, end) of the sorted list (K == 1: the whole set in its own order)
  unsigned int beg = 0, end = n;
  if (order) {
    const unsigned int target = (unsigned int)k + (lane & 1);  // lane 0: first key >= k, lane 1: first key >= k + 1
    unsigned int lo = 0, hi = n;
    if (lane < 2)
      while (lo < hi) {
        const unsigned int mid = lo + ((hi - lo) >> 1);
        if (__ldg(keys_sorted + mid) < target) lo = mid + 1; else hi = mid;
      }
    beg = __shfl_sync(0xffffffffu, lo, 0);
    end = __shfl_sync(0xffffffffu, lo, 1);
  }
  if (counts && lane == 0) counts[k] = (unsigned long long)(end - beg);
  double sum[DCH], c[DCH];
#pragma unroll
  for (int m = 0; m < DCH; m++) {
    const int e = lane + 32 * m;
    sum[m] = e < dim ? state[((size_t)k * dim + e) * 2] : 0.0;
    c[m] = e < dim ? state[((size_t)k * dim + e) * 2 + 1] : 0.0;
  }
  // lane's members of the batch starting at j: positions j + lane + 32 r (r < kRows) inside [j, j + batch) and [beg, end)
  constexpr int kRows = 4;
  auto member = [&](unsigned int j, unsigned int (&out)[kRows]) {
#pragma unroll
    for (int r = 0; r < kRows; r++) {
      const unsigned int o = lane + 32 * r;
      const bool ok = (int)o < batch && j < end && o < end - j;
      out[r] = !ok ? 0xffffffffu : order ? __ldg(order + j + o) : j + o;
    }
  };
  auto stage = [&](const unsigned int (&mine)[kRows], int slot) {  // every lane copies its own members' rows, asynchronously
#pragma unroll
    for (int r = 0; r < kRows; r++) {
      if (mine[r] == 0xffffffffu) continue;
      unsigned char *dst = buf + ((size_t)slot * batch + lane + 32 * r) * stride;
      if (F64) {
        const double *row = src.f64 + (unsigned long long)mine[r] * dim;
        for (unsigned int w = 0; w < words; w++) cp_async_8(dst + 8 * w, row + w);
      } else {
        const unsigned char *row = src.dense + (unsigned long long)mine[r] * stride;
        for (unsigned int w = 0; w < words; w++) cp_async_4(dst + 4 * w, row + 4 * w);
      }
    }
    cp_async_commit();
  };
  // value of element (lane + 32 m) of staged member s
  auto value = [&](const unsigned char *rows, unsigned int s, int m) -> double {
    const int e = lane + 32 * m;
    if (F64) return reinterpret_cast<const double *>(rows + s * stride)[e < dim ? e : 0];
    return s_val[rows[s * stride + (e < (int)stride ? e : 0)]];
  };
  // src/Quantizer.cpp:64-67:  y = x - c;  t = sum + y;  c = (t - sum) - y;  sum = t
  auto step = [&](int m, double x) {
    const double y = __dsub_rn(x, c[m]);
    const double t = __dadd_rn(sum[m], y);
    c[m] = __dsub_rn(__dsub_rn(t, sum[m]), y);
    sum[m] = t;
  };
  constexpr int G = DCH == 1 ? 8 : DCH == 2 ? 4 : 2;  // members whose values are fetched ahead of the chain
  unsigned int idx_next[kRows];
  member(beg, idx_next);
  stage(idx_next, 0);
  member(beg + batch, idx_next);
  int slot = 0;
  for (unsigned int j0 = beg; j0 < end; j0 += batch, slot ^= 1) {
    const unsigned int cnt = min((unsigned int)batch, end - j0);
    stage(idx_next, slot ^ 1);                       // batch j0 + batch (an empty group past the end)
    member(j0 + 2 * batch, idx_next);                // consumed one iteration later
    cp_async_wait<1>();                              // batch j0 has landed
    __syncwarp();
    const unsigned char *rows = buf + (size_t)slot * batch * stride;
    unsigned int s = 0;
    if (cnt >= (unsigned int)G) {
      double xa[G][DCH], xb[G][DCH];
#pragma unroll
      for (int g = 0; g < G; g++)
#pragma unroll
        for (int m = 0; m < DCH; m++) xa[g][m] = value(rows, g, m);
      for (; s + 2 * G <= cnt; s += G) {             // values of the next group load while this group's chain runs
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
          for (int m = 0; m < DCH; m++) xb[g][m] = value(rows, s + G + g, m);
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
          for (int m = 0; m < DCH; m++) step(m, xa[g][m]);
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
          for (int m = 0; m < DCH; m++) xa[g][m] = xb[g][m];
      }
#pragma unroll
      for (int g = 0; g < G; g++)
#pragma unroll
        for (int m = 0; m < DCH; m++) step(m, xa[g][m]);
      s += G;
    }
    for (; s < cnt; s++)
#pragma unroll
      for (int m = 0; m < DCH; m++) step(m, value(rows, s, m));
    __syncwarp();                                    // everyone is done with this slot before it is refilled
  }
  cp_async_wait<0>();
#pragma unroll
  for (int m = 0; m < DCH; m++) {
    const int e = lane + 32 * m;
    if (e < dim) {
      state[((size_t)k * dim + e) * 2] = sum[m];
      state[((size_t)k * dim + e) * 2 + 1] = c[m];
    }
  }
}

}  // namespace

size_t exact_sort_temp_bytes(size_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                  (uint32_t *)nullptr, (long long)n, 0, 32, (cudaStream_t)0);
  return bytes;
}

cudaError_t launch_exact_iota(uint32_t *iota, size_t n, int sm_count, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  unsigned long long blocks = (n + 1023) / 1024;
  if (blocks > (unsigned long long)sm_count * 8) blocks = (unsigned long long)sm_count * 8;
  iota_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(iota, n);
  count_launch();
  return cudaGetLastError();
}

// Stable sort of (assign[v], v) by cell: keys_out ascending, order = the members of cell 0, cell 1, ... each in
// ascending v.  CUB's radix sort is the one library call of this mode (stable by construction).
cudaError_t launch_exact_sort(const uint32_t *assign, uint32_t *keys_out, const uint32_t *iota, uint32_t *order, size_t n,
                              int key_bits, void *tmp, size_t tmp_bytes, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, assign, keys_out, iota, order, (long long)n, 0, key_bits, stream);
}

// state: K*dim pairs {sum, c}, read as the chain's initial state and overwritten with its final one.
// counts (may be null): members of every cell on this context.  Reads the dense byte copy of the training set
// (always made by set_image / set_vectors_u8) or, for general FP64 vectors, the doubles themselves.
cudaError_t launch_kahan_sums(const VecSource &src, const uint32_t *keys_sorted, const uint32_t *order, int K, int scaled,
                              double *state, unsigned long long *counts, cudaStream_t stream) {
  const bool f64 = src.f64 != nullptr;
  if (!f64 && !src.dense) return cudaErrorInvalidValue;
  const int dim = src.dim;
  const size_t row = f64 ? (size_t)dim * 8 : src.dense_stride;
  int batch = 128;  // long batches amortise the per-batch bookkeeping of the latency-bound chain
  while (batch > 4 && 2 * (size_t)batch * row > 40 * 1024) batch >>= 1;
  int warps = (int)((40 * 1024) / (2 * (size_t)batch * row));
  warps = warps < 1 ? 1 : warps > 4 ? 4 : warps;
  const int blocks = (K + warps - 1) / warps;
  const size_t smem = 2048 + (size_t)warps * 2 * batch * row;
#define QB_KAHAN(DCH)                                                                                                    \
  (f64 ? kahan_sums_kernel<DCH, true><<<blocks, 32 * warps, smem, stream>>>(src, keys_sorted, order, K, scaled, batch,  \
                                                                            state, counts)                              \
       : kahan_sums_kernel<DCH, false><<<blocks, 32 * warps, smem, stream>>>(src, keys_sorted, order, K, scaled, batch, \
                                                                             state, counts))
  if (dim <= 32)
    QB_KAHAN(1);
  else if (dim <= 64)
    QB_KAHAN(2);
  else if (dim <= 96)
    QB_KAHAN(3);
  else
    QB_KAHAN(6);
#undef QB_KAHAN
  count_launch();
  return cudaGetLastError();
}

}  // namespace qb
