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
//      (in-tree: qb200_sort.cu)
//   2. kahan_sums_kernel: one warp per cell, one lane per dimension, the four dependent FP64 operations of the
//      reference's loop per member (the chain is latency bound: ~N steps at K = 1, ~N/K at level K)
//   3. finalize_split_kernel divides the sums by n.
// The NORMAL colour space never needs it: its addends are integers, every partial sum is exact.
// Ranks of a sharded run continue each other's chains in rank order (see exact_centroid_sums in qb200_api.cu).
#include "qb200_launch.hpp"
#include "qb200_exact_fast.cuh"

#include <cmath>

namespace qb {

namespace {

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

// ------------------------------------------------------------------------------------------------
// Parallel evaluation of the same sums (SCALED lattice vectors): see qb200_exact_fast.cuh for the method.
// ------------------------------------------------------------------------------------------------
namespace {

using fx::i128;
using fx::u128;

// Workspace carved out of one device block by launch_kahan_sums_fast (all offsets 256-byte aligned).
struct FxWork {
  uint32_t *cell_beg;   // K + 1 positions into the sorted member list
  uint32_t *win_off;    // K + 1: first nominal window of every cell (windows are counted per cell, shared by its dimensions)
  u128 *sumX;           // per (window, dimension), chain-major: window sums, then (in place) their exclusive prefix
  signed char *need;    // per (window, dimension): highest state bit the window's members look at
  u128 *total;          // per chain: sum of X over the whole chain
  i128 *B;              // per chain (head)
  uint32_t *head_end;   // per chain (head): first member the segments cover
  signed char *head_je; // per chain (head): trailing zeros of the state there
  uint32_t *q_start;    // per chain (head): first window the chaining pass applies; 0xffffffff: finished by the head
  fx::SegRecord *rec;   // per (window, dimension), chain-major
  unsigned int *refine_list;   // storage indices of the records the first chaining pass marked
  unsigned int *refine_count;
  fx::Tables *tab;      // X_t table (device copy)
};

struct FxGeom {
  const uint8_t *dense;
  unsigned int stride;
  const uint32_t *order;  // null: identity (K == 1)
  int K, dim;
  unsigned int C;         // nominal window length
  unsigned int n;         // local vectors
};

struct FxAcc {
  const uint8_t *dense;
  const uint32_t *order;
  unsigned int stride;
  int e;
  __device__ __forceinline__ int operator()(unsigned int p) const {
    const unsigned int v = order ? __ldg(order + p) : p;
    return (int)(__ldg(dense + (size_t)v * stride + e) ^ 0x80u);  // t = byte ^ 0x80
  }
};

// record / window-array index of (cell k, dimension e, window q): contiguous along a chain
__device__ __forceinline__ size_t fx_index(const FxWork &w, int dim, int k, int e, unsigned int q) {
  const unsigned int w0 = w.win_off[k], cnt = w.win_off[k + 1] - w0;
  return (size_t)w0 * dim + (size_t)e * cnt + q;
}

__global__ void fx_cells_kernel(const uint32_t *__restrict__ keys_sorted, const unsigned int n, const int K, const unsigned int C,
                                FxWork w, unsigned long long *__restrict__ counts) {
  // one block: cell boundaries by binary search, then the exclusive scan of the cells' window counts
  __shared__ unsigned int s_carry;
  __shared__ unsigned int s_part[1024];
  for (int k = threadIdx.x; k <= K; k += blockDim.x) {
    unsigned int lo = 0, hi = n;
    if (keys_sorted) {
      while (lo < hi) {
        const unsigned int mid = lo + ((hi - lo) >> 1);
        if (__ldg(keys_sorted + mid) < (unsigned int)k) lo = mid + 1; else hi = mid;
      }
    } else {
      lo = k == 0 ? 0 : n;
    }
    w.cell_beg[k] = lo;
  }
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < K; base += blockDim.x) {
    const int k = base + threadIdx.x;
    unsigned int cnt = 0;
    if (k < K) {
      const unsigned int members = w.cell_beg[k + 1] - w.cell_beg[k];
      cnt = (members + C - 1) / C;
      if (counts) counts[k] = members;
    }
    s_part[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {  // Hillis-Steele inclusive scan
      const unsigned int add = (int)threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
      __syncthreads();
      s_part[threadIdx.x] += add;
      __syncthreads();
    }
    if (k < K) w.win_off[k] = s_carry + s_part[threadIdx.x] - cnt;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry += s_part[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    w.win_off[K] = s_carry;
    *w.refine_count = 0;
  }
}

// (global window, dimension) -> cell, window inside the cell
__device__ __forceinline__ void fx_locate(const FxWork &w, int K, unsigned int gw, int &k, unsigned int &q) {
  int lo = 0, hi = K;  // last cell with win_off <= gw
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (w.win_off[mid] <= gw) lo = mid; else hi = mid;
  }
  k = lo;
  q = gw - w.win_off[lo];
}

__global__ void __launch_bounds__(256) fx_pre_kernel(const FxGeom g, FxWork w) {
  __shared__ fx::Tables tab;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tab.X[i] = w.tab->X[i];
  __syncthreads();
  const unsigned long long total = (unsigned long long)w.win_off[g.K] * g.dim;
  for (unsigned long long id = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; id < total;
       id += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned int gw = (unsigned int)(id / (unsigned int)g.dim);
    const int e = (int)(id - (unsigned long long)gw * g.dim);
    int k;
    unsigned int q;
    fx_locate(w, g.K, gw, k, q);
    const unsigned int beg = w.cell_beg[k], end = w.cell_beg[k + 1];
    const unsigned int P = beg + q * g.C, P_end = end - P > g.C ? P + g.C : end;
    const FxAcc acc{g.dense, g.order, g.stride, e};
    u128 sx;
    int nd;
    fx::fx_window(acc, tab, P, P_end, sx, nd);
    const size_t idx = fx_index(w, g.dim, k, e, q);
    w.sumX[idx] = sx;
    w.need[idx] = (signed char)nd;
  }
}

// per chain (one WARP each): exclusive prefix of the window sums (in place) and the chain total.  Lanes take 32
// consecutive windows; the 128-bit values are scanned as three 32-bit limbs in 64-bit lanes (no carries inside the
// scan: a limb sum over at most 2^22 windows stays below 2^54) and recombined.
__global__ void __launch_bounds__(128) fx_scan_kernel(const FxGeom g, FxWork w) {
  const int chain = (int)((blockIdx.x * (unsigned int)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (chain >= g.K * g.dim) return;
  const int k = chain / g.dim, e = chain - k * g.dim;
  const unsigned int cnt = w.win_off[k + 1] - w.win_off[k];
  u128 *p = w.sumX + fx_index(w, g.dim, k, e, 0);
  u128 carry = 0;
  for (unsigned int q0 = 0; q0 < cnt; q0 += 128) {  // four consecutive windows per lane
    const unsigned int q = q0 + 4 * lane;
    u128 v[4], loc = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      v[i] = q + i < cnt ? p[q + i] : (u128)0;
      loc += v[i];
    }
    unsigned long long l0 = (unsigned long long)loc & 0xffffffffull, l1 = (unsigned long long)(loc >> 32) & 0xffffffffull,
                       l2 = (unsigned long long)(loc >> 64);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long a0 = __shfl_up_sync(0xffffffffu, l0, o), a1 = __shfl_up_sync(0xffffffffu, l1, o),
                               a2 = __shfl_up_sync(0xffffffffu, l2, o);
      if (lane >= o) {
        l0 += a0;
        l1 += a1;
        l2 += a2;
      }
    }
    const u128 incl = (u128)l0 + ((u128)l1 << 32) + ((u128)l2 << 64);
    u128 run = carry + incl - loc;  // exclusive prefix of this lane's first window
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (q + i < cnt) p[q + i] = run;
      run += v[i];
    }
    const unsigned long long t0 = __shfl_sync(0xffffffffu, l0, 31), t1 = __shfl_sync(0xffffffffu, l1, 31),
                             t2 = __shfl_sync(0xffffffffu, l2, 31);
    carry += (u128)t0 + ((u128)t1 << 32) + ((u128)t2 << 64);
  }
  if (lane == 0) w.total[chain] = carry;
}

__global__ void __launch_bounds__(128) fx_head_kernel(const FxGeom g, FxWork w, double *__restrict__ state) {
  __shared__ fx::Tables tab;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tab.X[i] = w.tab->X[i];
  __syncthreads();
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= g.K * g.dim) return;
  const int k = chain / g.dim, e = chain - k * g.dim;
  const unsigned int beg = w.cell_beg[k], end = w.cell_beg[k + 1];
  const FxAcc acc{g.dense, g.order, g.stride, e};
  const fx::HeadOut h = fx::fx_head(acc, tab, beg, end, g.C, state[2 * (size_t)chain], state[2 * (size_t)chain + 1]);
  w.B[chain] = h.B;
  w.head_end[chain] = h.pos_end;
  w.head_je[chain] = (signed char)h.je;
  w.q_start[chain] = h.done ? 0xffffffffu : h.q_start;
  if (h.done) {
    state[2 * (size_t)chain] = h.sum;
    state[2 * (size_t)chain + 1] = h.c;
  }
}

// Members of one (window, dimension) staged in shared memory, one byte column per thread: every later read of the
// speculative runs (anchor scans, one pass per class) is a shared-memory load instead of two dependent global ones.
constexpr int kFxRunThreads = 128;
struct FxStaged {
  const unsigned char *col;  // this thread's column: member p of the window at col[(p - base) * kFxRunThreads]
  unsigned int base;
  __device__ __forceinline__ int operator()(unsigned int p) const { return (int)col[(p - base) * kFxRunThreads]; }
};

__global__ void __launch_bounds__(kFxRunThreads) fx_runs_kernel(const FxGeom g, FxWork w) {
  extern __shared__ __align__(16) unsigned char fx_smem[];
  fx::Tables &tab = *reinterpret_cast<fx::Tables *>(fx_smem);
  unsigned char *stage = fx_smem + sizeof(fx::Tables);  // [C + kFxAnchorWin][kFxRunThreads] bytes
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tab.X[i] = w.tab->X[i];
  __syncthreads();
  const unsigned long long total = (unsigned long long)w.win_off[g.K] * g.dim;
  const unsigned long long rounds = (total + kFxRunThreads - 1) / kFxRunThreads;
  for (unsigned long long rnd = blockIdx.x; rnd < rounds; rnd += gridDim.x) {
    const unsigned long long id = rnd * kFxRunThreads + threadIdx.x;
    if (id >= total) continue;
    const unsigned int gw = (unsigned int)(id / (unsigned int)g.dim);
    const int e = (int)(id - (unsigned long long)gw * g.dim);
    int k;
    unsigned int q;
    fx_locate(w, g.K, gw, k, q);
    const int chain = k * g.dim + e;
    const unsigned int qs = w.q_start[chain];
    if (q < qs) {  // covered by the head (0xffffffff: the whole chain): an empty record, so that later passes skip it
      fx::SegRecord z;
      z.Eb = 0;
      z.begin = z.end = 0;
      z.w_base = 0;
      z.xs_low = 0;
      z.je = 0;
      z.top = -1;
      z.ncls = 1;
      z.flags = 0;
      w.rec[fx_index(w, g.dim, k, e, q)] = z;
      continue;
    }
    const unsigned int beg = w.cell_beg[k], end = w.cell_beg[k + 1], cnt = w.win_off[k + 1] - w.win_off[k];
    const unsigned int P = beg + q * g.C;
    const unsigned int len = end - P > g.C + fx::kFxAnchorWin ? g.C + fx::kFxAnchorWin : end - P;
    {
      const FxAcc src{g.dense, g.order, g.stride, e};
      unsigned char *col = stage + threadIdx.x;
#pragma unroll 8
      for (unsigned int i = 0; i < len; i++) col[i * kFxRunThreads] = (unsigned char)src(P + i);
    }
    const FxStaged acc{stage + threadIdx.x, P};
    const size_t idx = fx_index(w, g.dim, k, e, q);
    fx::SegRecord r;
    int je;
    u128 before;
    fx::fx_segment_bounds(acc, tab, beg, end, g.C, q, cnt, qs, w.head_end[chain], (int)w.head_je[chain], r.begin, r.end, je, before);
    r.je = (signed char)je;
    const int n0 = w.need[idx], n1 = q + 1 < cnt ? (int)w.need[idx + 1] : -1;
    r.top = (signed char)(n0 > n1 ? n0 : n1);
    r.Eb = (u128)(w.B[chain] + (i128)w.sumX[idx] + (i128)before);
    r.ncls = 0;
    r.flags = 0;
    r.xs_low = 0;
    r.w_base = 0;
    if (r.begin < r.end) fx::fx_run_segment_multi(acc, tab, r);
    w.rec[idx] = r;
  }
}

// Refinement round: the segments the first chaining pass marked (W outside their margin) are re-run around that
// pass's estimate of W, so that the second pass finds them covered.
__global__ void __launch_bounds__(kFxRunThreads) fx_refine_kernel(const FxGeom g, FxWork w) {
  extern __shared__ __align__(16) unsigned char fx_smem[];
  fx::Tables &tab = *reinterpret_cast<fx::Tables *>(fx_smem);
  unsigned char *stage = fx_smem + sizeof(fx::Tables);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tab.X[i] = w.tab->X[i];
  __syncthreads();
  const unsigned int total = *w.refine_count;
  for (unsigned int li = blockIdx.x * kFxRunThreads + threadIdx.x; li < total; li += gridDim.x * kFxRunThreads) {
    // records are chain-major: id is a storage index, the dimension follows from the cell's layout
    const unsigned long long id = w.refine_list[li];
    fx::SegRecord r = w.rec[id];
    const unsigned int gw_first = (unsigned int)(id / (unsigned int)g.dim);  // some window of the same cell
    int k;
    unsigned int q_unused;
    fx_locate(w, g.K, gw_first, k, q_unused);
    const unsigned int cnt = w.win_off[k + 1] - w.win_off[k];
    const int e = (int)((id - (unsigned long long)w.win_off[k] * g.dim) / cnt);
    const unsigned int len = r.end - r.begin;
    if (len > g.C + fx::kFxAnchorWin) {  // cannot be staged (never with the window lengths in use): leave it to the exact pass
      r.flags = 0;
      r.ncls = 0;
      w.rec[id] = r;
      continue;
    }
    const FxAcc src{g.dense, g.order, g.stride, e};
    unsigned char *col = stage + threadIdx.x;
#pragma unroll 8
    for (unsigned int i = 0; i < len; i++) col[i * kFxRunThreads] = (unsigned char)src(r.begin + i);
    const FxStaged acc{stage + threadIdx.x, r.begin};
    fx::fx_run_segment_multi(acc, tab, r);
    r.flags = 0;
    w.rec[id] = r;
  }
}

// Chaining: one WARP per chain, 32 segments at a time, one record per lane.  Where the batch allows it every lane
// turns its record into a 4-state map and the warp composes them with a parallel prefix (fx_compose, five shuffle
// rounds) instead of 32 dependent steps; the first segment whose validity interval does not contain W (or that is
// marked sequential) is re-run exactly - the warp stages its members in shared memory, one lane steps through them -
// and the prefix is redone behind it.  Batches whose class bits do not fit two adjacent positions are applied one
// record after the other (fx_apply).
constexpr int kFxChainWarps = 4;

__device__ __forceinline__ fx::Map4 fx_map_shfl_up(const fx::Map4 &m, int o) {
  fx::Map4 r;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    r.s[i] = __shfl_up_sync(0xffffffffu, m.s[i], o);
    r.dW[i] = __shfl_up_sync(0xffffffffu, m.dW[i], o);
    r.lo[i] = __shfl_up_sync(0xffffffffu, m.lo[i], o);
    r.hi[i] = __shfl_up_sync(0xffffffffu, m.hi[i], o);
  }
  return r;
}
__device__ __forceinline__ int fx_pick(const int (&v)[4], int s) { return s == 0 ? v[0] : s == 1 ? v[1] : s == 2 ? v[2] : v[3]; }

__global__ void __launch_bounds__(32 * kFxChainWarps) fx_chain_kernel(const FxGeom g, FxWork w, double *__restrict__ state, const int pass) {
  __shared__ fx::Tables tab;
  __shared__ unsigned char s_seg[kFxChainWarps][256 + fx::kFxAnchorWin];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tab.X[i] = w.tab->X[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chain = blockIdx.x * kFxChainWarps + warp;
  if (chain >= g.K * g.dim) return;
  const unsigned int qs = w.q_start[chain];
  if (qs == 0xffffffffu) return;  // finished by the head
  const int k = chain / g.dim, e = chain - k * g.dim;
  const unsigned int cnt = w.win_off[k + 1] - w.win_off[k];
  fx::SegRecord *rec = w.rec + fx_index(w, g.dim, k, e, 0);
  const FxAcc src{g.dense, g.order, g.stride, e};
  long long W = 0;  // the head's state is exact: the first segment is entered with W = 0 (all lanes carry W)
  int marked = 0;   // pass 0: segments marked for the refinement round (none: this pass's result is exact)

  // exact re-run of lane f's segment from the true state Eb + W; every lane returns the new W
  auto rerun = [&](const fx::SegRecord &mine, int f) {
    fx::SegRecord seg;
    seg.begin = __shfl_sync(0xffffffffu, mine.begin, f);
    seg.end = __shfl_sync(0xffffffffu, mine.end, f);
    const unsigned long long e_lo = __shfl_sync(0xffffffffu, (unsigned long long)mine.Eb, f);
    const unsigned long long e_hi = __shfl_sync(0xffffffffu, (unsigned long long)(mine.Eb >> 64), f);
    seg.Eb = ((u128)e_hi << 64) | e_lo;
    const unsigned int len = seg.end - seg.begin;
    long long Wn = W;
    if (len <= sizeof(s_seg[0])) {  // the warp stages the members, one lane steps through them
      for (unsigned int i = lane; i < len; i += 32) s_seg[warp][i] = (unsigned char)src(seg.begin + i);
      __syncwarp();
      if (lane == 0) {
        struct Staged {
          const unsigned char *p;
          unsigned int base;
          __device__ __forceinline__ int operator()(unsigned int q) const { return (int)p[q - base]; }
        };
        fx::fx_rerun(Staged{s_seg[warp], seg.begin}, tab, seg, Wn);
      }
      __syncwarp();
    } else if (lane == 0) {
      fx::fx_rerun(src, tab, seg, Wn);
    }
    W = __shfl_sync(0xffffffffu, Wn, 0);
  };

  fx::SegRecord r;
  // pass 0: lane f's segment is not covered - note W in its record for the refinement round and move on with an estimate
  auto mark = [&](int f, unsigned int q0) {
    long long Wn = W;
    if (lane == f) {
      fx::fx_mark_and_estimate(r, Wn);
      if (r.flags & fx::kFxRefine) {
        rec[q0 + f] = r;
        w.refine_list[atomicAdd(w.refine_count, 1u)] = (unsigned int)((rec + q0 + f) - w.rec);
      }
    }
    W = __shfl_sync(0xffffffffu, Wn, f);
    marked = 1;  // (sequential segments count as well: only the exact pass can run them)
  };
  auto load = [&](unsigned int q0) {
    const unsigned int q = q0 + lane;
    if (q < cnt) {
      r = rec[q];
    } else {
      r.begin = r.end = 0;
      r.ncls = 1;
      r.je = 0;
      r.top = -1;
      r.Eb = 0;
    }
  };
  for (unsigned int q0 = qs; q0 < cnt; q0 += 32) {
    load(q0);
    const bool live = r.begin < r.end, usable = live && r.ncls != 0;
    const unsigned int live_mask = __ballot_sync(0xffffffffu, live);
    if (!live_mask) continue;
    const int je_min = __reduce_min_sync(0xffffffffu, usable ? (int)r.je : 99);
    const int top_max = __reduce_max_sync(0xffffffffu, usable ? (int)r.top : -1);
    int jb = 0;
    const bool composable = je_min != 99 && fx::fx_batch_composable(je_min, 0, top_max, jb) && W < fx::kFxWLimit / 2 && W > -fx::kFxWLimit / 2;
    if (!composable) {  // one record after the other
      unsigned int todo = live_mask;
      while (todo) {
        const int f = __ffs(todo) - 1;
        todo &= todo - 1;
        long long Wn = W;
        int ok = 1;
        if (lane == f) ok = fx::fx_apply(r, Wn) ? 1 : 0;
        ok = __shfl_sync(0xffffffffu, ok, f);
        if (ok) {
          W = __shfl_sync(0xffffffffu, Wn, f);
        } else if (pass == 0) {
          mark(f, q0);
        } else {
          rerun(r, f);
        }
      }
      continue;
    }
    const fx::Map4 mine = fx::fx_map_of(r, jb);  // identity for empty records
    const unsigned int e16 = (unsigned int)((unsigned long long)r.Eb & 0xffffu);
    int j0 = 0;
    while (true) {
      const unsigned int rest = live_mask & (j0 >= 32 ? 0u : (0xffffffffu << j0));
      if (!rest) break;
      const int jl = __ffs(rest) - 1;  // first live record at or behind j0: its E gives the entry state
      const int s0 = (int)((((long long)__shfl_sync(0xffffffffu, e16, jl) + (W - (long long)__shfl_sync(0xffffffffu, r.w_base, jl))) >> jb) & 3);
      fx::Map4 pm = lane >= j0 ? mine : fx::fx_map_identity();
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const fx::Map4 left = fx_map_shfl_up(pm, o);
        if (lane >= o) pm = fx::fx_compose(left, pm);
      }
      const int lo = fx_pick(pm.lo, s0), hi = fx_pick(pm.hi, s0), dW = fx_pick(pm.dW, s0);
      const bool valid = lo <= hi && W >= lo && W <= hi;
      const unsigned int bad = __ballot_sync(0xffffffffu, !valid) & (j0 >= 32 ? 0u : (0xffffffffu << j0));
      if (!bad) {
        W += __shfl_sync(0xffffffffu, dW, 31);
        break;
      }
      const int f = __ffs(bad) - 1;  // records j0 .. f-1 are covered; f is re-run exactly
      if (f > j0) W += __shfl_sync(0xffffffffu, dW, f - 1);
      if ((live_mask >> f) & 1u) {
        if (pass == 0)
          mark(f, q0);
        else
          rerun(r, f);
      }
      j0 = f + 1;
    }
  }
  if (pass == 0 && marked) return;  // the exact pass redoes this chain after the refinement round
  if (lane == 0) {
    const u128 A = (u128)(w.B[chain] + (i128)w.total[chain] + (i128)W);
    double sum, c;
    fx::fx_state_to_pair(A, sum, c);
    state[2 * (size_t)chain] = sum;
    state[2 * (size_t)chain + 1] = c;
    if (pass == 0) w.q_start[chain] = 0xffffffffu;  // nothing was marked: done, the exact pass skips this chain
  }
}

// Long chains (few cells): one BLOCK per chain, kFxBlockWarps warps x 32 segments per round.  Every warp turns its 32
// records into maps and scans them as above; the warps' totals are exchanged through shared memory, each warp
// composes the totals in front of it to get its own entry (state, W) and checks its lanes; the first segment that is
// not covered is handled by its warp (marked in pass 0, re-run in pass 1) and the round is re-evaluated behind it
// (the per-record maps stay in registers, only one warp's scan and the small cross-warp composition are redone).
constexpr int kFxBlockWarps = 8;

__global__ void __launch_bounds__(32 * kFxBlockWarps) fx_chain_block_kernel(const FxGeom g, FxWork w, double *__restrict__ state,
                                                                            const int pass) {
  __shared__ fx::Tables tab;
  __shared__ unsigned char s_seg[256 + fx::kFxAnchorWin];
  __shared__ fx::Map4 s_tot[kFxBlockWarps];
  __shared__ int s_fail[kFxBlockWarps], s_je[kFxBlockWarps], s_top[kFxBlockWarps], s_first[kFxBlockWarps];
  __shared__ unsigned int s_live[kFxBlockWarps];
  __shared__ long long s_erel[kFxBlockWarps], s_W;
  __shared__ int s_marked;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tab.X[i] = w.tab->X[i];
  if (threadIdx.x == 0) s_marked = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chain = blockIdx.x;
  const unsigned int qs = w.q_start[chain];
  if (qs == 0xffffffffu) return;  // finished by the head or by pass 0 (uniform for the block)
  const int k = chain / g.dim, e = chain - k * g.dim;
  const unsigned int cnt = w.win_off[k + 1] - w.win_off[k];
  fx::SegRecord *rec = w.rec + fx_index(w, g.dim, k, e, 0);
  const FxAcc src{g.dense, g.order, g.stride, e};
  long long W = 0;
  fx::SegRecord r = {};
  unsigned int my_q0 = 0;

  auto rerun = [&](int f) {  // executed by ONE warp (all its lanes): exact re-run of lane f's segment from Eb + W
    fx::SegRecord seg;
    seg.begin = __shfl_sync(0xffffffffu, r.begin, f);
    seg.end = __shfl_sync(0xffffffffu, r.end, f);
    seg.w_base = __shfl_sync(0xffffffffu, r.w_base, f);
    const unsigned long long e_lo = __shfl_sync(0xffffffffu, (unsigned long long)r.Eb, f);
    const unsigned long long e_hi = __shfl_sync(0xffffffffu, (unsigned long long)(r.Eb >> 64), f);
    seg.Eb = ((u128)e_hi << 64) | e_lo;
    const unsigned int len = seg.end - seg.begin;
    long long Wn = W;
    if (len <= sizeof(s_seg)) {
      for (unsigned int i = lane; i < len; i += 32) s_seg[i] = (unsigned char)src(seg.begin + i);
      __syncwarp();
      if (lane == 0) {
        struct Staged {
          const unsigned char *p;
          unsigned int base;
          __device__ __forceinline__ int operator()(unsigned int q) const { return (int)p[q - base]; }
        };
        fx::fx_rerun(Staged{s_seg, seg.begin}, tab, seg, Wn);
      }
      __syncwarp();
    } else if (lane == 0) {
      fx::fx_rerun(src, tab, seg, Wn);
    }
    W = __shfl_sync(0xffffffffu, Wn, 0);
  };
  auto mark = [&](int f) {  // pass 0, one warp: note W in lane f's record for the refinement round, move on with an estimate
    long long Wn = W;
    if (lane == f) {
      fx::fx_mark_and_estimate(r, Wn);
      if (r.flags & fx::kFxRefine) {
        rec[my_q0 + f] = r;
        w.refine_list[atomicAdd(w.refine_count, 1u)] = (unsigned int)((rec + my_q0 + f) - w.rec);
      }
      s_marked = 1;
    }
    W = __shfl_sync(0xffffffffu, Wn, f);
  };

  for (unsigned int q0 = qs; q0 < cnt; q0 += 32 * kFxBlockWarps) {
    my_q0 = q0 + 32 * warp;
    const unsigned int q = my_q0 + lane;
    if (q < cnt) {
      r = rec[q];
    } else {
      r.begin = r.end = 0;
      r.ncls = 1;
      r.je = 0;
      r.top = -1;
      r.Eb = 0;
      r.w_base = 0;
      r.flags = 0;
    }
    const bool live = r.begin < r.end, usable = live && r.ncls != 0;
    const unsigned int live_mask = __ballot_sync(0xffffffffu, live);
    const int je_w = __reduce_min_sync(0xffffffffu, usable ? (int)r.je : 99);
    const int top_w = __reduce_max_sync(0xffffffffu, usable ? (int)r.top : -1);
    if (lane == 0) {
      s_je[warp] = je_w;
      s_top[warp] = top_w;
      s_live[warp] = live_mask;
      if (warp == 0) s_W = W;
    }
    __syncthreads();
    int je_min = 99, top_max = -1;
    unsigned int any_live = 0;
#pragma unroll
    for (int t = 0; t < kFxBlockWarps; t++) {
      je_min = min(je_min, s_je[t]);
      top_max = max(top_max, s_top[t]);
      any_live |= s_live[t];
    }
    int jb = 0;
    const bool composable = je_min != 99 && fx::fx_batch_composable(je_min, 0, top_max, jb) && W < fx::kFxWLimit / 2 && W > -fx::kFxWLimit / 2;
    if (!any_live) {
      __syncthreads();
      continue;
    }
    if (!composable) {  // the warps take turns, one record after the other
      for (int t = 0; t < kFxBlockWarps; t++) {
        if (warp == t) {
          unsigned int todo = live_mask;
          while (todo) {
            const int f = __ffs(todo) - 1;
            todo &= todo - 1;
            long long Wn = W;
            int ok = 1;
            if (lane == f) ok = fx::fx_apply(r, Wn) ? 1 : 0;
            ok = __shfl_sync(0xffffffffu, ok, f);
            if (ok)
              W = __shfl_sync(0xffffffffu, Wn, f);
            else if (pass == 0)
              mark(f);
            else
              rerun(f);
          }
          if (lane == 0) s_W = W;
        }
        __syncthreads();
        W = s_W;
      }
      __syncthreads();
      continue;
    }
    const fx::Map4 mine = fx::fx_map_of(r, jb);
    const long long erel = (long long)((unsigned long long)r.Eb & 0xffffu) - r.w_base;
    int ws = 0, ls = 0;  // the round is evaluated from (warp ws, lane ls) on
    while (ws < kFxBlockWarps) {
      const bool active = warp > ws || (warp == ws && lane >= ls);
      fx::Map4 pm = active ? mine : fx::fx_map_identity();
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const fx::Map4 left = fx_map_shfl_up(pm, o);
        if (lane >= o) pm = fx::fx_compose(left, pm);
      }
      const unsigned int act_mask = warp > ws ? 0xffffffffu : (warp == ws ? (ls >= 32 ? 0u : (0xffffffffu << ls)) : 0u);
      const unsigned int rest = live_mask & act_mask;
      const int fl = rest ? __ffs(rest) - 1 : -1;
      const long long erel_first = __shfl_sync(0xffffffffu, erel, fl < 0 ? 0 : fl);
      if (lane == 31) s_tot[warp] = pm;
      if (lane == 0) {
        s_first[warp] = fl;
        s_erel[warp] = erel_first;
      }
      __syncthreads();
      int fwl = -1;
#pragma unroll
      for (int t = kFxBlockWarps - 1; t >= 0; t--)
        if (t >= ws && s_first[t] >= 0) fwl = t;
      if (fwl < 0) {  // nothing live behind the start: the round is done
        __syncthreads();
        break;
      }
      const int s0 = (int)(((s_erel[fwl] + W) >> jb) & 3);
      fx::Map4 P = fx::fx_map_identity();
      for (int t = ws; t < warp; t++) P = fx::fx_compose(P, s_tot[t]);
      const int p_lo = fx_pick(P.lo, s0), p_hi = fx_pick(P.hi, s0);
      const bool ok_prev = p_lo <= p_hi && W >= p_lo && W <= p_hi;
      const int s_w = fx_pick(P.s, s0);
      const long long W_w = W + fx_pick(P.dW, s0);
      const int lo = fx_pick(pm.lo, s_w), hi = fx_pick(pm.hi, s_w), dW = fx_pick(pm.dW, s_w);
      const bool valid = ok_prev && lo <= hi && W_w >= lo && W_w <= hi;
      const unsigned int bad = __ballot_sync(0xffffffffu, !valid) & act_mask;
      if (lane == 0) s_fail[warp] = bad ? __ffs(bad) - 1 : 32;
      __syncthreads();
      int fw = -1;
#pragma unroll
      for (int t = kFxBlockWarps - 1; t >= 0; t--)
        if (t >= ws && s_fail[t] < 32) fw = t;
      if (fw < 0) {  // everything behind the start is covered
        fx::Map4 T = fx::fx_map_identity();
        for (int t = ws; t < kFxBlockWarps; t++) T = fx::fx_compose(T, s_tot[t]);
        W += fx_pick(T.dW, s0);
        __syncthreads();
        break;
      }
      const int f = s_fail[fw];
      if (warp == fw) {
        W = W_w + (f > 0 ? (long long)__shfl_sync(0xffffffffu, dW, f - 1) : 0ll);  // masked lanes are identities: dW = 0
        if ((live_mask >> f) & 1u) {
          if (pass == 0)
            mark(f);
          else
            rerun(f);
        }
        if (lane == 0) s_W = W;
      }
      __syncthreads();
      W = s_W;
      ws = fw;
      ls = f + 1;
      if (ls == 32) {
        ws++;
        ls = 0;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (pass == 0 && s_marked) return;  // the exact pass redoes this chain after the refinement round
  if (threadIdx.x == 0) {
    const u128 A = (u128)(w.B[chain] + (i128)w.total[chain] + (i128)W);
    double sum, c;
    fx::fx_state_to_pair(A, sum, c);
    state[2 * (size_t)chain] = sum;
    state[2 * (size_t)chain + 1] = c;
    if (pass == 0) w.q_start[chain] = 0xffffffffu;
  }
}

size_t fx_up(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

unsigned int exact_fast_window(size_t n, int K) {
  // Short windows: a re-run costs one window's members sequentially, the chaining pass one cheap step per window.
  (void)n;
  (void)K;
  return 128;
}

size_t exact_fast_workspace_bytes(size_t n, int K, int dim) {
  const unsigned int C = 128;  // the smallest window gives the most windows
  (void)exact_fast_window;
  const size_t windows = n / C + (size_t)K + 1, per = windows * (size_t)dim, chains = (size_t)K * dim;
  size_t b = 0;
  b += fx_up(((size_t)K + 1) * 4) * 2;
  b += fx_up(per * 16) + fx_up(per) + fx_up(chains * 16) * 2 + fx_up(chains * 4) * 2 + fx_up(chains);
  b += fx_up(per * sizeof(fx::SegRecord)) + fx_up(sizeof(fx::Tables)) + fx_up(per * 4) + 256;
  return b + 256;
}

// Same contract as launch_kahan_sums for SCALED lattice sources: state holds K*dim pairs {sum, c}, read as the
// chains' incoming state and overwritten with their final one; counts (may be null) receives the members per cell.
cudaError_t launch_kahan_sums_fast(const VecSource &src, const uint32_t *keys_sorted, const uint32_t *order, int K, double *state,
                                   unsigned long long *counts, void *workspace, size_t workspace_bytes, int sm_count,
                                   cudaStream_t stream) {
  if (!src.dense || src.f64) return cudaErrorInvalidValue;
  const size_t n = (size_t)src.n_local;
  const int dim = src.dim;
  if (workspace_bytes < exact_fast_workspace_bytes(n, K, dim)) return cudaErrorInvalidValue;
  const unsigned int C = exact_fast_window(n, K);
  const size_t windows = n / 128 + (size_t)K + 1, per = windows * (size_t)dim, chains = (size_t)K * dim;
  char *p = (char *)workspace;
  FxWork w;
  auto take = [&](size_t bytes) { char *r = p; p += fx_up(bytes); return r; };
  w.cell_beg = (uint32_t *)take(((size_t)K + 1) * 4);
  w.win_off = (uint32_t *)take(((size_t)K + 1) * 4);
  w.sumX = (u128 *)take(per * 16);
  w.need = (signed char *)take(per);
  w.total = (u128 *)take(chains * 16);
  w.B = (i128 *)take(chains * 16);
  w.head_end = (uint32_t *)take(chains * 4);
  w.head_je = (signed char *)take(chains);
  w.q_start = (uint32_t *)take(chains * 4);
  w.rec = (fx::SegRecord *)take(per * sizeof(fx::SegRecord));
  w.refine_list = (unsigned int *)take(per * 4);
  w.refine_count = (unsigned int *)take(256);
  w.tab = (fx::Tables *)take(sizeof(fx::Tables));
  static fx::Tables h_tab;
  static bool h_tab_ready = false;
  if (!h_tab_ready) {
    fx::fx_fill_tables(h_tab);
    h_tab_ready = true;
  }
  cudaError_t e = cudaMemcpyAsync(w.tab, &h_tab, sizeof h_tab, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  FxGeom g{src.dense, src.dense_stride, K > 1 ? order : nullptr, K, dim, C, (unsigned int)n};
  fx_cells_kernel<<<1, 1024, 0, stream>>>(K > 1 ? keys_sorted : nullptr, (unsigned int)n, K, C, w, counts);
  count_launch();
  const unsigned long long ids = (unsigned long long)(n / C + (size_t)K + 1) * dim;  // upper bound of (window, dimension) pairs
  unsigned long long blocks = (ids + 255) / 256;
  const unsigned long long cap = (unsigned long long)sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  const unsigned int chain_blocks = (unsigned int)((chains + 127) / 128);
  fx_pre_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(g, w);
  count_launch();
  fx_scan_kernel<<<(unsigned int)((chains * 32 + 127) / 128), 128, 0, stream>>>(g, w);
  count_launch();
  fx_head_kernel<<<chain_blocks, 128, 0, stream>>>(g, w, state);
  count_launch();
  unsigned long long rblocks_saved = 1;
  size_t smem_saved = 0;
  {
    const size_t smem = sizeof(fx::Tables) + (size_t)(C + fx::kFxAnchorWin) * kFxRunThreads;
    unsigned long long rblocks = (ids + kFxRunThreads - 1) / kFxRunThreads;
    const unsigned long long rcap = (unsigned long long)sm_count * 8;
    if (rblocks > rcap) rblocks = rcap;
    fx_runs_kernel<<<(unsigned int)rblocks, kFxRunThreads, smem, stream>>>(g, w);
    count_launch();
    rblocks_saved = rblocks;
    smem_saved = smem;
  }
  // long chains (few cells): one block of kFxBlockWarps warps per chain; short ones: one warp per chain
  const bool long_chains = n / (size_t)K >= (size_t)C * 32 * kFxBlockWarps && chains <= 65535;
  const unsigned int cblocks = (unsigned int)((chains + kFxChainWarps - 1) / kFxChainWarps);
  auto chain_pass = [&](int pass) {
    if (long_chains)
      fx_chain_block_kernel<<<(unsigned int)chains, 32 * kFxBlockWarps, 0, stream>>>(g, w, state, pass);
    else
      fx_chain_kernel<<<cblocks, 32 * kFxChainWarps, 0, stream>>>(g, w, state, pass);
    count_launch();
  };
  chain_pass(0);  // marks what its summaries do not cover
  fx_refine_kernel<<<(unsigned int)rblocks_saved, kFxRunThreads, smem_saved, stream>>>(g, w);
  count_launch();
  chain_pass(1);  // exact
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// small cells: the reference's sums of cells with at most kSmallCellMax members, next to the integer sums
// ------------------------------------------------------------------------------------------------
// The integer-sum centroids (S/255)/n differ from the reference's compensated sums in the last bits, which the
// auto centroid mode has to assume for every codevector it cannot prove exact (qb200_train).  Cells with a handful
// of members are where exact distance ties are STRUCTURAL rather than accidental (a 2-member cell of two equal-norm
// vectors is split into 1.2c / 0.8c, both equally far from either member), and they are also the cheap ones: their
// members are found by one pass over the assignment and their sums are the literal loop of
// Solution::sumInArea (/root/reference/src/Quantizer.cpp:59-70) over <= 8 values.  Sharded runs: a rank's members of
// a cell are contiguous in the cell's (ascending vector index) order, behind those of the lower ranks, so every rank
// writes its members' bytes at their final positions of a zeroed table and the table is summed over the ranks -
// the same all-reduce as the statistics.  How many members each rank holds travels as one nibble per rank in a
// 64-bit word per cell (<= 16 ranks, <= 8 members).
namespace {

__global__ void __launch_bounds__(256)
    small_count_kernel(const unsigned long long *__restrict__ stats_local, const int K, const int row_words, const int rank,
                       unsigned long long *__restrict__ packed) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const unsigned long long n = stats_local[(size_t)k * row_words];
  packed[k] = (n < 15ull ? n : 15ull) << (4 * rank);  // 9..15: "too many" (the reduced count decides anyway)
}

// flag[k] = the cell takes this path; off[k] = members of the cell on lower ranks.  Also clears the cell's member
// counter and its rows of the exchange table, and raises *any (this level's word) when at least one cell is small -
// the member scan returns at once otherwise.  Rows of cells that are not small keep stale words: never read.
__global__ void __launch_bounds__(256)
    small_mark_kernel(const unsigned long long *__restrict__ stats, const unsigned long long *__restrict__ packed, const int K,
                      const int row_words, const int rank, const int words, unsigned char *__restrict__ flag,
                      unsigned char *__restrict__ off, unsigned int *__restrict__ cnt, unsigned long long *__restrict__ table,
                      unsigned int *__restrict__ any) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const unsigned long long n = stats[(size_t)k * row_words];
  const bool small = n >= 1 && n <= (unsigned long long)kSmallCellMax;
  flag[k] = small ? 1 : 0;
  cnt[k] = 0;
  unsigned int before = 0;
  if (small) {
    if (packed) {
      const unsigned long long w = packed[k];
      for (int q = 0; q < rank; q++) before += (unsigned int)(w >> (4 * q)) & 15u;
    }
    for (int i = 0; i < kSmallCellMax * words; i++) table[(size_t)k * kSmallCellMax * words + i] = 0;
    *any = 1u;  // same value from every writer; cleared once per train (launch_small_cells_reset)
  }
  off[k] = (unsigned char)before;
}

__global__ void __launch_bounds__(256)
    small_collect_kernel(const uint32_t *__restrict__ assign, const unsigned int n, const unsigned char *__restrict__ flag,
                         unsigned int *__restrict__ cnt, uint32_t *__restrict__ list, const unsigned int *__restrict__ any) {
  if (!*any) return;
  for (unsigned int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    const uint32_t k = __ldg(assign + v) & 0x7fffffffu;
    if (!flag[k]) continue;
    const unsigned int pos = atomicAdd(cnt + k, 1u);
    if (pos < (unsigned int)kSmallCellMax) list[(size_t)k * kSmallCellMax + pos] = v;  // (always: local count <= reduced count)
  }
}

// one thread per cell: its local members in ascending order, their bytes t = L + 128 packed eight to a word
__global__ void __launch_bounds__(128)
    small_pack_kernel(const VecSource src, const int K, const int words, const unsigned char *__restrict__ flag,
                      const unsigned char *__restrict__ off, const unsigned int *__restrict__ cnt, const uint32_t *__restrict__ list,
                      unsigned long long *__restrict__ table, const unsigned int *__restrict__ any) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (!*any || k >= K || !flag[k]) return;
  unsigned int c = cnt[k];
  if (c > (unsigned int)kSmallCellMax) c = kSmallCellMax;
  uint32_t idx[kSmallCellMax];
#pragma unroll
  for (int i = 0; i < kSmallCellMax; i++) idx[i] = i < (int)c ? list[(size_t)k * kSmallCellMax + i] : 0xffffffffu;
#pragma unroll
  for (int i = 1; i < kSmallCellMax; i++)  // insertion sort with compile-time bounds: idx stays in registers
#pragma unroll
    for (int j = i; j > 0; j--) {
      const uint32_t a = idx[j - 1], b = idx[j];
      idx[j - 1] = a < b ? a : b;
      idx[j] = a < b ? b : a;
    }
  const int dim = src.dim;
#pragma unroll
  for (int i = 0; i < kSmallCellMax; i++) {
    if (i >= (int)c) break;
    const unsigned int slot = (unsigned int)off[k] + i;
    if (slot >= (unsigned int)kSmallCellMax) break;
    const uint8_t *row = src.dense + (unsigned long long)idx[i] * src.dense_stride;
    for (int w = 0; w < words; w++) {
      unsigned long long word = 0;
      for (int b = 0; b < 8 && 8 * w + b < dim; b++) word |= (unsigned long long)(uint8_t)(row[8 * w + b] ^ 0x80u) << (8 * b);
      table[((size_t)k * kSmallCellMax + slot) * words + w] = word;
    }
  }
}

// one thread per (cell, dimension): src/Quantizer.cpp:64-67 over the cell's members in vector order
__global__ void __launch_bounds__(256)
    small_sums_kernel(const unsigned long long *__restrict__ stats, const int K, const int dim, const int row_words, const int words,
                      const unsigned char *__restrict__ flag, const unsigned long long *__restrict__ table, double *__restrict__ sums) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)K * dim) return;
  const int k = (int)(i / dim), e = (int)(i - (size_t)k * dim);
  if (!flag[k]) return;
  const int n = (int)stats[(size_t)k * row_words];
  double sum = 0.0, c = 0.0;
  for (int m = 0; m < n; m++) {
    const unsigned int t = (unsigned int)(table[((size_t)k * kSmallCellMax + m) * words + (e >> 3)] >> (8 * (e & 7))) & 255u;
    const double x = __ddiv_rn((double)t, 255.0);  // ScaledColor: (L + 128) / 255 (src/ColorSpace.cpp:16-21)
    const double y = __dsub_rn(x, c);
    const double s = __dadd_rn(sum, y);
    c = __dsub_rn(__dsub_rn(s, sum), y);
    sum = s;
  }
  sums[i] = sum;
}

struct SmallWork {
  unsigned int *any;  // one word per split level (cleared once per train): some cell of that level is small
  unsigned int *cnt;
  uint32_t *list;
  unsigned char *flag, *off;
  unsigned long long *table;
  double *sums;
};
inline size_t small_up(size_t x) { return (x + 255) & ~(size_t)255; }
SmallWork small_carve(void *ws, int K, int dim) {
  const size_t words = (size_t)(dim + 7) / 8;
  char *p = (char *)ws;
  SmallWork w;
  w.any = (unsigned int *)p;
  p += 256;
  w.cnt = (unsigned int *)p;
  p += small_up((size_t)K * 4);
  w.list = (uint32_t *)p;
  p += small_up((size_t)K * kSmallCellMax * 4);
  w.flag = (unsigned char *)p;
  p += small_up((size_t)K);
  w.off = (unsigned char *)p;
  p += small_up((size_t)K);
  w.table = (unsigned long long *)p;
  p += small_up((size_t)K * kSmallCellMax * words * 8);
  w.sums = (double *)p;
  return w;
}

}  // namespace

size_t small_cells_table_words(int K, int dim) { return (size_t)K * kSmallCellMax * (size_t)((dim + 7) / 8); }
size_t small_cells_workspace_bytes(int K, int dim) {
  return small_up((size_t)K * 4) + small_up((size_t)K * kSmallCellMax * 4) + 2 * small_up((size_t)K) +
         small_up(small_cells_table_words(K, dim) * 8) + small_up((size_t)K * dim * 8) + 512;
}
unsigned long long *small_cells_table(void *ws, int K, int dim) { return small_carve(ws, K, dim).table; }
const unsigned char *small_cells_flags(void *ws, int K, int dim) { return small_carve(ws, K, dim).flag; }
const double *small_cells_sums(void *ws, int K, int dim) { return small_carve(ws, K, dim).sums; }

// once per train, before the first level
cudaError_t launch_small_cells_reset(void *ws, cudaStream_t stream) { return cudaMemsetAsync(ws, 0, 256, stream); }

// before the statistics are reduced: this rank's member counts, one nibble per rank (packed: K words, summed with the statistics)
cudaError_t launch_small_cells_count(const unsigned long long *stats_local, int K, int dim, int rank, unsigned long long *packed,
                                     cudaStream_t stream) {
  small_count_kernel<<<(K + 255) / 256, 256, 0, stream>>>(stats_local, K, dim + 2, rank, packed);
  count_launch();
  return cudaGetLastError();
}

// after the reduction: mark the small cells and write this rank's members into the (zeroed) table
cudaError_t launch_small_cells_collect(const VecSource &src, const uint32_t *assign, const unsigned long long *stats,
                                       const unsigned long long *packed, int K, int rank, int level, void *ws, int sm_count,
                                       cudaStream_t stream) {
  const int dim = src.dim, words = (dim + 7) / 8;
  const SmallWork w = small_carve(ws, K, dim);
  unsigned int *any = w.any + (level & 63);
  small_mark_kernel<<<(K + 255) / 256, 256, 0, stream>>>(stats, packed, K, dim + 2, rank, words, w.flag, w.off, w.cnt, w.table, any);
  count_launch();
  const unsigned int n = (unsigned int)src.n_local;
  if (n) {
    unsigned int blocks = (n + 255) / 256;
    const unsigned int cap = (unsigned int)sm_count * 8;
    if (blocks > cap) blocks = cap;
    small_collect_kernel<<<blocks, 256, 0, stream>>>(assign, n, w.flag, w.cnt, w.list, any);
    count_launch();
    small_pack_kernel<<<(K + 127) / 128, 128, 0, stream>>>(src, K, words, w.flag, w.off, w.cnt, w.list, w.table, any);
    count_launch();
  }
  return cudaGetLastError();
}

// after the table is reduced: the compensated sums of the small cells (small_cells_sums; cells not flagged are not written)
cudaError_t launch_small_cells_sums(const unsigned long long *stats, int K, int dim, void *ws, cudaStream_t stream) {
  const SmallWork w = small_carve(ws, K, dim);
  const size_t total = (size_t)K * dim;
  small_sums_kernel<<<(unsigned int)((total + 255) / 256), 256, 0, stream>>>(stats, K, dim, dim + 2, (dim + 7) / 8, w.flag, w.table, w.sums);
  count_launch();
  return cudaGetLastError();
}

}  // namespace qb
