// Hand-written sm_100a kernels of the LBG hot path and their launchers (the tensor-core filter lives in
// qb200_assign_tc.cu).
//
//   stage_codebook_kernel   FP64 codebook -> FP32 rows, bf16 limb tiles, transposed FP64 copy, max|C|
//   assign_kernel           Solution::assignCodeVectors   (/root/reference/src/Quantizer.cpp:24-32)
//                           brute-force FP32 (FFMA2) nearest-codevector FILTER, 16 < K < 128 (and whenever the tensor-core engine is off)
//   small_k_fused_kernel    the same filter + per-cell statistics in one pass for the first split levels (K <= 16)
//   resolve_bruteforce_kernel / resolve_kernel
//                           KDTree::nearestNeighbour      (/root/reference/src/KDTree.cpp:20-29 ->
//                           nanoflann.hpp:906-920,1188-1270,320-345) exact FP64 re-solve of the queries
//                           the filter could not decide: brute force, then the reference's tree walk for ties
//   accumulate_*            fixCodeVectors / updateDistortion as integer per-cell statistics
//                           (/root/reference/src/Quantizer.cpp:9-22,59-87)
//   finalize_split_kernel   centroids, distortions and the next split (src/Quantizer.cpp:81-85,134-138) on the device
//   pick_members_kernel / fetch_members_kernel   empty-cell repair (extension)
//   decode_kernel           decompress + getImageFromVectors (/root/reference/src/Compressor.cpp:64-92,156-165)
//
// Why a filter + resolver: the reference computes everything in FP64 (include/VectorOperations.hpp:10)
// and breaks exact ties by KD-tree traversal order.  Reduced precision cannot reproduce that bit-for-bit, so
// a filter pass only decides queries whose best and second-best scores differ by more than a rigorous
// bound on its own rounding error; all others are appended to a list and re-solved with the reference's
// arithmetic and tie order.
#include "qb200_launch.hpp"
#include "qb200_ptx.cuh"

#include <cfloat>
#include <cstdio>

namespace qb {

// ------------------------------------------------------------------------------------------------
// assign_kernel: FP32 filter
// ------------------------------------------------------------------------------------------------
// Lattice coordinates: every training-vector element is the integer L = (int8)byte in [-128,127]
// (SCALED value = (L+128)/255, NORMAL value = L), codevectors are mapped to the same axis
// (C = 255*c - 128 resp. C = c) on the host in FP64 and rounded once to FP32.  The score
//     s_k = |C_k|^2 - 2 <X, C_k>   ( = |X - C_k|^2 - |X|^2 )
// is one FFMA per dimension: row k of the staged codebook is [-2*C_k[0..DIM), |C_k|^2, pad].
// |s_k(fp32) - s_k(exact, FP64 codebook)| <= (DIM+3) * 2^-24 * (|X| + max_k|C_k|)^2, so a query is
// decided here only when  second - best > 2 * that bound (margin_coef carries the constant).
template <int DIM>
struct AssignCfg {
  static constexpr int ROW = ((DIM + 1 + 3) / 4) * 4;  // floats per staged codebook row
  static constexpr int Q = DIM <= 12 ? 4 : 2;          // queries per thread (even: processed as pairs)
  static constexpr int THREADS = DIM <= 27 ? 512 : 256;
};

// Inner loop, per thread and per PAIR of codevectors (k, k+1), for each pair of its queries (q0, q1):
//   (s_k[q0], s_k[q1])     = FFMA2 chain over the dimensions, addend initialised with |C_k|^2
//   (s_k+1[q0], s_k+1[q1]) = same with row k+1
// and per query, with lo/hi = min/max(s_k, s_k+1):
//   second = min3(second, hi, max(lo, best));  pair = lo < best ? k : pair;  best = min(best, lo)
// i.e. DIM/2 FFMA2 (fma pipe) + 3.5 min/max/select (alu pipe) per distance evaluation.  Which of
// the two codevectors of the winning pair it was is recovered after the loop by recomputing the
// two scores with the identical FMA sequence (bit-identical), so the loop carries no per-codevector
// index bookkeeping.  The staged codebook always holds an even number of rows (the host pads an odd
// K with a row that can never win).
template <int DIM, bool F64 = false>  // F64: general FP64 vectors, scored after rounding to FP32 (qb200_generic.cu)
__global__ void __launch_bounds__(AssignCfg<DIM>::THREADS, 1)
    assign_kernel(const VecSource src, const float *__restrict__ cb_rows, const int K, const int k_chunk,
                  const float margin_coef, const float *__restrict__ c_max_ptr, uint32_t *__restrict__ assign,
                  uint32_t *__restrict__ flag_list, unsigned int *__restrict__ flag_count,
                  const unsigned long long tiles, unsigned long long *__restrict__ stats, const int k_real) {
  using Cfg = AssignCfg<DIM>;
  const float c_max_norm = *c_max_ptr;
  constexpr int ROW = Cfg::ROW, Q = Cfg::Q, THREADS = Cfg::THREADS;
  static_assert(Q % 2 == 0, "queries are processed in pairs");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *s_cb = reinterpret_cast<float *>(smem_raw);
  __shared__ __align__(8) uint64_t s_bar;
  // Fused per-cell statistics (stats != null): the queries this kernel DECIDES are accumulated here into a
  // per-CTA table behind the codebook chunk; flagged queries are added by the resolver once it has decided them.
  unsigned long long *s_q = reinterpret_cast<unsigned long long *>(smem_raw + (((size_t)k_chunk * ROW * 4 + 127) & ~(size_t)127));
  int *s_n = reinterpret_cast<int *>(s_q + k_real);
  int *s_s = s_n + k_real;
  if (stats) {
    for (int i = threadIdx.x; i < k_real; i += THREADS) {
      s_q[i] = 0;
      s_n[i] = 0;
    }
    for (int i = threadIdx.x; i < k_real * DIM; i += THREADS) s_s[i] = 0;
  }  // made visible by the __syncthreads() that follows the barrier initialisation below

  const int tid = threadIdx.x;
  const int n_chunks = (K + k_chunk - 1) / k_chunk;
  uint32_t phase = 0;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // stage a codebook chunk with the TMA unit; one elected thread issues, everyone waits on the mbarrier
  auto stage_chunk = [&](int chunk) {
    const int k0 = chunk * k_chunk;
    const int kn = min(k_chunk, K - k0);
    if (tid == 0) {
      const uint32_t total = (uint32_t)kn * ROW * 4u;
      mbar_expect_tx(&s_bar, total);
      const char *g = reinterpret_cast<const char *>(cb_rows + (size_t)k0 * ROW);
      char *s = reinterpret_cast<char *>(s_cb);
      for (uint32_t off = 0; off < total; off += 32768u) {
        uint32_t n = min(32768u, total - off);
        tma_load_1d(s + off, g + off, n, &s_bar);
      }
    }
    mbar_wait(&s_bar, phase);
    phase ^= 1;
    return kn;
  };

  int staged = -1;
  if (n_chunks == 1) {
    stage_chunk(0);
    staged = 0;
  }

  for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const unsigned long long v0 = tile * (unsigned long long)(THREADS * Q);
    unsigned long long xp[Q / 2][DIM];  // (x[q0][e], x[q1][e]) pairs
    float xn[Q];
    bool live[Q];
#pragma unroll
    for (int qp = 0; qp < Q / 2; qp++) {
      float x[2][DIM];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int q = 2 * qp + h;
        const unsigned long long v = v0 + (unsigned long long)q * THREADS + tid;
        live[q] = v < src.n_local;
        xn[q] = 0.f;
        if (live[q]) {
          if constexpr (F64) {
            const double *xv = src.f64 + v * DIM;
#pragma unroll
            for (int e = 0; e < DIM; e++) x[h][e] = (float)xv[e];
          } else {
            gather_lattice<DIM>(src, v, x[h]);
          }
#pragma unroll
          for (int e = 0; e < DIM; e++) xn[q] = fmaf(x[h][e], x[h][e], xn[q]);
        } else {
#pragma unroll
          for (int e = 0; e < DIM; e++) x[h][e] = 0.f;
        }
      }
#pragma unroll
      for (int e = 0; e < DIM; e++) xp[qp][e] = pack2(x[0][e], x[1][e]);
    }
    float best[Q], second[Q];
    int bpair[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
      best[q] = FLT_MAX;
      second[q] = FLT_MAX;
      bpair[q] = 0;
    }

    for (int chunk = 0; chunk < n_chunks; chunk++) {
      int kn;
      if (n_chunks == 1) {
        kn = K;
      } else {
        if (staged != chunk) {
          __syncthreads();  // everyone is done reading the previous chunk
          kn = stage_chunk(chunk);
          staged = chunk;
        } else {
          kn = min(k_chunk, K - chunk * k_chunk);
        }
      }
      const int k0 = chunk * k_chunk;
      const float4 *rows = reinterpret_cast<const float4 *>(s_cb);
#pragma unroll 1
      for (int k = 0; k < kn; k += 2) {
        float c0[ROW], c1[ROW];
#pragma unroll
        for (int r = 0; r < ROW / 4; r++) {
          const float4 t = rows[k * (ROW / 4) + r];
          const float4 u = rows[(k + 1) * (ROW / 4) + r];
          c0[4 * r + 0] = t.x; c0[4 * r + 1] = t.y; c0[4 * r + 2] = t.z; c0[4 * r + 3] = t.w;
          c1[4 * r + 0] = u.x; c1[4 * r + 1] = u.y; c1[4 * r + 2] = u.z; c1[4 * r + 3] = u.w;
        }
        const int kg = k0 + k;
#pragma unroll
        for (int qp = 0; qp < Q / 2; qp++) {
          unsigned long long a0 = pack2(c0[DIM], c0[DIM]);
          unsigned long long a1 = pack2(c1[DIM], c1[DIM]);
#pragma unroll
          for (int e = 0; e < DIM; e++) {
            a0 = ffma2(xp[qp][e], pack2(c0[e], c0[e]), a0);
            a1 = ffma2(xp[qp][e], pack2(c1[e], c1[e]), a1);
          }
          float s0[2], s1[2];
          unpack2(a0, s0[0], s0[1]);
          unpack2(a1, s1[0], s1[1]);
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int q = 2 * qp + h;
            const float lo = fminf(s0[h], s1[h]), hi = fmaxf(s0[h], s1[h]);
            second[q] = fmin3(second[q], hi, fmaxf(lo, best[q]));
            bpair[q] = lo < best[q] ? kg : bpair[q];
            best[q] = fminf(best[q], lo);
          }
        }
      }
    }

#pragma unroll
    for (int q = 0; q < Q; q++) {
      const unsigned long long v = v0 + (unsigned long long)q * THREADS + tid;
      // which member of the winning pair: the same FMA sequence on the same operands gives the
      // same bits as the loop did (rows from global memory / L1: measured faster than divergent shared-memory reads)
      const float *r0 = cb_rows + (size_t)bpair[q] * ROW;
      float t0 = __ldg(r0 + DIM), t1 = __ldg(r0 + ROW + DIM);
#pragma unroll
      for (int e = 0; e < DIM; e++) {
        float lo_, hi_;
        unpack2(xp[q / 2][e], lo_, hi_);
        const float xe = (q & 1) ? hi_ : lo_;
        t0 = fmaf(xe, __ldg(r0 + e), t0);
        t1 = fmaf(xe, __ldg(r0 + ROW + e), t1);
      }
      const int bidx = bpair[q] + (t1 < t0 ? 1 : 0);
      float r = sqrtf(xn[q]) + c_max_norm;
      const float margin = margin_coef * r * r;
      const bool flag = live[q] && !((second[q] - best[q]) > margin);
      if (live[q]) assign[v] = (uint32_t)bidx;
      // warp-aggregated append
      const unsigned int m = __ballot_sync(0xffffffffu, flag);
      if (m) {
        const int lane = tid & 31;
        const int leader = __ffs(m) - 1;
        unsigned int basepos = 0;
        if (lane == leader) basepos = atomicAdd(flag_count, (unsigned int)__popc(m));
        basepos = __shfl_sync(0xffffffffu, basepos, leader);
        if (flag) flag_list[basepos + __popc(m & ((1u << lane) - 1u))] = (uint32_t)v;
      }
      if (stats) {
        const int a = (live[q] && !flag) ? bidx : -1;
        int L[DIM], qs = 0;
#pragma unroll
        for (int e = 0; e < DIM; e++) {
          float lo_, hi_;
          unpack2(xp[q / 2][e], lo_, hi_);
          L[e] = a >= 0 ? (int)((q & 1) ? hi_ : lo_) : 0;
          qs += L[e] * L[e];
        }
        int *row = s_s + (a >= 0 ? a : 0) * DIM;
        if (k_real <= 8) {  // few cells: combine the lanes that hit the same cell first (see accumulate_match_kernel)
          const unsigned int group = __match_any_sync(0xffffffffu, a);
          const bool leader = a >= 0 && (tid & 31) == __ffs(group) - 1;
#pragma unroll
          for (int e = 0; e < DIM; e++) {
            const int sv = __reduce_add_sync(group, L[e]);
            if (leader && sv != 0) atomicAdd(row + e, sv);
          }
          const unsigned int qsum = __reduce_add_sync(group, (unsigned int)qs);
          if (leader) {
            atomicAdd(s_n + a, __popc(group));
            atomicAdd(s_q + a, (unsigned long long)qsum);
          }
        } else if (a >= 0) {
#pragma unroll
          for (int e = 0; e < DIM; e++)
            if (L[e] != 0) atomicAdd(row + e, L[e]);
          atomicAdd(s_n + a, 1);
          atomicAdd(s_q + a, (unsigned long long)qs);
        }
      }
    }
  }
  if (stats) {
    __syncthreads();
    for (int i = threadIdx.x; i < k_real; i += THREADS) {
      if (s_n[i] != 0) {
        unsigned long long *row = stats + (size_t)i * (DIM + 2);
        atomicAdd(row, (unsigned long long)s_n[i]);
        atomicAdd(row + DIM + 1, s_q[i]);
      }
    }
    for (int i = threadIdx.x; i < k_real * DIM; i += THREADS) {
      const int sv = s_s[i];
      if (sv != 0) {
        const int k = i / DIM, e = i - k * DIM;
        atomicAdd(stats + (size_t)k * (DIM + 2) + 1 + e, (unsigned long long)(long long)sv);
      }
    }
  }
}

// refilter_kernel: the FP32 MIDDLE TIER of the tensor-core levels.  The tensor-core filter's margin is three times a
// deliberately generous bound on the tensor core's internal accumulation error (55 units of 2^-23 at dim 12); the
// FP32 FMA score's bound is (dim + 3) * 2^-24, an order of magnitude tighter.  Queries the tensor-core filter could
// not decide are therefore re-ranked here - same inner loop as assign_kernel (FFMA2 pairs against the staged
// rows, exact top-2 over ALL codevectors) - and only what is still inside the FP32 margin goes on to the FP64
// resolver, whose cost per query is K distance evaluations at FP64 rate.  Input: `list_in` (count on the device),
// output: final indices for the decided ones, `list_out` (appended, warp-aggregated) for the rest; entries keep the
// "undecided" mark in bit 31 until the resolver rewrites them.
// Work split: the list is cut into one contiguous slice per CTA; thread t takes items t*Q .. t*Q+Q-1 of the
// slice's current tile, so a short list keeps whole warps idle instead of wasting lanes on dead queries.
template <int DIM>
__global__ void __launch_bounds__(AssignCfg<DIM>::THREADS, 1)
    refilter_kernel(const VecSource src, const float *__restrict__ cb_rows, const int K, const int k_chunk,
                    const float margin_coef, const float *__restrict__ c_max_ptr, uint32_t *__restrict__ assign,
                    const uint32_t *__restrict__ list_in, const unsigned int *__restrict__ count_in,
                    uint32_t *__restrict__ list_out, unsigned int *__restrict__ count_out) {
  using Cfg = AssignCfg<DIM>;
  constexpr int ROW = Cfg::ROW, Q = Cfg::Q, THREADS = Cfg::THREADS;
  const unsigned int total = *count_in;
  const unsigned int per_cta = (total + gridDim.x - 1) / gridDim.x;
  const unsigned int s_begin = min(total, blockIdx.x * per_cta), s_end = min(total, s_begin + per_cta);
  if (s_begin >= s_end) return;
  const float c_max_norm = *c_max_ptr;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *s_cb = reinterpret_cast<float *>(smem_raw);
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x;
  const int n_chunks = (K + k_chunk - 1) / k_chunk;
  uint32_t phase = 0;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto stage_chunk = [&](int chunk) {
    const int k0 = chunk * k_chunk;
    const int kn = min(k_chunk, K - k0);
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)kn * ROW * 4u;
      mbar_expect_tx(&s_bar, bytes);
      const char *g = reinterpret_cast<const char *>(cb_rows + (size_t)k0 * ROW);
      char *s = reinterpret_cast<char *>(s_cb);
      for (uint32_t off = 0; off < bytes; off += 32768u) tma_load_1d(s + off, g + off, min(32768u, bytes - off), &s_bar);
    }
    mbar_wait(&s_bar, phase);
    phase ^= 1;
    return kn;
  };
  int staged = -1;
  for (unsigned int t0 = s_begin; t0 < s_end; t0 += THREADS * Q) {
    unsigned long long xp[Q / 2][DIM];
    float xn[Q];
    bool live[Q];
    uint32_t vq[Q];
#pragma unroll
    for (int qp = 0; qp < Q / 2; qp++) {
      float x[2][DIM];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int q = 2 * qp + h;
        const unsigned int item = t0 + (unsigned int)tid * Q + q;
        live[q] = item < s_end;
        vq[q] = live[q] ? __ldg(list_in + item) : 0u;
        xn[q] = 0.f;
        if (live[q]) {
          gather_lattice<DIM>(src, vq[q], x[h]);
#pragma unroll
          for (int e = 0; e < DIM; e++) xn[q] = fmaf(x[h][e], x[h][e], xn[q]);
        } else {
#pragma unroll
          for (int e = 0; e < DIM; e++) x[h][e] = 0.f;
        }
      }
#pragma unroll
      for (int e = 0; e < DIM; e++) xp[qp][e] = pack2(x[0][e], x[1][e]);
    }
    const bool warp_live = __any_sync(0xffffffffu, live[0]);  // items are handed out thread by thread: q = 0 first
    float best[Q], second[Q];
    int bpair[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
      best[q] = FLT_MAX;
      second[q] = FLT_MAX;
      bpair[q] = 0;
    }
    for (int chunk = 0; chunk < n_chunks; chunk++) {
      int kn;
      if (staged != chunk) {
        __syncthreads();  // everyone is done reading the previous chunk
        kn = stage_chunk(chunk);
        staged = chunk;
      } else {
        kn = min(k_chunk, K - chunk * k_chunk);
      }
      if (!warp_live) continue;  // (the barrier above is taken by every warp)
      const int k0 = chunk * k_chunk;
      const float4 *rows = reinterpret_cast<const float4 *>(s_cb);
#pragma unroll 1
      for (int k = 0; k < kn; k += 2) {
        float c0[ROW], c1[ROW];
#pragma unroll
        for (int r = 0; r < ROW / 4; r++) {
          const float4 t = rows[k * (ROW / 4) + r];
          const float4 u = rows[(k + 1) * (ROW / 4) + r];
          c0[4 * r + 0] = t.x; c0[4 * r + 1] = t.y; c0[4 * r + 2] = t.z; c0[4 * r + 3] = t.w;
          c1[4 * r + 0] = u.x; c1[4 * r + 1] = u.y; c1[4 * r + 2] = u.z; c1[4 * r + 3] = u.w;
        }
        const int kg = k0 + k;
#pragma unroll
        for (int qp = 0; qp < Q / 2; qp++) {
          unsigned long long a0 = pack2(c0[DIM], c0[DIM]);
          unsigned long long a1 = pack2(c1[DIM], c1[DIM]);
#pragma unroll
          for (int e = 0; e < DIM; e++) {
            a0 = ffma2(xp[qp][e], pack2(c0[e], c0[e]), a0);
            a1 = ffma2(xp[qp][e], pack2(c1[e], c1[e]), a1);
          }
          float s0[2], s1[2];
          unpack2(a0, s0[0], s0[1]);
          unpack2(a1, s1[0], s1[1]);
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int q = 2 * qp + h;
            const float lo = fminf(s0[h], s1[h]), hi = fmaxf(s0[h], s1[h]);
            second[q] = fmin3(second[q], hi, fmaxf(lo, best[q]));
            bpair[q] = lo < best[q] ? kg : bpair[q];
            best[q] = fminf(best[q], lo);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < Q; q++) {
      // which member of the winning pair: the same FMA sequence on the same operands gives the same bits as the loop
      const float *r0 = cb_rows + (size_t)bpair[q] * ROW;
      float t0s = __ldg(r0 + DIM), t1s = __ldg(r0 + ROW + DIM);
#pragma unroll
      for (int e = 0; e < DIM; e++) {
        float lo_, hi_;
        unpack2(xp[q / 2][e], lo_, hi_);
        const float xe = (q & 1) ? hi_ : lo_;
        t0s = fmaf(xe, __ldg(r0 + e), t0s);
        t1s = fmaf(xe, __ldg(r0 + ROW + e), t1s);
      }
      const int bidx = bpair[q] + (t1s < t0s ? 1 : 0);
      const float r = sqrtf(xn[q]) + c_max_norm;
      const bool flag = live[q] && !((second[q] - best[q]) > margin_coef * r * r);
      if (live[q]) assign[vq[q]] = (uint32_t)bidx | (flag ? 0x80000000u : 0u);
      const unsigned int m = __ballot_sync(0xffffffffu, flag);
      if (m) {
        const int lane = tid & 31, leader = __ffs(m) - 1;
        unsigned int basepos = 0;
        if (lane == leader) basepos = atomicAdd(count_out, (unsigned int)__popc(m));
        basepos = __shfl_sync(0xffffffffu, basepos, leader);
        if (flag) list_out[basepos + __popc(m & ((1u << lane) - 1u))] = vq[q];
      }
    }
  }
}

// Early split levels (K <= 16): filter + per-cell statistics in ONE high-occupancy pass.  With a handful of
// codevectors the work per vector is tiny, so the persistent one-CTA-per-SM kernel above and a separate
// accumulate pass are both dominated by their fixed costs; here every thread takes one vector per iteration,
// scores it against the K rows (shared memory, broadcast reads), applies the same margin rule, and the
// vectors it DECIDES are accumulated at once: lanes that chose the same cell are combined with
// __match_any_sync + __reduce_add_sync and only group leaders touch the per-CTA table (flagged vectors are
// added by the resolver).  stats may be null (assignment only).
//
// Statistics, KCAP > 0 (K <= KCAP): every thread keeps its OWN partial sums for all KCAP cells in registers - per cell
// (DIM+1)/2 words of two 16-bit lattice sums (t = L + 128 <= 255, so 256 vectors fit), a count and a sum of squares -
// and adds a vector to its cell with predicated integer adds: no shuffles, no atomics in the loop.  Every 256
// vectors (and at the end) the registers are summed across the warp (REDUX) into the per-CTA table.
// KCAP == 0: the register file cannot hold KCAP cells of this dimension; lanes that chose the same cell are combined
// per vector with __match_any_sync + __reduce_add_sync instead.
template <int DIM, int KCAP>
__global__ void __launch_bounds__(256)
    small_k_fused_kernel(const VecSource src, const float *__restrict__ cb_rows, const int K, const float margin_coef,
                         const float *__restrict__ c_max_ptr, uint32_t *__restrict__ assign,
                         uint32_t *__restrict__ flag_list, unsigned int *__restrict__ flag_count,
                         unsigned long long *__restrict__ stats) {
  constexpr int ROW = AssignCfg<DIM>::ROW, KMAX = 16;
  __shared__ __align__(16) float s_rows[KMAX * ROW];
  __shared__ unsigned long long s_q[KMAX];
  __shared__ int s_n[KMAX];
  __shared__ int s_s[KMAX * DIM];
  for (int i = threadIdx.x; i < K * ROW; i += blockDim.x) s_rows[i] = cb_rows[i];
  for (int i = threadIdx.x; i < KMAX; i += blockDim.x) {
    s_q[i] = 0;
    s_n[i] = 0;
  }
  for (int i = threadIdx.x; i < KMAX * DIM; i += blockDim.x) s_s[i] = 0;
  __syncthreads();
  const float c_max_norm = *c_max_ptr;
  const int lane = threadIdx.x & 31;
  constexpr int P = (DIM + 1) / 2, KC = KCAP > 0 ? KCAP : 1;
  unsigned int acc_s[KC][P], acc_n[KC], acc_q[KC];
  int pending = 0;  // vectors in the register accumulators since the last flush
#pragma unroll
  for (int k = 0; k < KC; k++) {
    acc_n[k] = 0;
    acc_q[k] = 0;
#pragma unroll
    for (int j = 0; j < P; j++) acc_s[k][j] = 0;
  }
  auto flush_registers = [&]() {  // warp-wide sums of the register accumulators -> per-CTA table, lane 0 adds
#pragma unroll
    for (int k = 0; k < KC; k++) {
      if (k < K) {
        const unsigned int n = __reduce_add_sync(0xffffffffu, acc_n[k]);
        if (n) {  // warp-uniform
          const unsigned int q = __reduce_add_sync(0xffffffffu, acc_q[k]);
          if (lane == 0) {
            atomicAdd(s_n + k, (int)n);
            atomicAdd(s_q + k, (unsigned long long)q);
          }
#pragma unroll
          for (int j = 0; j < P; j++) {
            const unsigned int lo = __reduce_add_sync(0xffffffffu, acc_s[k][j] & 0xffffu);
            const unsigned int hi = __reduce_add_sync(0xffffffffu, acc_s[k][j] >> 16);
            if (lane == 0) {
              if (lo) atomicAdd(s_s + k * DIM + 2 * j, (int)lo);
              if (2 * j + 1 < DIM && hi) atomicAdd(s_s + k * DIM + 2 * j + 1, (int)hi);
            }
          }
        }
      }
      acc_n[k] = 0;
      acc_q[k] = 0;
#pragma unroll
      for (int j = 0; j < P; j++) acc_s[k][j] = 0;
    }
    pending = 0;
  };
  const unsigned long long n_round = (src.n_local + 31ull) & ~31ull;  // whole warps stay in the loop
  for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_round;
       v += (unsigned long long)gridDim.x * blockDim.x) {
    const bool live = v < src.n_local;
    float x[DIM], xn = 0.f;
    if (live) {
      gather_lattice<DIM>(src, v, x);
    } else {
#pragma unroll
      for (int e = 0; e < DIM; e++) x[e] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < DIM; e++) xn = fmaf(x[e], x[e], xn);
    float best = FLT_MAX, second = FLT_MAX;
    int bidx = 0;
    for (int k = 0; k < K; k++) {
      const float4 *r4 = reinterpret_cast<const float4 *>(s_rows + k * ROW);
      float c[ROW];
#pragma unroll
      for (int q = 0; q < ROW / 4; q++) {
        const float4 t = r4[q];
        c[4 * q] = t.x; c[4 * q + 1] = t.y; c[4 * q + 2] = t.z; c[4 * q + 3] = t.w;
      }
      float sc = c[DIM];
#pragma unroll
      for (int e = 0; e < DIM; e++) sc = fmaf(x[e], c[e], sc);
      second = fminf(second, fmaxf(sc, best));
      bidx = sc < best ? k : bidx;
      best = fminf(best, sc);
    }
    const float rr = sqrtf(xn) + c_max_norm;
    const bool flag = live && !((second - best) > margin_coef * rr * rr);
    if (live) assign[v] = (uint32_t)bidx;
    const unsigned int m = __ballot_sync(0xffffffffu, flag);
    if (m) {
      const int leader = __ffs(m) - 1;
      unsigned int basepos = 0;
      if (lane == leader) basepos = atomicAdd(flag_count, (unsigned int)__popc(m));
      basepos = __shfl_sync(0xffffffffu, basepos, leader);
      if (flag) flag_list[basepos + __popc(m & ((1u << lane) - 1u))] = (uint32_t)v;
    }
    if (stats && KCAP > 0) {
      const int a = (live && !flag) ? bidx : -1;
      unsigned int px[P], qs = 0;
#pragma unroll
      for (int j = 0; j < P; j++) {
        const int L0 = (int)x[2 * j], L1 = 2 * j + 1 < DIM ? (int)x[2 * j + 1 < DIM ? 2 * j + 1 : 2 * j] : -128;
        qs += (unsigned int)(L0 * L0) + (2 * j + 1 < DIM ? (unsigned int)(L1 * L1) : 0u);
        px[j] = (unsigned int)(L0 + 128) | ((unsigned int)(L1 + 128) << 16);
      }
#pragma unroll
      for (int k = 0; k < KC; k++) {
        const bool hit = a == k;
        acc_n[k] += hit ? 1u : 0u;
        acc_q[k] += hit ? qs : 0u;
#pragma unroll
        for (int j = 0; j < P; j++) acc_s[k][j] += hit ? px[j] : 0u;
      }
      if (++pending == 256) flush_registers();  // 256 * 255 < 2^16: the packed 16-bit sums cannot overflow
    } else if (stats) {
      const int a = (live && !flag) ? bidx : -1;
      const unsigned int group = __match_any_sync(0xffffffffu, a);
      const bool lead = a >= 0 && lane == __ffs(group) - 1;
      // t = L + 128 in [0, 255]: two coordinates per register, 16 bits each (32 lanes x 255 < 2^16), so one REDUX
      // sums two coordinates; the table holds sums of t and is turned back into sums of L when it is flushed
      int *row = s_s + (a >= 0 ? a : 0) * DIM;
      int qs = 0;
#pragma unroll
      for (int e = 0; e < DIM; e += 2) {
        const int L0 = a >= 0 ? (int)x[e] : -128;
        const int L1 = (e + 1 < DIM && a >= 0) ? (int)x[e + 1 < DIM ? e + 1 : e] : -128;
        qs += (a >= 0 ? L0 * L0 : 0) + ((e + 1 < DIM && a >= 0) ? L1 * L1 : 0);
        const unsigned int sv = __reduce_add_sync(group, (unsigned int)(L0 + 128) | ((unsigned int)(L1 + 128) << 16));
        if (lead) {
          if (sv & 0xffffu) atomicAdd(row + e, (int)(sv & 0xffffu));
          if (e + 1 < DIM && (sv >> 16)) atomicAdd(row + e + 1, (int)(sv >> 16));
        }
      }
      const unsigned int qsum = __reduce_add_sync(group, (unsigned int)qs);
      if (lead) {
        atomicAdd(s_n + a, __popc(group));
        atomicAdd(s_q + a, (unsigned long long)qsum);
      }
    }
  }
  if (stats) {
    if (KCAP > 0) flush_registers();
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
      if (s_n[i] != 0) {
        unsigned long long *row = stats + (size_t)i * (DIM + 2);
        atomicAdd(row, (unsigned long long)s_n[i]);
        atomicAdd(row + DIM + 1, s_q[i]);
      }
    }
    for (int i = threadIdx.x; i < K * DIM; i += blockDim.x) {
      const int k = i / DIM, e = i - k * DIM;
      const int sv = s_s[i] - 128 * s_n[k];  // sum of t -> sum of L
      if (sv != 0) atomicAdd(stats + (size_t)k * (DIM + 2) + 1 + e, (unsigned long long)(long long)sv);
    }
  }
}

// Generic-dimension filter (any dim <= kMaxDim that has no template instance): one query per
// thread, the query lives in shared memory (column per thread), codebook rows streamed from L2.
// Same score, same margin rule; slower, only there so that every block shape works.
__global__ void __launch_bounds__(128, 1)
    assign_generic_kernel(const VecSource src, const float *__restrict__ cb_rows, const int K, const int row,
                          const float margin_coef, const float *__restrict__ c_max_ptr, uint32_t *__restrict__ assign,
                          uint32_t *__restrict__ flag_list, unsigned int *__restrict__ flag_count) {
  const float c_max_norm = *c_max_ptr;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *s_x = reinterpret_cast<float *>(smem_raw);  // [dim][128]
  const int dim = src.dim, tid = threadIdx.x;
  for (unsigned long long v0 = (unsigned long long)blockIdx.x * 128; v0 < src.n_local;
       v0 += (unsigned long long)gridDim.x * 128) {
    const unsigned long long v = v0 + tid;
    const bool live = v < src.n_local;
    float xn = 0.f;
    if (live && src.f64) {  // general FP64 vectors: the filter sees them rounded to FP32 (the margin covers it)
      const double *xv = src.f64 + v * (unsigned long long)dim;
      for (int e = 0; e < dim; e++) {
        const float f = (float)xv[e];
        s_x[e * 128 + tid] = f;
        xn = fmaf(f, f, xn);
      }
    } else if (live) {
      unsigned long long base, img;
      vec_base(src, v, base, img);
      for (int e = 0; e < dim; e++) {
        float f = (float)load_lattice(src, img, base, e);
        s_x[e * 128 + tid] = f;
        xn = fmaf(f, f, xn);
      }
    } else {
      for (int e = 0; e < dim; e++) s_x[e * 128 + tid] = 0.f;
    }
    float best = FLT_MAX, second = FLT_MAX;
    int bidx = 0;
    for (int k = 0; k < K; k++) {
      const float *c = cb_rows + (size_t)k * row;
      float s = __ldg(c + dim);
      for (int e = 0; e < dim; e++) s = fmaf(s_x[e * 128 + tid], __ldg(c + e), s);
      second = fminf(second, fmaxf(s, best));
      const bool better = s < best;
      best = fminf(best, s);
      bidx = better ? k : bidx;
    }
    float r = sqrtf(xn) + c_max_norm;
    const bool flag = live && !((second - best) > margin_coef * r * r);
    if (live) assign[v] = (uint32_t)bidx;
    const unsigned int m = __ballot_sync(0xffffffffu, flag);
    if (m) {
      const int lane = tid & 31, leader = __ffs(m) - 1;
      unsigned int basepos = 0;
      if (lane == leader) basepos = atomicAdd(flag_count, (unsigned int)__popc(m));
      basepos = __shfl_sync(0xffffffffu, basepos, leader);
      if (flag) flag_list[basepos + __popc(m & ((1u << lane) - 1u))] = (uint32_t)v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// resolve_kernel: exact FP64 nearest neighbour in nanoflann's traversal order
// ------------------------------------------------------------------------------------------------
// Phase B: one warp per query phase A (below) could not decide.  Arithmetic is spelled with the round-to-nearest intrinsics so
// that nvcc cannot contract a*b+c into an FMA: the reference's x86-64 build has none.
// Adds one query to its cell's statistics row {n, S[dim], Q} (used by the resolver when the filter kernel
// accumulates the queries it decided itself).
// Called by ALL lanes of the warp that resolved the query: lane l adds dimensions l, l + 32, ... (one atomic each, side by
// side), lane 0 the count and the sum of squares.  (One lane doing all dim + 2 atomics in turn made the resolver 2.6x
// slower on 48-dimensional vectors, where ~2 % of the queries are flagged.)
__device__ __forceinline__ void add_query_stats(const VecSource &src, unsigned long long img, unsigned long long base,
                                                unsigned long long *row, const int lane) {
  const int dim = src.dim;
  unsigned int q = 0;
  for (int e = lane; e < dim; e += 32) {
    const int L = load_lattice(src, img, base, e);
    q += (unsigned int)(L * L);
    if (L != 0) atomicAdd(row + 1 + e, (unsigned long long)(long long)L);
  }
  q = __reduce_add_sync(0xffffffffu, q);
  if (lane == 0) {
    atomicAdd(row, 1ull);
    atomicAdd(row + dim + 1, (unsigned long long)q);
  }
}

__device__ __forceinline__ double sq_diff(double a, double b) {
  const double d = __dsub_rn(a, b);
  return __dmul_rn(d, d);
}

// L2_Adaptor::operator() (nanoflann.hpp:320-345) with worst_dist = -1: groups of four summed as
// ((d0^2 + d1^2) + d2^2) + d3^2 and added to the running result, then a scalar tail.
__device__ __forceinline__ double nanoflann_l2(const double *a, const double *__restrict__ b, int dim) {
  double result = 0.0;
  int d = 0;
  for (; d + 3 < dim; d += 4) {
    const double g = __dadd_rn(__dadd_rn(__dadd_rn(sq_diff(a[d], b[d]), sq_diff(a[d + 1], b[d + 1])),
                                         sq_diff(a[d + 2], b[d + 2])),
                               sq_diff(a[d + 3], b[d + 3]));
    result = __dadd_rn(result, g);
  }
  for (; d < dim; d++) result = __dadd_rn(result, sq_diff(a[d], b[d]));
  return result;
}

// Compile-time dimension, row-major codevector: fully unrolled.
template <int DIM>
__device__ __forceinline__ double nanoflann_l2_fixed(const double *a, const double *b) {
  double result = 0.0;
#pragma unroll
  for (int d = 0; d + 3 < DIM; d += 4) {
    const double g = __dadd_rn(__dadd_rn(__dadd_rn(sq_diff(a[d], b[d]), sq_diff(a[d + 1], b[d + 1])), sq_diff(a[d + 2], b[d + 2])),
                               sq_diff(a[d + 3], b[d + 3]));
    result = __dadd_rn(result, g);
  }
#pragma unroll
  for (int d = DIM & ~3; d < DIM; d++) result = __dadd_rn(result, sq_diff(a[d], b[d]));
  return result;
}

// Compile-time dimension: fully unrolled (loads issued back to back, the query stays in registers).
template <int DIM>
__device__ __forceinline__ double nanoflann_l2_t_fixed(const double *a, const double *__restrict__ bt, size_t stride) {
  double result = 0.0;
#pragma unroll
  for (int d = 0; d + 3 < DIM; d += 4) {
    const double g = __dadd_rn(__dadd_rn(__dadd_rn(sq_diff(a[d], bt[d * stride]), sq_diff(a[d + 1], bt[(d + 1) * stride])),
                                         sq_diff(a[d + 2], bt[(d + 2) * stride])),
                               sq_diff(a[d + 3], bt[(d + 3) * stride]));
    result = __dadd_rn(result, g);
  }
#pragma unroll
  for (int d = DIM & ~3; d < DIM; d++) result = __dadd_rn(result, sq_diff(a[d], bt[d * stride]));
  return result;
}

// Same arithmetic on the TRANSPOSED codebook (element e of codevector k at bt[e * stride + k]): lanes that
// hold consecutive k read consecutive doubles, so a warp-wide load touches 2 cache lines instead of 32.
__device__ __forceinline__ double nanoflann_l2_t(const double *a, const double *__restrict__ bt, size_t stride, int dim) {
  double result = 0.0;
  int d = 0;
  for (; d + 3 < dim; d += 4) {
    const double g = __dadd_rn(__dadd_rn(__dadd_rn(sq_diff(a[d], bt[d * stride]), sq_diff(a[d + 1], bt[(d + 1) * stride])),
                                         sq_diff(a[d + 2], bt[(d + 2) * stride])),
                               sq_diff(a[d + 3], bt[(d + 3) * stride]));
    result = __dadd_rn(result, g);
  }
  for (; d < dim; d++) result = __dadd_rn(result, sq_diff(a[d], bt[d * stride]));
  return result;
}


// Phase A of the resolver: one WARP per flagged query, exact FP64 distances (nanoflann's arithmetic,
// so the values are the ones the reference's leaf loop computes) to ALL K codevectors, lanes strided
// over k.  nanoflann's search is exact, so whenever the smallest distance is separated from every
// other one the tree walk must return that codevector and no walk is needed: the only FP64 roundings
// that differ between the walk and this loop are in the walk's pruning bound (mindistsq + cut_dist -
// dists[feat], nanoflann.hpp:1254-1262), whose intermediates are distances from the query to actual
// codevector coordinates, i.e. bounded by dmax = max_k dist_k: an add and a subtract per tree level on values of at
// most 2 * dmax, each rounding <= 2^-52 * dmax (depth <= kResolveDepthCap = 512: 2^-42 * dmax in all, 2^-46 at the
// depths that occur), plus the (dim + 1) * 2^-53 relative error of a distance (<= 2^-45.4 at dim 192) - so a subtree
// holding a codevector more than ~2^-41.5 * dmax closer than the running best is never pruned.  A query is decided
// here when it has exactly one candidate within band = 2^-36 * dmax of the minimum (45x above that worst case, many
// orders below the FP32 filter's margin; a wider band only sends more queries to the walk).  EXACT ties (all candidates bitwise equal) are decided by the tree's
// visiting order without a walk (see below); what remains - distinct distances closer than the band, or
// more than 32 candidates - goes to the tie list and phase B walks the tree.
template <int DIMCAP, int DIMT>  // DIMT != 0: dimension known at compile time (== DIMCAP)
__global__ void __launch_bounds__(128)
    resolve_bruteforce_kernel(const VecSource src, const int scaled, const double *__restrict__ cbt, const int K,
                              const KdDevice tree, const uint32_t *__restrict__ flag_list,
                              const unsigned int *__restrict__ flag_count, uint32_t *__restrict__ assign,
                              uint32_t *__restrict__ tie_list, unsigned int *__restrict__ tie_count,
                              unsigned int *__restrict__ changed, unsigned long long *__restrict__ stats,
                              uint32_t *__restrict__ result, unsigned int *__restrict__ sensitive,
                              const unsigned char *__restrict__ cv_exact, const int tree_robust) {
  // The filter's guess for a flagged query may carry the "undecided" mark in bit 31 (tensor-core finalise kernels: the
  // statistics pass that runs next to this kernel skips marked entries).  Every flagged query gets its exact index
  // written here or in phase B - to `result` when given (committed to `assign` once that pass is done), else in place.
  const int dim = DIMT ? DIMT : src.dim;
  const unsigned int total = *flag_count;
  const int lane = threadIdx.x & 31;
  const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned int n_warps = (gridDim.x * blockDim.x) >> 5;
  double x[DIMCAP];
  for (unsigned int f = warp; f < total; f += n_warps) {
    const unsigned long long v = flag_list[f];
    unsigned long long base = 0, img = 0;
    const double *xv = src.f64 ? src.f64 + v * (unsigned long long)dim : nullptr;  // general FP64 vectors: as they are
    if (!xv) vec_base(src, v, base, img);
#pragma unroll
    for (int e = 0; e < (DIMT ? DIMT : dim); e++) {
      if (xv) {
        x[e] = xv[e];
      } else {
        const double L = (double)load_lattice(src, img, base, e);
        x[e] = scaled ? __ddiv_rn(__dadd_rn(L, 128.0), 255.0) : L;
      }
    }
    double d1 = DBL_MAX, d2 = DBL_MAX, dmax = 0.0;  // this lane's smallest, second smallest, largest
    int k1 = 0;
    for (int k = lane; k < K; k += 32) {
      const double d = DIMT ? nanoflann_l2_t_fixed<DIMT ? DIMT : 1>(x, cbt + k, (size_t)K) : nanoflann_l2_t(x, cbt + k, (size_t)K, dim);
      dmax = fmax(dmax, d);
      if (d < d1) {
        d2 = d1;
        d1 = d;
        k1 = k;
      } else if (d < d2) {
        d2 = d;
      }
    }
    double wmin = d1, wmax = dmax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wmin = fmin(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
      wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    }
    const double lim = wmin + wmax * 1.4551915228366852e-11;  // 2^-36
    const int mine = (d1 <= lim ? 1 : 0) + (d2 <= lim ? 1 : 0);
    const unsigned int holders = __ballot_sync(0xffffffffu, mine > 0);
    const unsigned int multi = __ballot_sync(0xffffffffu, mine > 1);
    // Decisions the LAST BITS of the codebook could change (counted for the auto centroid mode): a second codevector
    // within lim2 = 2^-44 dmax of the minimum.  (Centroids from the integer sums are within 4e-16 = 2^-51 relative of the
    // reference's compensated sums; that moves a distance by less than 2^-48 dmax.)  Refined below: when every
    // codevector inside that band is one the integer path reproduces BIT FOR BIT (children of a cell whose members
    // are all one vector, dead cells: cv_exact) and the minimum is attained once, the exact search returns that
    // minimum whatever the other codevectors' last bits are - not sensitive.
    const double lim2 = wmin + wmax * 5.6843418860808015e-14;  // 2^-44
    int win = -1;
    if (__popc(holders) == 1 && multi == 0) {
      win = __shfl_sync(0xffffffffu, k1, __ffs(holders) - 1);
    } else {
      // EXACT ties only (every candidate's distance is bitwise the minimum: duplicated codevectors - dead
      // cells - or children 1.2c / 0.8c of a single-member cell): the walk keeps the FIRST candidate it
      // visits (leaf test `dist < worst` and KNNResultSet::addPoint are strict, nanoflann.hpp:1219-1224,
      // :121) and it cannot prune that candidate's subtree, whose bound is below the running worst by the
      // band.  Visiting order = near child first at every inner node (nanoflann.hpp:1242-1251), leaf points
      // in vind order - so descend from the root through the subtrees that still contain candidates.
      unsigned int my_pos = 0xffffffffu;  // up to 32 candidates, one per lane
      int n_cand = 0;
      bool exact = true;
      int near2 = 0, at_min = 0, inexact2 = 0;  // census of the 2^-44 band (this lane's share)
      for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        bool cand = false;
        if (k < K) {
          const double d = DIMT ? nanoflann_l2_t_fixed<DIMT ? DIMT : 1>(x, cbt + k, (size_t)K) : nanoflann_l2_t(x, cbt + k, (size_t)K, dim);
          cand = d <= lim;
          exact = exact && (!cand || d == wmin);
          if (d <= lim2) {
            near2++;
            at_min += d == wmin;
            inexact2 += !(cv_exact && cv_exact[k]);
          }
        }
        const unsigned int m = __ballot_sync(0xffffffffu, cand);
        const int slot = n_cand + __popc(m & ((1u << lane) - 1u));
        const unsigned int pos = cand ? tree.inv[k] : 0xffffffffu;
        // hand candidate `pos` to lane `slot` (slots >= 32 are dropped and force the slow walk)
#pragma unroll 1
        for (unsigned int mm = m; mm; mm &= mm - 1) {
          const int srcl = __ffs(mm) - 1;
          const int s_slot = __shfl_sync(0xffffffffu, slot, srcl);
          const unsigned int s_pos = __shfl_sync(0xffffffffu, pos, srcl);
          if (lane == s_slot) my_pos = s_pos;
        }
        n_cand += __popc(m);
      }
      if (sensitive) {
        near2 = __reduce_add_sync(0xffffffffu, near2);
        at_min = __reduce_add_sync(0xffffffffu, at_min);
        inexact2 = __reduce_add_sync(0xffffffffu, inexact2);
      }
      bool fragile = false;  // a step of the descent below within rounding noise of going the other way
      bool path_fragile = false;  // an inner node on the way to the first-visited candidate whose split the host's census
                                  // calls fragile (kKdNodeFragile): only what lies UNDER such a node can be arranged
                                  // differently in the reference's tree
      const bool all_exact = __all_sync(0xffffffffu, exact);
      if (all_exact && n_cand <= 32) {
        int node = 0, lo = 0, hi = K;
        for (int guard = 0; guard < 4096; guard++) {
          const KdNode nd = tree.nodes[node];
          if (nd.child1 < 0 && nd.child2 < 0) break;
          path_fragile = path_fragile || (nd.a & kKdNodeFragile) != 0;
          const int mid = nd.b;
          const bool in1 = my_pos != 0xffffffffu && (int)my_pos >= lo && (int)my_pos < mid;
          const bool in2 = my_pos != 0xffffffffu && (int)my_pos >= mid && (int)my_pos < hi;
          const bool any1 = __any_sync(0xffffffffu, in1), any2 = __any_sync(0xffffffffu, in2);
          bool go1;
          if (any1 && any2) {
            double val = 0.0;  // x[feat] by selection: a dynamic index would push x[] out of registers
            const int feat = nd.a & kKdFeatMask;
#pragma unroll
            for (int e = 0; e < (DIMT ? DIMT : DIMCAP); e++) val = (e == feat) ? x[e] : val;
            const double side = __dadd_rn(__dsub_rn(val, nd.divlow), __dsub_rn(val, nd.divhigh));
            go1 = side < 0;  // nearer child first
            // Robust against last-bit changes of the codebook?  Yes when the margin is comfortable - or when both
            // plane coordinates ARE coordinates of tied candidates (the query midway between the children 1.2c / 0.8c
            // of its own single-member cell makes `side` a pure rounding residue, but of bit-reproducible operands)
            // AND no other codevector's coordinate is within rounding noise of them (an inexact point a last bit above
            // 0.8c in the reference's codebook would be the plane there: the host's census bits of this node).
            if (sensitive && fabs(side) <= 1e-9 * (fabs(val) + fabs(nd.divlow) + fabs(nd.divhigh))) {
              double coord = 0.0;
              if (in1 || in2) coord = cbt[(size_t)feat * K + tree.vind[my_pos]];
              const bool lo_ok = (nd.a & kKdDivLowExact) && __any_sync(0xffffffffu, in1 && coord == nd.divlow);
              const bool hi_ok = (nd.a & kKdDivHighExact) && __any_sync(0xffffffffu, in2 && coord == nd.divhigh);
              fragile = fragile || !(lo_ok && hi_ok);
            }
          } else {
            go1 = any1;
          }
          if (go1) {
            node = nd.child1;
            hi = mid;
            if (!in1) my_pos = 0xffffffffu;
          } else {
            node = nd.child2;
            lo = mid;
            if (!in2) my_pos = 0xffffffffu;
          }
        }
        unsigned int first = my_pos;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        win = (int)tree.vind[first];
      }
      // Sensitive to the last bits of the codebook?  Yes when a codevector in the 2^-44 band is not reproduced bit for
      // bit by the integer path.  When they all are and the minimum is attained once, no (the exact search returns
      // it).  When it is attained several times the visiting order decides: that order is the same for any codebook
      // within a few ulps when every step of the descent is robust and the tree's shape is - as a whole
      // (KdHostTree::min_margin) or at least at every node from the root to the winner's leaf: a fragile split only
      // rearranges its own subtree, which the walk enters after the winner's unless it lies on that path.
      if (sensitive && lane == 0 && near2 > 1) {
        const bool order_safe = (tree_robust || !path_fragile) && all_exact && n_cand <= 32 && !fragile;
        // (Bit-reproducible candidates at DISTINCT distances inside the band are taken as decided by the exact search:
        //  it returns the smaller one unless the walk's pruning bound for that candidate's subtree is within ~2^-46 dmax
        //  of its full distance - a box corner reached along every dimension - AND the bound's plane coordinates differ
        //  between the codebooks.  Counting them as sensitive was tried: the 1.2c / 0.8c pair of a two-member cell is
        //  such a pair, an ulp apart, and every large noise train would take the exact repeat for it.)
        if (inexact2 > 0 || (at_min > 1 && !order_safe)) atomicAdd(sensitive, 1u);
        // diagnostics: sensitive[1] counts the first kind, sensitive[2] collects why a visiting order was not safe
        if (inexact2 > 0) atomicAdd(sensitive + 1, 1u);
        if (inexact2 == 0 && at_min > 1 && !order_safe)
          atomicOr(sensitive + 2, (tree_robust ? 0u : 1u) | (all_exact ? 0u : 2u) | (n_cand <= 32 ? 0u : 4u) | (fragile ? 8u : 0u));
      }
    }
    if (win >= 0) {
      if (lane == 0) {
        const uint32_t old = assign[v] & 0x7fffffffu;
        (result ? result : assign)[v] = (uint32_t)win;
        if (old != (uint32_t)win) atomicAdd(changed, 1u);
      }
      if (stats) add_query_stats(src, img, base, stats + (size_t)win * (dim + 2), lane);
    } else if (lane == 0) {
      tie_list[atomicAdd(tie_count, 1u)] = (uint32_t)v;
    }
  }
}

struct Frame {
  int node, other, feat, phase;
  double mind, cut, saved;
};

template <int DIMCAP, int DEPTHCAP, int DIMT>  // DIMT != 0: dimension known at compile time (== DIMCAP)
__global__ void __launch_bounds__(128)
    resolve_kernel(const VecSource src, const int scaled, const double *cb, const int K, const KdDevice tree,
                   const uint32_t *__restrict__ flag_list, const unsigned int *__restrict__ flag_count,
                   uint32_t *__restrict__ assign, unsigned int *__restrict__ changed,
                   unsigned long long *__restrict__ stats, const unsigned int stage_bytes, uint32_t *__restrict__ result) {
  // One WARP per query: every lane runs the same (sequential) tree walk; at a leaf the lanes compute the
  // distances of its <= 10 points in parallel.  nanoflann's leaf loop (read worstDist once, add points in
  // order, strict comparisons) keeps the first point that attains the leaf minimum, and only if that
  // minimum is strictly below the best so far - which is what the lane-ordered reduction below returns.
  // The walk is a chain of dependent loads (node -> vind -> codevector), so a block that has work first
  // copies tree and codebook into shared memory when they fit (stage_bytes != 0).
  const int dim = DIMT ? DIMT : src.dim;
  const unsigned int total = *flag_count;
  const int lane = threadIdx.x & 31;
  const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned int n_warps = (gridDim.x * blockDim.x) >> 5;
  if (((blockIdx.x * blockDim.x) >> 5) >= total) return;  // no query for any warp of this block
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const KdNode *nodes = tree.nodes;
  const unsigned int *vind = tree.vind;
  if (stage_bytes) {
    const size_t nb_nodes = (size_t)tree.n_nodes * sizeof(KdNode), nb_cb = (size_t)K * dim * 8, nb_vind = (size_t)K * 4;
    // bulk copies by the TMA unit (one thread issues, everyone waits on the mbarrier): ~130 KB at K = 1024 arrive in
    // a few microseconds, where a copy loop of this 128-thread block took ~100 us - longer than the walks themselves
    __shared__ __align__(8) uint64_t s_bar;
    if (threadIdx.x == 0) {
      mbar_init(&s_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t b0 = (uint32_t)nb_cb, b1 = (uint32_t)nb_nodes, b2 = (uint32_t)((nb_vind + 15) & ~(size_t)15);
      mbar_expect_tx(&s_bar, b0 + b1 + b2);  // cb and nodes are multiples of 16 bytes; vind is followed by its inverse
      auto bulk = [&](unsigned char *d, const void *g, uint32_t bytes) {
        for (uint32_t off = 0; off < bytes; off += 32768u)
          tma_load_1d(d + off, reinterpret_cast<const char *>(g) + off, min(32768u, bytes - off), &s_bar);
      };
      bulk(smem_raw, cb, b0);
      bulk(smem_raw + b0, tree.nodes, b1);
      bulk(smem_raw + b0 + b1, tree.vind, b2);
    }
    mbar_wait_bounded(&s_bar, 0, 9);
    cb = reinterpret_cast<const double *>(smem_raw);
    nodes = reinterpret_cast<const KdNode *>(smem_raw + nb_cb);
    vind = reinterpret_cast<const unsigned int *>(smem_raw + nb_cb + nb_nodes);
  }
  double x[DIMCAP], dists[DIMCAP];
  Frame stack[DEPTHCAP];
  for (unsigned int f = warp; f < total; f += n_warps) {
    const unsigned long long v = flag_list[f];
    unsigned long long base = 0, img = 0;
    const double *xv = src.f64 ? src.f64 + v * (unsigned long long)dim : nullptr;
    if (!xv) vec_base(src, v, base, img);
#pragma unroll
    for (int e = 0; e < (DIMT ? DIMT : dim); e++) {
      if (xv) {
        x[e] = xv[e];
      } else {
        const double L = (double)load_lattice(src, img, base, e);
        // ScaledColor::RGBtoColorSpace: ((double)c + 128.0) / 255 (src/ColorSpace.cpp:16-21);
        // a past-the-end element is the literal 0.0 and (-128 + 128)/255 == 0.0 as well.
        x[e] = scaled ? __ddiv_rn(__dadd_rn(L, 128.0), 255.0) : L;
      }
    }
    // computeInitialDistances (nanoflann.hpp:1188-1205)
    double distsq = 0.0;
#pragma unroll
    for (int e = 0; e < (DIMT ? DIMT : dim); e++) {
      dists[e] = 0.0;
      if (x[e] < tree.bbox_low[e]) {
        dists[e] = sq_diff(x[e], tree.bbox_low[e]);
        distsq = __dadd_rn(distsq, dists[e]);
      }
      if (x[e] > tree.bbox_high[e]) {
        dists[e] = sq_diff(x[e], tree.bbox_high[e]);
        distsq = __dadd_rn(distsq, dists[e]);
      }
    }
    // KNNResultSet with capacity 1 (nanoflann.hpp:78-144)
    double best = DBL_MAX;
    unsigned int best_idx = 0;
    bool have = false;
    // searchLevel (nanoflann.hpp:1213-1270), recursion unrolled onto an explicit stack
    int sp = 0;
    stack[0].node = 0;
    stack[0].mind = distsq;
    stack[0].phase = 0;
    while (sp >= 0) {
      Frame &fr = stack[sp];
      const KdNode nd = nodes[fr.node];
      if (fr.phase == 0) {
        if (nd.child1 < 0 && nd.child2 < 0) {
          const int p = nd.a + lane;
          double dist = DBL_MAX;
          unsigned int index = 0;
          if (p < nd.b) {
            index = vind[p];
            dist = DIMT ? nanoflann_l2_fixed<DIMT ? DIMT : 1>(x, cb + (size_t)index * dim) : nanoflann_l2(x, cb + (size_t)index * dim, dim);
          }
          // leaf minimum, first position on ties (lanes are in vind order)
          double m = dist;
          int ml = lane;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {  // all 32 lanes must end with the same (m, ml): they share the walk's state
            const double om = __shfl_xor_sync(0xffffffffu, m, o);
            const int ol = __shfl_xor_sync(0xffffffffu, ml, o);
            if (om < m || (om == m && ol < ml)) {
              m = om;
              ml = ol;
            }
          }
          const unsigned int mi = __shfl_sync(0xffffffffu, index, ml);
          if (nd.b > nd.a && (!have || m < best)) {  // dist < worstDist at leaf entry, then addPoint's strict '>'
            best = m;
            best_idx = mi;
            have = true;
          }
          sp--;
          continue;
        }
        const int feat = nd.a & kKdFeatMask;
        double val;
        if (DIMT) {  // by selection: a dynamic index would push x[] out of registers
          val = 0.0;
#pragma unroll
          for (int e = 0; e < (DIMT ? DIMT : 1); e++) val = (e == feat) ? x[e] : val;
        } else {
          val = x[feat];
        }
        const double diff1 = __dsub_rn(val, nd.divlow);
        const double diff2 = __dsub_rn(val, nd.divhigh);
        int first;
        if (__dadd_rn(diff1, diff2) < 0) {
          first = nd.child1;
          fr.other = nd.child2;
          fr.cut = sq_diff(val, nd.divhigh);
        } else {
          first = nd.child2;
          fr.other = nd.child1;
          fr.cut = sq_diff(val, nd.divlow);
        }
        fr.feat = feat;
        fr.phase = 1;
        const double m = fr.mind;
        sp++;
        stack[sp].node = first;
        stack[sp].mind = m;
        stack[sp].phase = 0;
      } else if (fr.phase == 1) {
        const double dst = dists[fr.feat];
        fr.saved = dst;
        const double m2 = __dsub_rn(__dadd_rn(fr.mind, fr.cut), dst);  // mindistsq + cut_dist - dst
        dists[fr.feat] = fr.cut;
        fr.phase = 2;
        if (m2 <= best) {  // mindistsq*epsError <= worstDist(), epsError == 1
          const int other = fr.other;
          sp++;
          stack[sp].node = other;
          stack[sp].mind = m2;
          stack[sp].phase = 0;
        }
      } else {
        dists[fr.feat] = fr.saved;
        sp--;
      }
    }
    if (lane == 0) {
      const uint32_t old = assign[v] & 0x7fffffffu;
      (result ? result : assign)[v] = best_idx;
      if (old != best_idx) atomicAdd(changed, 1u);
    }
    if (stats) add_query_stats(src, img, base, stats + (size_t)best_idx * (dim + 2), lane);
  }
}

// ------------------------------------------------------------------------------------------------
// accumulate: per-cell integer statistics  {n_k, S_k[d] = sum L, Q_k = sum_d sum L^2}
// ------------------------------------------------------------------------------------------------
// Known dimensions: per-CTA privatised table in shared memory, but the lanes of a warp that hit the SAME
// cell are combined first - __match_any_sync groups them, __reduce_add_sync (REDUX) sums each coordinate
// inside every group at once, and only the group leaders issue shared-memory atomics.  Early split
// levels (K = 1, 2, 4, ... where a warp's 32 vectors fall into a handful of cells) need ~1 atomic per
// vector instead of dim + 2.  MATCH costs one step per distinct value, so this kernel is used for K <= 8 only.
// Cells [k_base, k_base + k_count) only: codebooks whose table exceeds shared memory take blockIdx.y slices.
template <int DIM>
__global__ void __launch_bounds__(1024, 1)
    accumulate_match_kernel(const VecSource src, const uint32_t *__restrict__ assign, const int K, const int k_slice,
                            unsigned long long *__restrict__ stats) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int k_base = blockIdx.y * k_slice;
  const int k_count = min(k_slice, K - k_base);
  unsigned long long *s_q = reinterpret_cast<unsigned long long *>(smem_raw);  // [k_slice]
  int *s_n = reinterpret_cast<int *>(s_q + k_slice);                           // [k_slice]
  int *s_s = s_n + k_slice;                                                    // [k_slice][DIM]
  for (int i = threadIdx.x; i < k_slice; i += blockDim.x) {
    s_q[i] = 0;
    s_n[i] = 0;
  }
  for (int i = threadIdx.x; i < k_slice * DIM; i += blockDim.x) s_s[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned long long n_round = (src.n_local + 31ull) & ~31ull;  // whole warps stay in the loop
  for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_round;
       v += (unsigned long long)gridDim.x * blockDim.x) {
    int a = -1;
    int L[DIM];
    int q = 0;
    if (v < src.n_local) {
      a = (assign ? (int)assign[v] : 0) - k_base;
      if (a < 0 || a >= k_count) a = -1;
    }
    if (a >= 0) {
      gather_lattice<DIM>(src, v, L);
#pragma unroll
      for (int e = 0; e < DIM; e++) q += L[e] * L[e];
    } else {
#pragma unroll
      for (int e = 0; e < DIM; e++) L[e] = 0;
    }
    const unsigned int group = __match_any_sync(0xffffffffu, a);
    const bool leader = a >= 0 && lane == __ffs(group) - 1;
    int *row = s_s + (a >= 0 ? a : 0) * DIM;
    // two coordinates per REDUX: t = L + 128 in [0, 255], 16 bits each (the table holds sums of t, see the flush)
#pragma unroll
    for (int e = 0; e < DIM; e += 2) {
      const unsigned int t0 = a >= 0 ? (unsigned int)(L[e] + 128) : 0u;
      const unsigned int t1 = (e + 1 < DIM && a >= 0) ? (unsigned int)(L[e + 1 < DIM ? e + 1 : e] + 128) : 0u;
      const unsigned int sv = __reduce_add_sync(group, t0 | (t1 << 16));
      if (leader) {
        if (sv & 0xffffu) atomicAdd(row + e, (int)(sv & 0xffffu));
        if (e + 1 < DIM && (sv >> 16)) atomicAdd(row + e + 1, (int)(sv >> 16));
      }
    }
    const unsigned int qs = __reduce_add_sync(group, (unsigned int)q);  // <= 32 * DIM * 128^2: fits
    if (leader) {
      atomicAdd(s_n + a, __popc(group));
      atomicAdd(s_q + a, (unsigned long long)qs);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k_count; i += blockDim.x) {
    if (s_n[i] != 0) {
      unsigned long long *row = stats + (size_t)(k_base + i) * (DIM + 2);
      atomicAdd(row, (unsigned long long)s_n[i]);
      atomicAdd(row + DIM + 1, s_q[i]);
    }
  }
  for (int i = threadIdx.x; i < k_count * DIM; i += blockDim.x) {
    const int k = i / DIM, e = i - k * DIM;
    const int sv = s_s[i] - 128 * s_n[k];  // sum of t -> sum of L
    if (sv != 0) atomicAdd(stats + (size_t)(k_base + k) * (DIM + 2) + 1 + e, (unsigned long long)(long long)sv);
  }
}

// Any dimension: per-CTA privatised table in shared memory (cells [k_base, k_base + k_count) only,
// so codebooks whose table exceeds shared memory are handled by blockIdx.y slices that each re-read
// the 4 + dim bytes per vector), shared-memory atomics while streaming, 64-bit global atomics of
// the non-zero entries at the end.
template <int DIMT>
__global__ void __launch_bounds__(256)
    accumulate_smem_kernel(const VecSource src, const uint32_t *__restrict__ assign, const int K,
                           const int k_slice, unsigned long long *__restrict__ stats) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int dim = DIMT ? DIMT : src.dim;
  const int k_base = blockIdx.y * k_slice;
  const int k_count = min(k_slice, K - k_base);
  unsigned long long *s_q = reinterpret_cast<unsigned long long *>(smem_raw);  // [k_slice]
  int *s_n = reinterpret_cast<int *>(s_q + k_slice);                           // [k_slice]
  int *s_s = s_n + k_slice;                                                    // [k_slice][dim]
  for (int i = threadIdx.x; i < k_slice; i += blockDim.x) {
    s_q[i] = 0;
    s_n[i] = 0;
  }
  for (int i = threadIdx.x; i < k_slice * dim; i += blockDim.x) s_s[i] = 0;
  __syncthreads();
  for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < src.n_local;
       v += (unsigned long long)gridDim.x * blockDim.x) {
    const int a = (assign ? (int)assign[v] : 0) - k_base;
    if (a < 0 || a >= k_count) continue;
    int q = 0;
    int *row = s_s + a * dim;
    if constexpr (DIMT != 0) {
      int Lv[DIMT ? DIMT : 1];
      gather_lattice<DIMT>(src, v, Lv);
#pragma unroll
      for (int e = 0; e < DIMT; e++) {
        q += Lv[e] * Lv[e];
        if (Lv[e] != 0) atomicAdd(row + e, Lv[e]);
      }
    } else {
      unsigned long long base, img;
      vec_base(src, v, base, img);
      for (int e = 0; e < dim; e++) {
        const int L = load_lattice(src, img, base, e);
        q += L * L;
        if (L != 0) atomicAdd(row + e, L);
      }
    }
    atomicAdd(s_n + a, 1);
    atomicAdd(s_q + a, (unsigned long long)q);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k_count; i += blockDim.x) {
    if (s_n[i] != 0) {
      unsigned long long *row = stats + (size_t)(k_base + i) * (dim + 2);
      atomicAdd(row, (unsigned long long)s_n[i]);
      atomicAdd(row + dim + 1, s_q[i]);
    }
  }
  for (int i = threadIdx.x; i < k_count * dim; i += blockDim.x) {
    const int sv = s_s[i];
    if (sv != 0) {
      const int k = i / dim, e = i - k * dim;
      atomicAdd(stats + (size_t)(k_base + k) * (dim + 2) + 1 + e, (unsigned long long)(long long)sv);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// decode: indices + codebook bytes -> RGB bytes, and the report's pixel-domain squared error
// ------------------------------------------------------------------------------------------------
// getImageFromVectors (src/Compressor.cpp:64-92) PUSHES block elements into pixels in (i, j, x, y)
// loop order; on shapes where ySize is not a multiple of h a block overflows in y and its writes
// wrap into the next x line, so a pixel can be written twice and the later write wins.  One thread
// per pixel PULLS instead: it enumerates the (x, y) pairs with x*ySize + y == p that some block
// covers and takes the one the sequential loops would have executed last.
__global__ void __launch_bounds__(256)
    decode_kernel(const DecodeGeom g, const uint8_t *__restrict__ orig, const uint32_t *__restrict__ assign,
                  const uint8_t *__restrict__ cb_bytes, uint8_t *__restrict__ out,
                  unsigned long long *__restrict__ sq_err) {
  unsigned long long err = 0;
  const unsigned long long total = g.n_pixels * (unsigned long long)g.n_images;
  const unsigned long long per_image_vecs = (unsigned long long)g.wB * g.hB;
  const long long y_cover = (long long)g.hB * g.h;  // blocks cover y in [0, y_cover)
  const int dim = 3 * g.w * g.h;
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long img = t / g.n_pixels;
    const long long p = (long long)(t - img * g.n_pixels);
    long long x = p / g.ySize, y = p - x * g.ySize;
    // candidates: (x, y), (x-1, y+ySize), ... while y < y_cover; keep the lexicographically largest (i, j, x, y)
    long long bi = -1, bj = -1, bx = 0, by = 0;
    while (x >= 0 && y < y_cover) {
      const long long i = x / g.w, j = y / g.h;
      if (i > bi || (i == bi && (j > bj || (j == bj && (x > bx || (x == bx && y > by)))))) {
        bi = i; bj = j; bx = x; by = y;
      }
      x -= 1;
      y += g.ySize;
    }
    uint8_t px[3] = {0, 0, 0};  // a pixel nobody writes keeps RGB{} == 0
    if (bi >= 0) {
      const unsigned long long vec = img * per_image_vecs + (unsigned long long)bi * g.hB + (unsigned long long)bj;
      const int e = (int)(((bx - bi * g.w) * g.h + (by - bj * g.h)) * 3);
      const uint8_t *c = cb_bytes + (size_t)assign[vec] * dim + e;
      px[0] = __ldg(c); px[1] = __ldg(c + 1); px[2] = __ldg(c + 2);
    }
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      if (out) out[t * 3 + ch] = px[ch];
      // the report compares the bytes as SIGNED chars (src/Compressor.cpp:141-143)
      const int d = (int)(signed char)__ldg(orig + t * 3 + ch) - (int)(signed char)px[ch];
      err += (unsigned long long)(d * d);
    }
  }
  for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
  if ((threadIdx.x & 31) == 0 && err) atomicAdd(sq_err, err);
}

// ------------------------------------------------------------------------------------------------
// pack_vectors_kernel: getBlocksAsVectorsFromImage (/root/reference/src/Compressor.cpp:31-62) as BYTES
// ------------------------------------------------------------------------------------------------
// One thread per 4-byte word of the dense N x stride copy of the training set: gathers four block elements
// with the reference's layout rule (any shape: padding and the y-overflow wrap are resolved here, once) and
// stores them coalesced.  The reference materialises 240-byte double vectors at this point; this is 12 bytes.
__global__ void __launch_bounds__(256) pack_vectors_kernel(const VecSource src, uint8_t *__restrict__ dense, const int stride) {
  const int words = stride >> 2;
  const unsigned long long total = src.n_local * (unsigned long long)words;
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long v = t / (unsigned int)words;
    const int wd = (int)(t - v * (unsigned int)words);
    unsigned int packed = 0;
    if (src.fast) {
      const signed char *p = fast_vec_ptr(src, v);
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int e = 4 * wd + b;
        if (e < src.dim) packed |= (unsigned int)(unsigned char)__ldg(p + src.elem_off[e]) << (8 * b);
      }
    } else {
      unsigned long long base, img;
      vec_base(src, v, base, img);
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int e = 4 * wd + b;
        if (e < src.dim) packed |= (unsigned int)(unsigned char)load_lattice(src, img, base, e) << (8 * b);
      }
    }
    reinterpret_cast<unsigned int *>(dense)[t] = packed;
  }
}

// ------------------------------------------------------------------------------------------------
// stage_codebook_kernel: FP64 codebook -> what the filters consume, on the device (one thread per row)
// ------------------------------------------------------------------------------------------------
// Lattice coordinates C = 255*c - 128 (SCALED) or c (NORMAL), rounded once to FP32.
//   rows32   k_rows32 rows of `row32` floats: [-2*C_k, |C_k|^2, 0 ..]; rows >= K can never win
//   tc_out   (optional) k_rows_tc rows as three bf16 limbs in the UMMA K-major no-swizzle layout, per N tile
//            of 256 rows, per limb, per 16-wide K block: 8-row x 16-byte core matrices, the two core matrices
//            of a K block 128 B apart, 8-row groups 256 B apart.  Limbs of v: bf16(v), bf16(v - hi),
//            bf16(v - hi - mid) - together 24 mantissa bits.
//   cb_t     the FP64 codebook transposed ([dim][K]) for the brute-force resolver
//   c_max    max_k |C_k| (slightly rounded up), as float bits via atomicMax (non-negative floats order as ints)
__device__ __forceinline__ unsigned short bf16_rn_bits(float f) {
  unsigned int u = __float_as_uint(f);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (unsigned short)(u >> 16);
}
__global__ void __launch_bounds__(128)
    stage_codebook_kernel(const double *__restrict__ cb, const int K, const int k_rows32, const int k_rows_tc,
                          const int dim, const int scaled, float *__restrict__ rows32, const int row32,
                          unsigned char *__restrict__ tc_out, const int kblocks, unsigned int *__restrict__ c_max_bits,
                          double *__restrict__ cb_t) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_rows = k_rows32 > k_rows_tc ? k_rows32 : k_rows_tc;
  if (k >= n_rows) return;
  const size_t block = (size_t)256 * 32;  // bytes of one (N tile, limb, K block)
  const int jt = k >> 8, r = k & 255;
  auto put_tc = [&](int e, double v, bool single) {
    if (!tc_out || k >= k_rows_tc) return;
    const int kb = e >> 4, kk = e & 15;
    double rem = v;
#pragma unroll
    for (int l = 0; l < 3; l++) {
      const unsigned short b = (single && l > 0) ? (unsigned short)0 : bf16_rn_bits((float)rem);
      rem = __dsub_rn(rem, (double)__uint_as_float((unsigned int)b << 16));
      unsigned char *p = tc_out + (size_t)((jt * 3 + l) * kblocks + kb) * block + (size_t)(r >> 3) * 256 +
                         (size_t)(kk >> 3) * 128 + (size_t)(r & 7) * 16 + (size_t)(kk & 7) * 2;
      *reinterpret_cast<unsigned short *>(p) = b;
    }
  };
  double n2 = 0.0;
  for (int e = 0; e < dim; e++) {
    float Cf = 0.f;
    if (k < K) {
      const double c = cb[(size_t)k * dim + e];
      cb_t[(size_t)e * K + k] = c;  // transposed FP64 copy for the brute-force resolver
      Cf = (float)(scaled ? __dsub_rn(__dmul_rn(255.0, c), 128.0) : c);
      n2 = __dadd_rn(n2, __dmul_rn((double)Cf, (double)Cf));
    }
    if (k < k_rows32) rows32[(size_t)k * row32 + e] = -2.0f * Cf;
    put_tc(e, -2.0 * (double)Cf, false);
  }
  const bool pad = k >= K;
  if (k < k_rows32) {
    rows32[(size_t)k * row32 + dim] = pad ? 3.0e38f : (float)n2;
    for (int e = dim + 1; e < row32; e++) rows32[(size_t)k * row32 + e] = 0.f;
  }
  put_tc(dim, pad ? 3.0e38 : n2, pad);
  for (int e = dim + 1; e < kblocks * 16; e++) put_tc(e, 0.0, true);
  if (!pad) atomicMax(c_max_bits, __float_as_uint((float)(sqrt(n2) * 1.000001 + 1e-3)));
}

// ------------------------------------------------------------------------------------------------
// finalize_split_kernel: fixCodeVectors + the two updateDistortion values + the next split, on the device
// ------------------------------------------------------------------------------------------------
// Device twin of qb200_finalize_level (qb200_api.cu) followed by the split of src/Quantizer.cpp:134-138, so
// that the HEAD-schedule train needs no host round trip between levels: from the reduced integer statistics
// it writes the centroids (same IEEE operations as the host version: ((double)S_t / unit) / n, zero vector for
// an empty cell), the next level's codebook (1.2 c | 0.8 c) and {distortion before fix, after fix, dead cells}.
// One block; per-thread partial sums in a fixed order + a fixed shared-memory tree: deterministic.
struct LevelSummary {
  double dist_pre, dist_post;
  unsigned int dead_cells, pad;
  unsigned long long n_total_seen;
};
__global__ void __launch_bounds__(1024)
    finalize_split_kernel(const unsigned long long *__restrict__ stats, const double *__restrict__ cb_pre,
                          const double *__restrict__ exact_state, const int K, const int dim, const int scaled, const double n_total, const double f_up, const double f_dn,
                          double *__restrict__ cb_post, double *__restrict__ cb_next,
                          LevelSummary *__restrict__ summary, unsigned char *__restrict__ exact_next,
                          const unsigned char *__restrict__ small_flag, const double *__restrict__ small_sums) {
  __shared__ double s_pre[1024], s_post[1024];
  __shared__ unsigned int s_dead[1024];
  __shared__ unsigned long long s_n[1024];
  const double unit = scaled ? 255.0 : 1.0;
  double acc_pre = 0.0, acc_post = 0.0;
  unsigned int dead = 0;
  unsigned long long seen = 0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const unsigned long long *row = stats + (size_t)k * (dim + 2);
    const unsigned long long n = row[0];
    seen += n;
    dead += n == 0;
    long long s_all = 0;
    for (int e = 0; e < dim; e++) s_all += (long long)row[1 + e];
    // Q in the colour space's lattice t (SCALED: t = L + 128): sum t^2 = Q_L + 256 S_L + 128^2 dim n
    const double Qt = scaled ? (double)row[dim + 1] + 256.0 * (double)s_all + 16384.0 * (double)dim * (double)n
                             : (double)row[dim + 1];
    // all members the same vector (see qb200_finalize_level): the reference's sum of n equal terms is fl(n * v)
    bool same = scaled && n > 0 && !exact_state;
    if (same) {
      unsigned long long q = 0;
      for (int e = 0; e < dim && same; e++) {
        const long long S = (long long)row[1 + e];
        const long long m = S / (long long)n;
        same = m * (long long)n == S;
        q += (unsigned long long)(m * m);
      }
      same = same && row[dim + 1] == n * q;
    }
    // children of this cell whose centroid the integer path reproduces bit for bit (one repeated vector, or empty);
    // NORMAL sums are integers: always exact
    // small cell: its compensated sums were computed beside the integer ones (qb200_exact.cu, "small cells")
    const bool small = small_flag && small_flag[k] && n > 0 && !exact_state;
    if (exact_next) {
      const unsigned char ex = (!scaled || n == 0 || same || small || exact_state) ? 1 : 0;
      exact_next[k] = ex;
      exact_next[K + k] = ex;
    }
    double st2 = 0.0, cross = 0.0, c2 = 0.0;
    for (int e = 0; e < dim; e++) {
      const long long St = (long long)row[1 + e] + (scaled ? (long long)(128ull * n) : 0ll);
      // exact_state: the reference's compensated sum itself (qb200_exact.cu), divided as in src/Quantizer.cpp:84-85
      const double c = !n ? 0.0
                          : exact_state ? __ddiv_rn(exact_state[((size_t)k * dim + e) * 2], (double)n)
                          : small ? __ddiv_rn(small_sums[(size_t)k * dim + e], (double)n)
                          : same ? __ddiv_rn(__dmul_rn((double)n, __ddiv_rn((double)(St / (long long)n), unit)), (double)n)
                                 : __ddiv_rn(__ddiv_rn((double)St, unit), (double)n);
      cb_post[(size_t)k * dim + e] = c;
      if (cb_next) {
        cb_next[(size_t)k * dim + e] = __dmul_rn(c, f_up);
        cb_next[((size_t)K + k) * dim + e] = __dmul_rn(c, f_dn);
      }
      st2 += (double)St * (double)St;
      if (cb_pre) {
        const double cp = cb_pre[(size_t)k * dim + e];
        cross += cp * (double)St;
        c2 += cp * cp;
      }
    }
    if (n) acc_post += Qt - st2 / (double)n;
    if (cb_pre) acc_pre += Qt - 2.0 * unit * cross + unit * unit * (double)n * c2;
  }
  s_pre[threadIdx.x] = acc_pre;
  s_post[threadIdx.x] = acc_post;
  s_dead[threadIdx.x] = dead;
  s_n[threadIdx.x] = seen;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_pre[threadIdx.x] += s_pre[threadIdx.x + o];
      s_post[threadIdx.x] += s_post[threadIdx.x + o];
      s_dead[threadIdx.x] += s_dead[threadIdx.x + o];
      s_n[threadIdx.x] += s_n[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double denom = unit * unit * n_total * (double)dim;
    summary->dist_pre = s_pre[0] / denom;
    summary->dist_post = s_post[0] / denom;
    summary->dead_cells = s_dead[0];
    summary->pad = 0;
    summary->n_total_seen = s_n[0];
  }
}

// ------------------------------------------------------------------------------------------------
// empty-cell repair (extension, QB200_MODE_FULL_REPAIR): pick one member of each donor cell
// ------------------------------------------------------------------------------------------------
// The README's repair step ("random vector from the area of biggest distortion", README.md:31; the dead helper
// getDistortionInArea, src/Quantizer.cpp:34-44) is not implemented in the reference, so there is no RNG to
// match: the member is the one with the smallest 32-bit hash of (global vector index, seed, round) - a uniform
// choice that does not depend on how the vectors are sharded.  key = hash << 32 | global index, one atomicMin.
__device__ __forceinline__ unsigned int mix_hash(unsigned long long x) {  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return (unsigned int)(x >> 32);
}
__global__ void __launch_bounds__(256)
    pick_members_kernel(const VecSource src, const uint32_t *__restrict__ assign, const int *__restrict__ slot_of_cell,
                        const unsigned long long seed, unsigned long long *__restrict__ keys) {
  for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < src.n_local;
       v += (unsigned long long)gridDim.x * blockDim.x) {
    const int slot = slot_of_cell[assign[v]];
    if (slot < 0) continue;
    const unsigned long long gv = src.first_vec + v;  // index inside the whole (unsharded) training set
    const unsigned long long key = ((unsigned long long)mix_hash(gv * 0x9E3779B97F4A7C15ull + seed) << 32) | (gv & 0xffffffffull);
    atomicMin(keys + slot, key);
  }
}
// out[i*dim + e] = lattice value + 128 of local vector local_idx[i] (or 0 when local_idx[i] is not on this rank)
__global__ void fetch_members_kernel(const VecSource src, const long long *__restrict__ local_idx, const int count,
                                     unsigned long long *__restrict__ out) {
  const int i = blockIdx.x;
  if (i >= count) return;
  const long long v = local_idx[i];
  for (int e = threadIdx.x; e < src.dim; e += blockDim.x) {
    unsigned long long w = 0;
    if (v >= 0) {
      unsigned long long base, img;
      vec_base(src, (unsigned long long)v, base, img);
      w = (unsigned long long)(load_lattice(src, img, base, e) + 128);
    }
    out[(size_t)i * src.dim + e] = w;
  }
}

// ------------------------------------------------------------------------------------------------
// FP32 FMA peak probe (roofline denominator measured in the same run)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) ffma_probe_kernel(float *out, int iters, const float m, const float c) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      a0 = fmaf(a0, m, c);
      a1 = fmaf(a1, m, c);
      a2 = fmaf(a2, m, c);
      a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c);
      a5 = fmaf(a5, m, c);
      a6 = fmaf(a6, m, c);
      a7 = fmaf(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static int g_launch_count = 0;
int launch_count() { return g_launch_count; }
void reset_launch_count() { g_launch_count = 0; }
void count_launch() { g_launch_count++; }

int assign_row_floats(int dim) { return ((dim + 1 + 3) / 4) * 4; }

template <int DIM>
static cudaError_t launch_assign_t(const AssignLaunch &a) {
  using Cfg = AssignCfg<DIM>;
  if (a.k_real <= 16) {  // early split levels: one fused high-occupancy pass
    unsigned long long blocks = (a.src.n_local + 255) / 256;
    const unsigned long long cap = (unsigned long long)a.sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (a.fused_out) *a.fused_out = a.stats != nullptr;
    if (blocks == 0) return cudaSuccess;
    // register accumulators for as many cells as ~128 registers hold ((DIM+1)/2 + 2 words per cell)
    constexpr int kFit = 128 / ((DIM + 1) / 2 + 2);
    // (measured, dim 12: K = 16 needs 173 registers, one CTA per SM, and is slower than the REDUX path: stop at 8)
    constexpr int kCap = kFit >= 8 ? 8 : kFit >= 4 ? 4 : kFit >= 2 ? 2 : 0;
#define QB_SMALLK(KC)                                                                                              \
  small_k_fused_kernel<DIM, KC><<<(unsigned int)blocks, 256, 0, a.stream>>>(a.src, a.cb_rows, a.k_real, a.margin_coef, \
                                                                            a.c_max_ptr, a.assign, a.flag_list,      \
                                                                            a.flag_count, a.stats)
    if (a.stats == nullptr || a.k_real > kCap)
      QB_SMALLK(0);
    else if (a.k_real <= 2 && kCap >= 2)
      QB_SMALLK(kCap >= 2 ? 2 : 0);
    else if (a.k_real <= 4 && kCap >= 4)
      QB_SMALLK(kCap >= 4 ? 4 : 0);
    else
      QB_SMALLK(kCap >= 8 ? 8 : 0);
#undef QB_SMALLK
    g_launch_count++;
    return cudaGetLastError();
  }
  const size_t row_bytes = (size_t)Cfg::ROW * 4;
  const size_t smem_cap = 200 * 1024;
  int k_chunk = a.K;
  if ((size_t)k_chunk * row_bytes > smem_cap) k_chunk = (int)(smem_cap / row_bytes) & ~7;  // even
  size_t smem = (size_t)k_chunk * row_bytes;
  // fused statistics only when the whole codebook is one chunk and the table fits behind it
  const size_t table = ((smem + 127) & ~(size_t)127) - smem + (size_t)a.k_real * (12 + 4 * (size_t)DIM);
  // (measured: below 32 cells the in-kernel atomics of this one-CTA-per-SM kernel cost more than the separate,
  //  high-occupancy accumulate pass)
  const bool fuse = a.stats != nullptr && a.k_real >= 32 && k_chunk == a.K && smem + table <= smem_cap + 16 * 1024;
  if (fuse) smem += table;
  if (a.fused_out) *a.fused_out = fuse;
  auto kernel = a.src.f64 ? assign_kernel<DIM, true> : assign_kernel<DIM, false>;
  {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_cap + 16 * 1024));
    if (e != cudaSuccess) return e;
  }
  const unsigned long long per_tile = (unsigned long long)Cfg::THREADS * Cfg::Q;
  const unsigned long long tiles = (a.src.n_local + per_tile - 1) / per_tile;
  unsigned long long grid = tiles < (unsigned long long)a.sm_count ? tiles : (unsigned long long)a.sm_count;
  if (grid == 0) return cudaSuccess;
  kernel<<<(unsigned int)grid, Cfg::THREADS, smem, a.stream>>>(
      a.src, a.cb_rows, a.K, k_chunk, a.margin_coef, a.c_max_ptr, a.assign, a.flag_list, a.flag_count, tiles,
      fuse ? a.stats : nullptr, a.k_real);
  g_launch_count++;
  return cudaGetLastError();
}

cudaError_t launch_assign(const AssignLaunch &a) {
  if (a.fused_out) *a.fused_out = false;
  // FP64 vectors: the tiled kernel where the dimension has an instance and the codebook is past the fused
  // small-K pass (which accumulates integer statistics), else the generic kernel
  switch (a.src.f64 && a.k_real <= 16 ? -1 : a.src.dim) {
    case 3: return launch_assign_t<3>(a);
    case 6: return launch_assign_t<6>(a);
    case 9: return launch_assign_t<9>(a);
    case 12: return launch_assign_t<12>(a);
    case 24: return launch_assign_t<24>(a);
    case 27: return launch_assign_t<27>(a);
    case 48: return launch_assign_t<48>(a);
    default: break;
  }
  const size_t smem = (size_t)a.src.dim * 128 * 4;
  {
    cudaError_t e = cudaFuncSetAttribute(assign_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kMaxDim * 128 * 4);
    if (e != cudaSuccess) return e;
  }
  unsigned long long blocks = (a.src.n_local + 127) / 128;
  const unsigned long long cap = (unsigned long long)a.sm_count * 2;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) return cudaSuccess;
  assign_generic_kernel<<<(unsigned int)blocks, 128, smem, a.stream>>>(a.src, a.cb_rows, a.K,
                                                                        assign_row_floats(a.src.dim), a.margin_coef,
                                                                        a.c_max_ptr, a.assign, a.flag_list,
                                                                        a.flag_count);
  g_launch_count++;
  return cudaGetLastError();
}

template <int DIM>
static cudaError_t launch_refilter_t(const VecSource &src, const float *cb_rows, int K, float margin_coef, const float *c_max_ptr,
                                     uint32_t *assign, const uint32_t *list_in, const unsigned int *count_in, uint32_t *list_out,
                                     unsigned int *count_out, int sm_count, cudaStream_t stream) {
  using Cfg = AssignCfg<DIM>;
  const size_t row_bytes = (size_t)Cfg::ROW * 4, smem_cap = 200 * 1024;
  int k_chunk = K;
  if ((size_t)k_chunk * row_bytes > smem_cap) k_chunk = (int)(smem_cap / row_bytes) & ~7;
  cudaError_t e = cudaFuncSetAttribute(refilter_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap);
  if (e != cudaSuccess) return e;
  refilter_kernel<DIM><<<(unsigned int)sm_count, Cfg::THREADS, (size_t)k_chunk * row_bytes, stream>>>(
      src, cb_rows, K, k_chunk, margin_coef, c_max_ptr, assign, list_in, count_in, list_out, count_out);
  g_launch_count++;
  return cudaGetLastError();
}

// FP32 re-rank of the queries in list_in (see refilter_kernel).  K: staged rows (even).  Returns cudaErrorNotSupported
// for dimensions without a template instance (the caller then feeds list_in straight to the resolver).
cudaError_t launch_refilter(const VecSource &src, const float *cb_rows, int K, float margin_coef, const float *c_max_ptr,
                            uint32_t *assign, const uint32_t *list_in, const unsigned int *count_in, uint32_t *list_out,
                            unsigned int *count_out, int sm_count, cudaStream_t stream) {
  switch (src.dim) {
#define QB_RF(D) case D: return launch_refilter_t<D>(src, cb_rows, K, margin_coef, c_max_ptr, assign, list_in, count_in, list_out, count_out, sm_count, stream)
    QB_RF(3); QB_RF(6); QB_RF(9); QB_RF(12); QB_RF(24); QB_RF(27); QB_RF(48);
#undef QB_RF
    default: return cudaErrorNotSupported;
  }
}

// assign[v] = result[v] for every flagged query (see resolve_bruteforce_kernel)
__global__ void commit_resolved_kernel(const uint32_t *__restrict__ flag_list, const unsigned int *__restrict__ flag_count,
                                       const uint32_t *__restrict__ result, uint32_t *__restrict__ assign) {
  const unsigned int total = *flag_count;
  for (unsigned int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
    const uint32_t v = flag_list[f];
    assign[v] = result[v];
  }
}

cudaError_t launch_commit_resolved(const uint32_t *flag_list, const unsigned int *flag_count, const uint32_t *result,
                                   uint32_t *assign, int sm_count, cudaStream_t stream) {
  commit_resolved_kernel<<<(unsigned int)sm_count, 256, 0, stream>>>(flag_list, flag_count, result, assign);
  g_launch_count++;
  return cudaGetLastError();
}

cudaError_t launch_resolve(const VecSource &src, int scaled, const double *cb, const double *cbt, int K,
                           const KdDevice &tree,
                           const uint32_t *flag_list, const unsigned int *flag_count, uint32_t *assign,
                           uint32_t *tie_list, unsigned int *tie_count, unsigned int *changed,
                           unsigned long long *stats, uint32_t *result, unsigned int *sensitive, const unsigned char *cv_exact,
                           int tree_robust, int sm_count, cudaStream_t stream) {
  // phase A: brute force, one warp per flagged query (the count is only known on the device)
  const unsigned int blocks_a = (unsigned int)sm_count * 8;
#define QB_RESOLVE_A(CAP, DT)                                                                                        \
  resolve_bruteforce_kernel<CAP, DT><<<blocks_a, 128, 0, stream>>>(src, scaled, cbt, K, tree, flag_list, flag_count, \
                                                                   assign, tie_list, tie_count, changed, stats, result, sensitive, cv_exact, tree_robust)
  switch (src.dim) {
    case 3: QB_RESOLVE_A(3, 3); break;
    case 6: QB_RESOLVE_A(6, 6); break;
    case 9: QB_RESOLVE_A(9, 9); break;
    case 12: QB_RESOLVE_A(12, 12); break;
    case 24: QB_RESOLVE_A(24, 24); break;
    case 27: QB_RESOLVE_A(27, 27); break;
    case 48: QB_RESOLVE_A(48, 48); break;
    default:
      if (src.dim <= 16)
        QB_RESOLVE_A(16, 0);
      else if (src.dim <= 48)
        QB_RESOLVE_A(48, 0);
      else
        QB_RESOLVE_A(kMaxDim, 0);
  }
#undef QB_RESOLVE_A
  g_launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // phase B: the reference's tree walk for the queries phase A left undecided
  const unsigned int blocks = (unsigned int)sm_count * 4;
  const bool deep = tree.depth > 92;
  // tree + codebook in shared memory when they fit (the walk is latency-bound on dependent loads)
  size_t stage = (size_t)K * src.dim * 8 + (size_t)tree.n_nodes * sizeof(KdNode) + (((size_t)K * 4 + 15) & ~(size_t)15);
  if (stage > 160 * 1024 || ((size_t)K * src.dim * 8) % 16 != 0) stage = 0;
  auto go = [&](auto kernel) -> cudaError_t {
    if (stage) {
      cudaError_t e2 = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      if (e2 != cudaSuccess) return e2;
    }
    kernel<<<blocks, 128, stage, stream>>>(src, scaled, cb, K, tree, tie_list, tie_count, assign, changed, stats,
                                           (unsigned int)stage, result);
    return cudaGetLastError();
  };
  switch (src.dim) {
    case 3: e = deep ? go(resolve_kernel<3, kResolveDepthCap, 3>) : go(resolve_kernel<3, 96, 3>); break;
    case 6: e = deep ? go(resolve_kernel<6, kResolveDepthCap, 6>) : go(resolve_kernel<6, 96, 6>); break;
    case 9: e = deep ? go(resolve_kernel<9, kResolveDepthCap, 9>) : go(resolve_kernel<9, 96, 9>); break;
    case 12: e = deep ? go(resolve_kernel<12, kResolveDepthCap, 12>) : go(resolve_kernel<12, 96, 12>); break;
    case 24: e = deep ? go(resolve_kernel<24, kResolveDepthCap, 24>) : go(resolve_kernel<24, 96, 24>); break;
    case 27: e = deep ? go(resolve_kernel<27, kResolveDepthCap, 27>) : go(resolve_kernel<27, 96, 27>); break;
    case 48: e = deep ? go(resolve_kernel<48, kResolveDepthCap, 48>) : go(resolve_kernel<48, 96, 48>); break;
    default:
      if (src.dim <= 16)
        e = deep ? go(resolve_kernel<16, kResolveDepthCap, 0>) : go(resolve_kernel<16, 96, 0>);
      else if (src.dim <= 48)
        e = deep ? go(resolve_kernel<48, kResolveDepthCap, 0>) : go(resolve_kernel<48, 96, 0>);
      else
        e = deep ? go(resolve_kernel<kMaxDim, kResolveDepthCap, 0>) : go(resolve_kernel<kMaxDim, 96, 0>);
  }
  g_launch_count++;
  return e;
}

template <int DIM>
static cudaError_t launch_acc_match(const VecSource &src, const uint32_t *assign, int K, int k_slice, int slices,
                                    size_t smem, unsigned long long *stats, int sm_count, cudaStream_t stream) {
  cudaError_t err = cudaFuncSetAttribute(accumulate_match_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (err != cudaSuccess) return err;
  unsigned long long blocks = (src.n_local + 1023) / 1024;
  if (blocks > (unsigned long long)sm_count) blocks = (unsigned long long)sm_count;
  if (blocks == 0) return cudaSuccess;
  dim3 grid((unsigned int)blocks, (unsigned int)slices);
  accumulate_match_kernel<DIM><<<grid, 1024, smem, stream>>>(src, assign, K, k_slice, stats);
  g_launch_count++;
  return cudaGetLastError();
}

cudaError_t launch_accumulate(const VecSource &src, const uint32_t *assign, int K, unsigned long long *stats,
                              int sm_count, cudaStream_t stream) {
  if (assign == nullptr && K != 1) return cudaErrorInvalidValue;
  const int dim = src.dim;
  const size_t per_cell = 8 + 4 + 4 * (size_t)dim;
  const size_t smem_cap = 200 * 1024;
  int k_slice = K;
  if ((size_t)k_slice * per_cell > smem_cap) k_slice = (int)(smem_cap / per_cell);
  const int slices = (K + k_slice - 1) / k_slice;
  const size_t smem = (size_t)k_slice * per_cell;
  if (K <= 8) {  // few cells: combine equal cells inside each warp first (match + redux), see accumulate_match_kernel
    switch (dim) {
      case 3: return launch_acc_match<3>(src, assign, K, k_slice, slices, smem, stats, sm_count, stream);
      case 6: return launch_acc_match<6>(src, assign, K, k_slice, slices, smem, stats, sm_count, stream);
      case 9: return launch_acc_match<9>(src, assign, K, k_slice, slices, smem, stats, sm_count, stream);
      case 12: return launch_acc_match<12>(src, assign, K, k_slice, slices, smem, stats, sm_count, stream);
      case 24: return launch_acc_match<24>(src, assign, K, k_slice, slices, smem, stats, sm_count, stream);
      case 27: return launch_acc_match<27>(src, assign, K, k_slice, slices, smem, stats, sm_count, stream);
      case 48: return launch_acc_match<48>(src, assign, K, k_slice, slices, smem, stats, sm_count, stream);
      default: break;
    }
  }
  auto kernel = accumulate_smem_kernel<0>;
  switch (dim) {
    case 3: kernel = accumulate_smem_kernel<3>; break;
    case 6: kernel = accumulate_smem_kernel<6>; break;
    case 9: kernel = accumulate_smem_kernel<9>; break;
    case 12: kernel = accumulate_smem_kernel<12>; break;
    case 24: kernel = accumulate_smem_kernel<24>; break;
    case 27: kernel = accumulate_smem_kernel<27>; break;
    case 48: kernel = accumulate_smem_kernel<48>; break;
    default: break;
  }
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap);
  if (err != cudaSuccess) return err;
  unsigned long long blocks = (src.n_local + 255) / 256;
  unsigned long long per_sm = (200 * 1024) / (smem + 1024);
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  unsigned long long cap = (unsigned long long)sm_count * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) return cudaSuccess;
  dim3 grid((unsigned int)blocks, (unsigned int)slices);
  kernel<<<grid, 256, smem, stream>>>(src, assign, K, k_slice, stats);
  g_launch_count++;
  return cudaGetLastError();
}

cudaError_t launch_decode(const DecodeGeom &g, const uint8_t *orig, const uint32_t *assign, const uint8_t *cb_bytes,
                          uint8_t *out, unsigned long long *sq_err, int sm_count, cudaStream_t stream) {
  const unsigned long long total = g.n_pixels * (unsigned long long)g.n_images;
  unsigned long long blocks = (total + 255) / 256;
  const unsigned long long cap = (unsigned long long)sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) return cudaSuccess;
  decode_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(g, orig, assign, cb_bytes, out, sq_err);
  g_launch_count++;
  return cudaGetLastError();
}

cudaError_t launch_pack_vectors(const VecSource &src, uint8_t *dense, int stride, int sm_count, cudaStream_t stream) {
  const unsigned long long total = src.n_local * (unsigned long long)(stride / 4);
  unsigned long long blocks = (total + 255) / 256;
  const unsigned long long cap = (unsigned long long)sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) return cudaSuccess;
  pack_vectors_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(src, dense, stride);
  g_launch_count++;
  return cudaGetLastError();
}

cudaError_t launch_stage_codebook(const double *cb, int K, int k_rows32, int k_rows_tc, int dim, int scaled,
                                  float *rows32, unsigned char *tc_out, float *c_max, double *cb_t,
                                  cudaStream_t stream) {
  const int n_rows = k_rows32 > k_rows_tc ? k_rows32 : k_rows_tc;
  if (n_rows == 0) return cudaSuccess;
  stage_codebook_kernel<<<(n_rows + 127) / 128, 128, 0, stream>>>(cb, K, k_rows32, k_rows_tc, dim, scaled, rows32,
                                                                   assign_row_floats(dim), tc_out, tc_kblocks(dim),
                                                                   reinterpret_cast<unsigned int *>(c_max), cb_t);
  g_launch_count++;
  return cudaGetLastError();
}

cudaError_t launch_finalize_split(const unsigned long long *stats, const double *cb_pre, const double *exact_state, int K,
                                  int dim, int scaled, double n_total, double f_up, double f_dn, double *cb_post,
                                  double *cb_next, void *summary, unsigned char *exact_next, const unsigned char *small_flag,
                                  const double *small_sums, cudaStream_t stream) {
  finalize_split_kernel<<<1, 1024, 0, stream>>>(stats, cb_pre, exact_state, K, dim, scaled, n_total, f_up, f_dn, cb_post,
                                                cb_next, reinterpret_cast<LevelSummary *>(summary), exact_next, small_flag, small_sums);
  g_launch_count++;
  return cudaGetLastError();
}

cudaError_t launch_pick_members(const VecSource &src, const uint32_t *assign, const int *slot_of_cell,
                                unsigned long long seed, unsigned long long *keys, int sm_count, cudaStream_t stream) {
  unsigned long long blocks = (src.n_local + 255) / 256;
  const unsigned long long cap = (unsigned long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) return cudaSuccess;
  pick_members_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(src, assign, slot_of_cell, seed, keys);
  g_launch_count++;
  return cudaGetLastError();
}

// One thread per 32-bit word of the stream: word w holds stream bits [32w, 32w + 32).
__global__ void pack_indices_kernel(const uint32_t *__restrict__ assign, const unsigned long long n, const int bits,
                                    uint32_t *__restrict__ out, const unsigned long long out_words) {
  for (unsigned long long w = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; w < out_words;
       w += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long lo = 32ull * w, hi = lo + 32ull;
    uint32_t word = 0;
    for (unsigned long long i = lo / (unsigned)bits; i < n && i * (unsigned)bits < hi; i++) {
      const unsigned long long pos = i * (unsigned)bits;
      const uint32_t a = __ldg(assign + i);
      word |= pos >= lo ? a << (unsigned)(pos - lo) : a >> (unsigned)(lo - pos);
    }
    out[w] = word;
  }
}

cudaError_t launch_pack_indices(const uint32_t *assign, unsigned long long n, int bits, uint32_t *out,
                                unsigned long long out_words, int sm_count, cudaStream_t stream) {
  if (out_words == 0) return cudaSuccess;
  unsigned long long blocks = (out_words + 255) / 256;
  if (blocks > (unsigned long long)sm_count * 16) blocks = (unsigned long long)sm_count * 16;
  pack_indices_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(assign, n, bits, out, out_words);
  g_launch_count++;
  return cudaGetLastError();
}

cudaError_t launch_fetch_members(const VecSource &src, const long long *local_idx, int count, unsigned long long *out,
                                 cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  fetch_members_kernel<<<count, 64, 0, stream>>>(src, local_idx, count, out);
  g_launch_count++;
  return cudaGetLastError();
}

cudaError_t launch_ffma_probe(float *out, int blocks, int iters, float m, float c, cudaStream_t stream) {
  ffma_probe_kernel<<<blocks, 512, 0, stream>>>(out, iters, m, c);
  g_launch_count++;
  return cudaGetLastError();
}

}  // namespace qb
