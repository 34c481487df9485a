// assign_tc_kernel: the nearest-codevector FILTER on the 5th-generation tensor cores (tcgen05).
//
// Replaces the distance loop of Solution::assignCodeVectors (/root/reference/src/Quantizer.cpp:24-32,
// i.e. KDTree::nearestNeighbour per vector, src/KDTree.cpp:20-29) for codebooks of 16 or more
// entries.  Like the CUDA-core assign_kernel it only DECIDES queries whose two best scores are
// further apart than a bound on its own rounding error; the rest go to the exact FP64 resolver.
//
// Score  s_k = |C_k|^2 - 2 <X, C_k>  as a 128 x N x 16*KB GEMM per tile:
//   A (128 queries x 16*KB, bf16)  row = [x_0 .. x_dim-1, 1, 0 ..]    lattice bytes are exact in bf16
//   B (N codevectors x 16*KB, bf16) row = limb_l([-2 C_k, |C_k|^2, 0 ..]),  l = 0,1,2
// -2C and |C|^2 are split on the host into three bf16 limbs (hi + mid + lo carries 24 mantissa bits);
// the three limbs are three accumulating MMAs into the same FP32 accumulator in tensor memory.
// All products are exact in FP32; the error is limb truncation + FP32 accumulation of 3*(dim+1) terms.
//
// Roles inside the one persistent CTA per SM (416 threads):
//   warp 0        barrier setup, TMEM allocation, one lane issues TMA staging of the codebook and all MMAs
//   warps 1-4     "service": thread r owns query r of a tile - gathers its bytes, writes the bf16 A row,
//                 and after the epilogue merges/finalises the result (margin test, index, flag list)
//   warps 5-12    epilogue: two warps per TMEM lane quarter, each scanning half of the N columns with
//                 tcgen05.ld and tracking (best, second, chunk-of-8 index) in registers with min/max ops:
//                 2.75 alu operations per distance evaluation - this, not the tensor pipe, is the bound
// Pipelines: A tile double-buffered in shared memory (a_full/a_empty), accumulator double-buffered in
// TMEM (2 x 256 columns; tmem_full/tmem_empty), per-tile results double-buffered (res_full/res_empty).
//
// The member of the winning chunk of 8 is identified by the service thread with eight FP32 scores
// from the FP32 rows (the same rows the CUDA-core kernel uses): a query that passes the margin test has
// its best score separated from every other score of the chunk by more than the tensor-core bound, which
// is itself more than twice the FP32 bound, so the FP32 argmin over the chunk is the same codevector.
#include <cfloat>

#include "qb200_launch.hpp"
#include "qb200_ptx.cuh"

namespace qb {

namespace {

constexpr int kTcThreads = 32 * 13;
constexpr int kTileQ = 128;  // queries per tile = MMA M = TMEM lanes

template <int DIM>
struct TcCfg {
  static constexpr int KB = (DIM + 1 + 15) / 16;  // 16-wide K blocks per limb
  static constexpr int A_BYTES = KB * 4096;       // one A tile: KB blocks of 128 rows x 32 B
  static constexpr int ROW_BYTES = KB * 3 * 32;   // staged bytes per codevector (3 limbs)
  static constexpr int ROW32 = ((DIM + 1 + 3) / 4) * 4;
};

struct TcShared {
  uint64_t b_full, a_full[2], a_empty[2], tmem_full[2], tmem_empty[2], res_full[2], res_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  float res_best[2][2][kTileQ], res_second[2][2][kTileQ];
  int res_chunk[2][2][kTileQ];
};

// top-2 of eight scores merged into the running (best, second); `chunk` remembers which group of 8 held the best
__device__ __forceinline__ void top2_update8(const float *a, int cid, float &best, float &second, int &chunk) {
  const float l01 = fminf(a[0], a[1]), h01 = fmaxf(a[0], a[1]);
  const float l23 = fminf(a[2], a[3]), h23 = fmaxf(a[2], a[3]);
  const float l45 = fminf(a[4], a[5]), h45 = fmaxf(a[4], a[5]);
  const float l67 = fminf(a[6], a[7]), h67 = fmaxf(a[6], a[7]);
  const float m1a = fminf(l01, l23), m2a = fmin3(fmaxf(l01, l23), h01, h23);
  const float m1b = fminf(l45, l67), m2b = fmin3(fmaxf(l45, l67), h45, h67);
  const float m1 = fminf(m1a, m1b), m2 = fmin3(fmaxf(m1a, m1b), m2a, m2b);
  second = fmin3(second, m2, fmaxf(best, m1));
  chunk = m1 < best ? cid : chunk;
  best = fminf(best, m1);
}

template <int DIM>
__global__ void __launch_bounds__(kTcThreads, 1)
    assign_tc_kernel(const VecSource src, const unsigned char *__restrict__ b_staged, const float *__restrict__ rows32,
                     const int k_rows, const int k_base, const int n_tile, const int first_pass, const int last_pass,
                     const float margin_coef, const float c_max_norm, float *__restrict__ state,
                     uint32_t *__restrict__ assign, uint32_t *__restrict__ flag_list,
                     unsigned int *__restrict__ flag_count, const unsigned long long tiles) {
  using Cfg = TcCfg<DIM>;
  constexpr int KB = Cfg::KB;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // [ B chunk | A tile 0 | A tile 1 | TcShared ]
  const uint32_t b_bytes = (uint32_t)k_rows * Cfg::ROW_BYTES;
  unsigned char *s_b = smem_raw;
  unsigned char *s_a = smem_raw + ((b_bytes + 1023u) & ~1023u);
  TcShared &sh = *reinterpret_cast<TcShared *>(s_a + 2 * Cfg::A_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles_n = k_rows / n_tile;  // N tiles per query tile

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&sh.b_full, 1);
      for (int i = 0; i < 2; i++) {
        mbar_init(&sh.a_full[i], kTileQ);
        mbar_init(&sh.a_empty[i], 1);
        mbar_init(&sh.tmem_full[i], 1);
        mbar_init(&sh.tmem_empty[i], 256);
        mbar_init(&sh.res_full[i], 256);
        mbar_init(&sh.res_empty[i], kTileQ);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    tmem_alloc(&sh.tmem_base, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_base;

  if (warp == 0) {
    // =========================== TMA staging + MMA issue (one lane) ===========================
    if (lane == 0) {
      mbar_expect_tx(&sh.b_full, b_bytes);
      for (uint32_t off = 0; off < b_bytes; off += 32768u)
        tma_load_1d(s_b + off, b_staged + off, min(32768u, b_bytes - off), &sh.b_full);
      mbar_wait_bounded(&sh.b_full, 0, 1);
      const uint32_t idesc = umma_idesc_bf16_f32(n_tile);
      const uint32_t a_addr = smem_u32(s_a), b_addr = smem_u32(s_b);
      const uint32_t b_tile_bytes = (uint32_t)n_tile * 32u;  // one (limb, K block) of one N tile
      unsigned int it = 0, tile_seq = 0;
      for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, tile_seq++) {
        const int ab = tile_seq & 1;
        mbar_wait_bounded(&sh.a_full[ab], (tile_seq >> 1) & 1, 2);
        tc_fence_after();
        for (int jt = 0; jt < n_tiles_n; jt++, it++) {
          const int buf = it & 1;
          mbar_wait_bounded(&sh.tmem_empty[buf], ((it >> 1) & 1) ^ 1, 3);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256u;
#pragma unroll
          for (int l = 0; l < 3; l++) {
#pragma unroll
            for (int kb = 0; kb < KB; kb++) {
              const uint64_t a_desc = umma_smem_desc(a_addr + ab * Cfg::A_BYTES + kb * 4096, 128, 256);
              const uint64_t b_desc =
                  umma_smem_desc(b_addr + (uint32_t)((jt * 3 + l) * KB + kb) * b_tile_bytes, 128, 256);
              umma_bf16(d_tmem, a_desc, b_desc, idesc, (l | kb) != 0);
            }
          }
          umma_commit(&sh.tmem_full[buf]);
        }
        umma_commit(&sh.a_empty[ab]);
      }
    }
  } else if (warp <= 4) {
    // =========================== service: produce A rows, finalise results ===========================
    const int r = (warp - 1) * 32 + lane;  // query row inside a tile
    const uint32_t row_off = (uint32_t)(r >> 3) * 256u + (uint32_t)(r & 7) * 16u;
    float x_cur[DIM], x_next[DIM];
    bool live_cur = false, live_next = false;
    auto gather = [&](unsigned long long tile, float *x, bool &live) {
      const unsigned long long v = tile * kTileQ + r;
      live = v < src.n_local;
      if (live) {
        unsigned long long base, img;
        vec_base(src, v, base, img);
#pragma unroll
        for (int e = 0; e < DIM; e++) x[e] = (float)load_lattice(src, img, base, e);
      } else {
#pragma unroll
        for (int e = 0; e < DIM; e++) x[e] = 0.f;
      }
    };
    auto produce = [&](unsigned int tile_seq, const float *x) {
      const int ab = tile_seq & 1;
      if (tile_seq >= 2) mbar_wait_bounded(&sh.a_empty[ab], ((tile_seq >> 1) & 1) ^ 1, 4);
      unsigned char *a_tile = s_a + ab * Cfg::A_BYTES;
#pragma unroll
      for (int kb = 0; kb < KB; kb++) {
        uint32_t w[8];  // 16 bf16: small integers are exact in bf16 = upper half of their fp32 pattern
#pragma unroll
        for (int p = 0; p < 8; p++) {
          const int e0 = kb * 16 + 2 * p, e1 = e0 + 1;
          const uint32_t lo = e0 < DIM ? (__float_as_uint(x[e0 < DIM ? e0 : 0]) >> 16) : (e0 == DIM ? 0x3F80u : 0u);
          const uint32_t hi = e1 < DIM ? (__float_as_uint(x[e1 < DIM ? e1 : 0]) >> 16) : (e1 == DIM ? 0x3F80u : 0u);
          w[p] = lo | (hi << 16);
        }
        uint4 *dst = reinterpret_cast<uint4 *>(a_tile + kb * 4096 + row_off);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);  // k 0..7 of this block
        dst[8] = make_uint4(w[4], w[5], w[6], w[7]);  // k 8..15: next core matrix (+128 B)
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(&sh.a_full[ab]);
    };

    unsigned int tile_seq = 0;
    unsigned long long tile = blockIdx.x;
    if (tile < tiles) {
      gather(tile, x_cur, live_cur);
      produce(0, x_cur);
    }
    for (; tile < tiles; tile += gridDim.x, tile_seq++) {
      const unsigned long long next = tile + gridDim.x;
      if (next < tiles) {
        gather(next, x_next, live_next);
        produce(tile_seq + 1, x_next);
      }
      // results of this tile
      const int rb = tile_seq & 1;
      mbar_wait_bounded(&sh.res_full[rb], (tile_seq >> 1) & 1, 5);
      const float b0 = sh.res_best[rb][0][r], b1 = sh.res_best[rb][1][r];
      const float s0 = sh.res_second[rb][0][r], s1 = sh.res_second[rb][1][r];
      const int c0 = sh.res_chunk[rb][0][r], c1 = sh.res_chunk[rb][1][r];
      mbar_arrive(&sh.res_empty[rb]);
      const float best = fminf(b0, b1);
      const float second = fmin3(fmaxf(b0, b1), s0, s1);
      const int chunk = b1 < b0 ? c1 : c0;
      const unsigned long long v = tile * kTileQ + r;
      bool flag = false;
      if (!last_pass) {
        if (live_cur) {
          state[v * 3 + 0] = best;
          state[v * 3 + 1] = second;
          state[v * 3 + 2] = __int_as_float(chunk);
        }
      } else {
        float xn = 0.f;
#pragma unroll
        for (int e = 0; e < DIM; e++) xn = fmaf(x_cur[e], x_cur[e], xn);
        const float rr = sqrtf(xn) + c_max_norm;
        flag = live_cur && !((second - best) > margin_coef * rr * rr);
        int bidx = chunk * 8;
        if (live_cur && !flag) {
          float sb = FLT_MAX;
          const float4 *rows = reinterpret_cast<const float4 *>(rows32 + (size_t)chunk * 8 * Cfg::ROW32);
#pragma unroll
          for (int c = 0; c < 8; c++) {
            float cr[Cfg::ROW32];
#pragma unroll
            for (int q4 = 0; q4 < Cfg::ROW32 / 4; q4++) {
              const float4 t = __ldg(rows + c * (Cfg::ROW32 / 4) + q4);
              cr[4 * q4] = t.x; cr[4 * q4 + 1] = t.y; cr[4 * q4 + 2] = t.z; cr[4 * q4 + 3] = t.w;
            }
            float s = cr[DIM];
#pragma unroll
            for (int e = 0; e < DIM; e++) s = fmaf(x_cur[e], cr[e], s);
            if (s < sb) {
              sb = s;
              bidx = chunk * 8 + c;
            }
          }
        }
        if (live_cur) assign[v] = (uint32_t)bidx;
      }
      const unsigned int m = __ballot_sync(0xffffffffu, flag);
      if (m) {
        const int leader = __ffs(m) - 1;
        unsigned int basepos = 0;
        if (lane == leader) basepos = atomicAdd(flag_count, (unsigned int)__popc(m));
        basepos = __shfl_sync(0xffffffffu, basepos, leader);
        if (flag) flag_list[basepos + __popc(m & ((1u << lane) - 1u))] = (uint32_t)v;
      }
#pragma unroll
      for (int e = 0; e < DIM; e++) x_cur[e] = x_next[e];
      live_cur = live_next;
    }
  } else {
    // =========================== epilogue: TMEM -> registers -> running top-2 ===========================
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, 32*quarter + 32) are the ones this warp may read
    const int half = (warp - 5) >> 2;      // which half of the N columns
    const int r = quarter * 32 + lane;     // query row = TMEM lane
    const int half_cols = n_tile >> 1;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    unsigned int it = 0, tile_seq = 0;
    for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, tile_seq++) {
      float best = FLT_MAX, second = FLT_MAX;
      int chunk = k_base >> 3;
      if (!first_pass && half == 0) {
        const unsigned long long v = tile * kTileQ + r;
        if (v < src.n_local) {
          best = state[v * 3 + 0];
          second = state[v * 3 + 1];
          chunk = __float_as_int(state[v * 3 + 2]);
        }
      }
      // Flat sequence of column chunks (<= 32 columns each) over the N tiles of this query tile, software
      // pipelined over two register buffers: the tcgen05.ld of chunk g+1 is in flight while chunk g is reduced.
      const int col0 = half * half_cols;
      const int ncol = min(32, half_cols);
      const int ch_shift = half_cols >= 128 ? 2 : (half_cols >= 64 ? 1 : 0);  // chunks per N tile = 1 << ch_shift
      const int total = n_tiles_n << ch_shift;
      auto issue = [&](int g, float *v) {
        const int jt = g >> ch_shift, c = (g & ((1 << ch_shift) - 1)) * 32;
        const unsigned int itg = it + jt;
        const int buf = itg & 1;
        if (c == 0) {
          mbar_wait_bounded(&sh.tmem_full[buf], (itg >> 1) & 1, 6);
          tc_fence_after();
        }
        const uint32_t taddr = lane_addr + (uint32_t)(buf * 256 + col0 + c);
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (q * 8 < ncol) tmem_ld8(taddr + q * 8, v + q * 8);
      };
      auto landed = [&](int g, float *v) {  // after this the chunk is in registers
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; q++) tmem_ld_pin8(v + q * 8);
        if ((g & ((1 << ch_shift) - 1)) == (1 << ch_shift) - 1) {  // last chunk of its N tile: hand the buffer back
          tc_fence_before();
          mbar_arrive(&sh.tmem_empty[(it + (g >> ch_shift)) & 1]);
        }
      };
      auto reduce = [&](int g, const float *v) {
        const int cid = (k_base + (g >> ch_shift) * n_tile + col0 + (g & ((1 << ch_shift) - 1)) * 32) >> 3;
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (q * 8 < ncol) top2_update8(v + q * 8, cid + q, best, second, chunk);
      };
      float va[32], vb[32];
      issue(0, va);
      for (int g = 0; g < total; g += 2) {
        landed(g, va);
        if (g + 1 < total) issue(g + 1, vb);
        reduce(g, va);
        if (g + 1 < total) {
          landed(g + 1, vb);
          if (g + 2 < total) issue(g + 2, va);
          reduce(g + 1, vb);
        }
      }
      it += n_tiles_n;
      const int rb = tile_seq & 1;
      if (tile_seq >= 2) mbar_wait_bounded(&sh.res_empty[rb], ((tile_seq >> 1) & 1) ^ 1, 7);
      sh.res_best[rb][half][r] = best;
      sh.res_second[rb][half][r] = second;
      sh.res_chunk[rb][half][r] = chunk;
      mbar_arrive(&sh.res_full[rb]);  // arrive has release semantics: the stores above are visible to the waiter
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
int tc_kblocks(int dim) { return (dim + 1 + 15) / 16; }
size_t tc_row_bytes(int dim) { return (size_t)tc_kblocks(dim) * 3 * 32; }
bool tc_supported(int dim, int K) {
  switch (dim) {
    case 3: case 6: case 9: case 12: case 24: case 27: case 48: return K >= 16;
    default: return false;
  }
}
int tc_padded_rows(int K) { return K < 256 ? ((K + 15) / 16) * 16 : ((K + 255) / 256) * 256; }
int tc_n_tile(int K) { return K < 256 ? ((K + 15) / 16) * 16 : 256; }
// Rows per launch: what fits in shared memory next to the two A tiles, a multiple of the N tile.
int tc_chunk_rows(int dim, int K) {
  const int kp = tc_padded_rows(K), nt = tc_n_tile(K);
  const size_t budget = 200 * 1024 - 2 * (size_t)tc_kblocks(dim) * 4096;
  int rows = (int)(budget / tc_row_bytes(dim));
  rows = (rows / nt) * nt;
  return rows < kp ? rows : kp;
}

template <int DIM>
static cudaError_t launch_tc_t(const AssignTcLaunch &a) {
  using Cfg = TcCfg<DIM>;
  const int kp = tc_padded_rows(a.K), nt = tc_n_tile(a.K), chunk = tc_chunk_rows(DIM, a.K);
  const unsigned long long tiles = (a.src.n_local + kTileQ - 1) / kTileQ;
  if (tiles == 0) return cudaSuccess;
  const size_t smem_max = (((size_t)chunk * Cfg::ROW_BYTES + 1023) & ~(size_t)1023) + 2 * Cfg::A_BYTES + sizeof(TcShared) + 1024;
  cudaError_t e = cudaFuncSetAttribute(assign_tc_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
  if (e != cudaSuccess) return e;
  const unsigned int grid = (unsigned int)(tiles < (unsigned long long)a.sm_count ? tiles : a.sm_count);
  for (int k0 = 0; k0 < kp; k0 += chunk) {
    const int rows = kp - k0 < chunk ? kp - k0 : chunk;
    assign_tc_kernel<DIM><<<grid, kTcThreads, smem_max, a.stream>>>(
        a.src, a.b_staged + (size_t)k0 * Cfg::ROW_BYTES, a.rows32, rows, k0, nt, k0 == 0, k0 + rows >= kp, a.margin_coef,
        a.c_max_norm, a.state, a.assign, a.flag_list, a.flag_count, tiles);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_assign_tc(const AssignTcLaunch &a) {
  switch (a.src.dim) {
    case 3: return launch_tc_t<3>(a);
    case 6: return launch_tc_t<6>(a);
    case 9: return launch_tc_t<9>(a);
    case 12: return launch_tc_t<12>(a);
    case 24: return launch_tc_t<24>(a);
    case 27: return launch_tc_t<27>(a);
    case 48: return launch_tc_t<48>(a);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace qb
