// assign_tc_kernel: the nearest-codevector FILTER on the 5th-generation tensor cores (tcgen05).
//
// Replaces the distance loop of Solution::assignCodeVectors (/root/reference/src/Quantizer.cpp:24-32,
// i.e. KDTree::nearestNeighbour per vector, src/KDTree.cpp:20-29) for codebooks of 64 or more
// entries (padded to a multiple of 256 with rows that cannot win).  Like the CUDA-core assign_kernel it only DECIDES queries whose two best scores are
// further apart than a bound on its own rounding error; the rest go to the exact FP64 resolver.
//
// Score  s_k = |C_k|^2 - 2 <X, C_k>  as a 128 x N x 16*KB GEMM per tile:
//   A (128 queries x 16*KB, bf16)  row = [x_0 .. x_dim-1, 1, 0 ..]    lattice bytes are exact in bf16
//   B (N codevectors x 16*KB, bf16) row = limb_l([-2 C_k, |C_k|^2, 0 ..]),  l = 0,1,2
// -2C and |C|^2 are split on the host into three bf16 limbs (hi + mid + lo carries 24 mantissa bits);
// the three limbs are three accumulating MMAs into the same FP32 accumulator in tensor memory.
// All products are exact in FP32; the error is limb truncation + FP32 accumulation of 3*(dim+1) terms.
//
// Roles inside the one persistent CTA per SM (544 threads with one producer group):
//   warp 0        barrier setup, TMEM allocation, one lane issues TMA staging of the codebook and all MMAs
//   warps 1-4     producers (kProdGroups groups of four): thread r gathers the bytes of query r
//                 of a tile and writes its bf16 A row
//   next 4 warps  mergers: thread r merges the two column halves of query r and stores its record
//                 (best, second, chunk) - tc_finalize_kernel turns records into indices and flags
//   last 8 warps  epilogue: two warps per TMEM lane quarter, each scanning half of the N columns with
//                 tcgen05.ld.  Only CHUNK MINIMA are tracked: every 8 scores are reduced to their minimum m
//                 (3 FMNMX3 + 1 FMNMX) and (best, runner-up, chunk index) run over the chunk minima
//                 (runner = min(runner, max(best, m)); chunk = m < best ? id : chunk; best = min(best, m)):
//                 9 alu operations per 8 distance evaluations (1.125 each; the round-1 kernel tracked the true
//                 second-best score at 2.75 each and was bound by exactly that).  The true second-best score is
//                 either the runner-up chunk minimum or the second-best score INSIDE the winning chunk; the
//                 latter is found by the finalise kernel, which rescans the winning chunk in FP32 anyway.
// Pipelines: ring of 4 A tiles in shared memory (a_full/a_empty), accumulator double-buffered in
// TMEM (2 x 256 columns; tmem_full/tmem_empty), ring of 4 per-tile result records (res_full/res_empty).
//
// The member of the winning chunk of 8 is identified by tc_finalize_kernel with eight FP32 scores
// from the FP32 rows (the same rows the CUDA-core kernel uses).  A query is DECIDED when both gaps exceed the
// tensor-core margin: (runner-up chunk minimum - best chunk minimum), tensor-core scores, and (second - best)
// of the eight FP32 scores of the winning chunk.  The margin is three times the tensor-core error bound, which
// is itself more than twice the FP32 bound, so in both comparisons the winner beats everything else by more
// than the rounding error of either side: the FP32 argmin over the chunk is the exact nearest codevector.
#include <cfloat>
#include <cstdlib>

#include "qb200_launch.hpp"
#include "qb200_ptx.cuh"

namespace qb {

namespace {

constexpr int kProdGroups = 1;          // producer groups of 4 warps taking alternate tiles (2 measured slower: register cap)
// MMA + producer + merger + epilogue warps; the epilogue has S warps per TMEM lane quarter, each scanning 1/S of the columns
constexpr int tc_threads(int S) { return 32 * (1 + 4 * kProdGroups + 4 + 4 * S); }

constexpr int kAStages = 4;            // A tiles in flight (shared memory ring)
constexpr int kResStages = 4;          // per-tile result records in flight
// Bit 31 of a provisional index: the query is on the flag list.  The exact resolver always rewrites such entries;
// until then the statistics pass, which may run concurrently with it, skips them (its cell test fails).
constexpr uint32_t kUndecided = 0x80000000u;
constexpr int kTileQ = 128;  // queries per tile = MMA M = TMEM lanes
constexpr int kTileN = 256;  // codevectors per MMA = accumulator columns per TMEM buffer

template <int DIM>
struct TcCfg {
  static constexpr int KB = (DIM + 1 + 15) / 16;  // 16-wide K blocks per limb
  static constexpr int A_BYTES = KB * 4096;       // one A tile: KB blocks of 128 rows x 32 B
  static constexpr int ROW_BYTES = KB * 3 * 32;   // staged bytes per codevector (3 limbs)
  static constexpr int ROW32 = ((DIM + 1 + 3) / 4) * 4;
};

template <int S>
struct TcShared {
  uint64_t b_full, a_full[kAStages], a_empty[kAStages], tmem_full[2], tmem_empty[2], res_full[kResStages],
      res_empty[kResStages];
  uint32_t tmem_base;
  uint32_t pad;
  float res_best[kResStages][S][kTileQ], res_second[kResStages][S][kTileQ];
  int res_chunk[kResStages][S][kTileQ];
};

// minimum of eight scores merged into the running (best, runner-up) over CHUNK MINIMA; `chunk` remembers which group
// of 8 held the best.  9 alu-pipe operations per 8 scores.
__device__ __forceinline__ void chunkmin_update8(const float *a, int cid, float &best, float &runner, int &chunk) {
  const float m = fminf(fmin3(a[0], a[1], a[2]), fmin3(fmin3(a[3], a[4], a[5]), a[6], a[7]));
  runner = fminf(runner, fmaxf(best, m));
  chunk = m < best ? cid : chunk;
  best = fminf(best, m);
}

// What the FUSED variant needs on top: the FP32 rows (staged in shared memory next to the bf16 limbs) and the outputs
// of the finalise step, which its merger warps then do themselves - no per-query record leaves the SM.
struct TcFuse {
  const float *rows32;
  float margin_coef;
  const float *c_max_ptr;
  uint32_t *assign, *flag_list;
  unsigned int *flag_count;
  int keep_records;  // diagnostics (qb200_debug_filter_records): store the per-query records as the unfused path does
};

template <int DIM, int S, bool FUSE>
__global__ void __launch_bounds__(tc_threads(S), 1)
    assign_tc_kernel(const VecSource src, const unsigned char *__restrict__ b_staged, const int k_rows, const int k_base,
                     const int first_pass, float *__restrict__ state, const unsigned long long tiles, const TcFuse fuse) {
  using Cfg = TcCfg<DIM>;
  constexpr int KB = Cfg::KB;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // [ B chunk | (FUSE: FP32 rows) | A tile ring | TcShared ]
  const uint32_t b_bytes = (uint32_t)k_rows * Cfg::ROW_BYTES;
  const uint32_t r_bytes = FUSE ? (uint32_t)k_rows * Cfg::ROW32 * 4u : 0u;
  unsigned char *s_b = smem_raw;
  float4 *s_rows4 = reinterpret_cast<float4 *>(smem_raw + ((b_bytes + 1023u) & ~1023u));
  unsigned char *s_a = smem_raw + ((b_bytes + 1023u) & ~1023u) + ((r_bytes + 1023u) & ~1023u);
  TcShared<S> &sh = *reinterpret_cast<TcShared<S> *>(s_a + kAStages * Cfg::A_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles_n = k_rows / kTileN;  // N tiles per query tile (the host pads the codebook to a multiple)

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&sh.b_full, 1);
      for (int i = 0; i < kAStages; i++) {
        mbar_init(&sh.a_full[i], kTileQ);
        mbar_init(&sh.a_empty[i], 1);
      }
      for (int i = 0; i < 2; i++) {
        mbar_init(&sh.tmem_full[i], 1);
        mbar_init(&sh.tmem_empty[i], 128 * S);
      }
      for (int i = 0; i < kResStages; i++) {
        mbar_init(&sh.res_full[i], 128 * S);
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
      const uint32_t idesc = umma_idesc_bf16_f32(kTileN);
      const uint32_t a_addr = smem_u32(s_a), b_addr = smem_u32(s_b);
      const uint32_t b_tile_bytes = (uint32_t)kTileN * 32u;  // one (limb, K block) of one N tile
      unsigned int it = 0, tile_seq = 0;
      for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, tile_seq++) {
        const int ab = tile_seq % kAStages;
        mbar_wait_bounded<20000>(&sh.a_full[ab], (tile_seq / kAStages) & 1, 2);
        tc_fence_after();
        for (int jt = 0; jt < n_tiles_n; jt++, it++) {
          const int buf = it & 1;
          mbar_wait_bounded<20000>(&sh.tmem_empty[buf], ((it >> 1) & 1) ^ 1, 3);
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
  } else if (warp <= 4 * kProdGroups) {
    // =========================== producers: gather query bytes, write bf16 A rows ===========================
    // kProdGroups groups of 128 threads take alternate tiles: a tile costs one global-memory round trip per
    // thread, two groups keep two tiles in flight.
    const int group = (warp - 1) >> 2;
    const int r = ((warp - 1) & 3) * 32 + lane;  // query row inside a tile
    const uint32_t row_off = (uint32_t)(r >> 3) * 256u + (uint32_t)(r & 7) * 16u;
    // The query's bytes come as whole words from the dense copy of the training set (the host only selects this
    // kernel for lattice sources, which always have one).  The loads of the NEXT tile are issued before the
    // current one is converted and stored, so a tile does not cost a full global-memory round trip.
    constexpr int WORDS = (DIM + 3) / 4;
    auto load_words = [&](unsigned long long tile, uint32_t (&w)[WORDS]) {
      const unsigned long long v = tile * kTileQ + r;
#pragma unroll
      for (int i = 0; i < WORDS; i++) w[i] = 0u;
      if (tile < tiles && v < src.n_local) {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(src.dense + v * (unsigned long long)src.dense_stride);
        if constexpr (WORDS % 4 == 0) {
#pragma unroll
          for (int i = 0; i < WORDS / 4; i++) {
            const uint4 t = __ldg(reinterpret_cast<const uint4 *>(p) + i);
            w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < WORDS; i++) w[i] = __ldg(p + i);
        }
      }
    };
    // bf16 bit pattern of element e of the A row: the lattice value (exact in bf16), then the constant 1, then zeros
    auto a_elem = [&](const uint32_t (&w)[WORDS], int e) -> uint32_t {
      if (e < DIM) return __float_as_uint((float)(int)(signed char)(w[(e < DIM ? e : 0) >> 2] >> (8 * (e & 3)))) >> 16;
      return e == DIM ? 0x3F80u : 0u;
    };
    const unsigned long long step = (unsigned long long)kProdGroups * gridDim.x;
    unsigned int tile_seq = group;
    unsigned long long tile = blockIdx.x + (unsigned long long)group * gridDim.x;
    uint32_t cur[WORDS], nxt[WORDS];
    load_words(tile, cur);
    for (; tile < tiles; tile += step, tile_seq += kProdGroups) {
      load_words(tile + step, nxt);
      const int ab = tile_seq % kAStages;
      const unsigned int use = tile_seq / kAStages;
      if (use >= 1) mbar_wait_bounded<20000>(&sh.a_empty[ab], (use - 1) & 1, 4);
      unsigned char *a_tile = s_a + ab * Cfg::A_BYTES;
#pragma unroll
      for (int kb = 0; kb < KB; kb++) {
        uint32_t w[8];  // 16 bf16: small integers are exact in bf16 = upper half of their fp32 pattern
#pragma unroll
        for (int p = 0; p < 8; p++) w[p] = a_elem(cur, kb * 16 + 2 * p) | (a_elem(cur, kb * 16 + 2 * p + 1) << 16);
        uint4 *dst = reinterpret_cast<uint4 *>(a_tile + kb * 4096 + row_off);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);  // k 0..7 of this block
        dst[8] = make_uint4(w[4], w[5], w[6], w[7]);  // k 8..15: next core matrix (+128 B)
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(&sh.a_full[ab]);
#pragma unroll
      for (int i = 0; i < WORDS; i++) cur[i] = nxt[i];
    }
  } else if (warp <= 4 * kProdGroups + 4) {
    // =========================== mergers: combine the two column halves, store the per-query record ===========================
    // No dependent global loads here (they would serialise one tile per memory round trip); the margin
    // test and the index inside the winning chunk are done by tc_finalize_kernel after the last pass.
    const int r = (warp - 1 - 4 * kProdGroups) * 32 + lane;
    unsigned int tile_seq = 0;
    constexpr int WORDS = (DIM + 3) / 4, ROW32 = Cfg::ROW32, NF4 = ROW32 / 4;
    float c_max_norm = 0.f;
    // slot of float4 number q of row `row`: 16-float rows sit 64 bytes apart, so lanes reading the same q of unrelated
    // rows would hit two of the eight 16-byte bank groups; rotating by the row number spreads them over all eight
    auto slot = [](int row, int q) { return NF4 == 4 ? ((q + (row >> 1)) & 3) : (NF4 == 2 ? ((q + (row >> 2)) & 1) : q); };
    if (FUSE) {
      // the four merger warps copy the FP32 rows into shared memory themselves (they have nothing else to do until
      // the first tile is through the pipeline) and meet at their own named barrier
      const float4 *g = reinterpret_cast<const float4 *>(fuse.rows32);
      for (int i = r; i < k_rows * NF4; i += kTileQ) {
        const int row = i / NF4, q = i - row * NF4;
        s_rows4[row * NF4 + slot(row, q)] = __ldg(g + i);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      c_max_norm = *fuse.c_max_ptr;
    }
    for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, tile_seq++) {
      const unsigned long long v = tile * kTileQ + r;
      const int rb = tile_seq % kResStages;
      uint32_t xw[WORDS];
      if (FUSE) {  // the query's bytes again (L2 hit: the producers read them a moment ago), before the wait
#pragma unroll
        for (int i = 0; i < WORDS; i++) xw[i] = 0u;
        if (v < src.n_local) {
          const uint32_t *p = reinterpret_cast<const uint32_t *>(src.dense + v * (unsigned long long)src.dense_stride);
#pragma unroll
          for (int i = 0; i < WORDS; i++) xw[i] = __ldg(p + i);
        }
      }
      mbar_wait_bounded<20000>(&sh.res_full[rb], (tile_seq / kResStages) & 1, 5);
      float best = sh.res_best[rb][0][r], runner = sh.res_second[rb][0][r];
      int chunk = sh.res_chunk[rb][0][r];
#pragma unroll
      for (int i = 1; i < S; i++) {  // the column parts hold disjoint chunks: same rule as the running update
        const float b = sh.res_best[rb][i][r];
        runner = fmin3(runner, sh.res_second[rb][i][r], fmaxf(best, b));
        chunk = b < best ? sh.res_chunk[rb][i][r] : chunk;
        best = fminf(best, b);
      }
      mbar_arrive(&sh.res_empty[rb]);
      if (!FUSE || fuse.keep_records) {
        if (v < src.n_local) reinterpret_cast<float4 *>(state)[v] = make_float4(best, runner, __int_as_float(chunk), 0.f);
        if (!FUSE) continue;
      }
      // ---- fused finalise: FP32 rescoring of the winning chunk's eight rows from shared memory, margin test on both
      // gaps (tensor-core scores across chunks, FP32 scores inside the chunk), index and flag - see tc_finalize_kernel
      bool flag = false;
      if (v < src.n_local) {
        float x[DIM], xn = 0.f;
#pragma unroll
        for (int e = 0; e < DIM; e++) {
          x[e] = (float)(int)(signed char)(xw[e >> 2] >> (8 * (e & 3)));
          xn = fmaf(x[e], x[e], xn);
        }
        const float rr = sqrtf(xn) + c_max_norm;
        const float margin = fuse.margin_coef * rr * rr;
        float sb = FLT_MAX, s2 = FLT_MAX;
        int bidx = chunk * 8;
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const int row = chunk * 8 + c;
          float cr[ROW32];
#pragma unroll
          for (int q = 0; q < NF4; q++) {
            const float4 t = s_rows4[row * NF4 + slot(row, q)];
            cr[4 * q] = t.x; cr[4 * q + 1] = t.y; cr[4 * q + 2] = t.z; cr[4 * q + 3] = t.w;
          }
          float sc = cr[DIM];
#pragma unroll
          for (int e = 0; e < DIM; e++) sc = fmaf(x[e], cr[e], sc);
          s2 = fminf(s2, fmaxf(sb, sc));
          if (sc < sb) {
            sb = sc;
            bidx = row;
          }
        }
        flag = !((runner - best) > margin && (s2 - sb) > margin);
        fuse.assign[v] = (uint32_t)bidx | (flag ? kUndecided : 0u);
      }
      const unsigned int m = __ballot_sync(0xffffffffu, flag);
      if (m) {
        const int leader = __ffs(m) - 1;
        unsigned int basepos = 0;
        if (lane == leader) basepos = atomicAdd(fuse.flag_count, (unsigned int)__popc(m));
        basepos = __shfl_sync(0xffffffffu, basepos, leader);
        if (flag) fuse.flag_list[basepos + __popc(m & ((1u << lane) - 1u))] = (uint32_t)v;
      }
    }
  } else {
    // =========================== epilogue: TMEM -> registers -> running top-2 ===========================
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, 32*quarter + 32) are the ones this warp may read
    const int half = (warp - 5 - 4 * kProdGroups) >> 2;  // which half of the N columns
    const int r = quarter * 32 + lane;     // query row = TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    unsigned int it = 0, tile_seq = 0;
    for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, tile_seq++) {
      float best = FLT_MAX, second = FLT_MAX;
      int chunk = k_base >> 3;
      if (!first_pass && half == 0) {
        const unsigned long long v = tile * kTileQ + r;
        if (v < src.n_local) {
          const float4 rec = reinterpret_cast<const float4 *>(state)[v];
          best = rec.x;
          second = rec.y;
          chunk = __float_as_int(rec.z);
        }
      }
      // Per N tile this warp scans its kTileN / S columns as four pieces of CW, software-pipelined over two register
      // buffers: the tcgen05.ld of the next piece (the next N tile's first piece included) is in flight while
      // the current one is reduced.  Everything but the buffer address and the chunk base is a constant.
      constexpr int part_cols = kTileN / S, CW = part_cols / 4;
      const int col0 = half * part_cols;
      auto loadp = [&](uint32_t taddr, float *v) {
#pragma unroll
        for (int q = 0; q < CW / 8; q++) tmem_ld8(taddr + q * 8, v + q * 8);
      };
      auto landed = [&](float *v) {  // after this the piece is in registers
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < CW / 8; q++) tmem_ld_pin8(v + q * 8);
      };
      auto reduce = [&](int cid, const float *v) {
#pragma unroll
        for (int q = 0; q < CW / 8; q++) chunkmin_update8(v + q * 8, cid + q, best, second, chunk);
      };
      auto acquire = [&](unsigned int itg) -> uint32_t {  // wait for N tile `itg`, return its first column address
        const int buf = itg & 1;
        mbar_wait_bounded(&sh.tmem_full[buf], (itg >> 1) & 1, 6);
        tc_fence_after();
        return lane_addr + (uint32_t)(buf * 256 + col0);
      };
      float va[CW], vb[CW];
      uint32_t taddr = acquire(it);
      loadp(taddr, va);
      int cid = (k_base + col0) >> 3;
      for (int jt = 0; jt < n_tiles_n; jt++, cid += kTileN / 8) {
        landed(va);
        loadp(taddr + CW, vb);
        reduce(cid, va);
        landed(vb);
        loadp(taddr + 2 * CW, va);
        reduce(cid + CW / 8, vb);
        landed(va);
        loadp(taddr + 3 * CW, vb);
        reduce(cid + 2 * (CW / 8), va);
        landed(vb);  // the whole N tile is in registers: hand the buffer back to the MMA
        tc_fence_before();
        mbar_arrive(&sh.tmem_empty[(it + jt) & 1]);
        if (jt + 1 < n_tiles_n) {
          taddr = acquire(it + jt + 1);
          loadp(taddr, va);
        }
        reduce(cid + 3 * (CW / 8), vb);
      }
      it += n_tiles_n;
      const int rb = tile_seq % kResStages;
      if (tile_seq >= kResStages) mbar_wait_bounded(&sh.res_empty[rb], ((tile_seq / kResStages) - 1) & 1, 7);
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

// After the last pass: per query, read its (best, second, chunk) record, apply the margin test and - for
// decided queries - find the member of the winning chunk of 8 with FP32 scores from the FP32 rows.
// Undecided queries go to the flag list for the exact resolver.
//
// The chunk's 8 rows are contiguous (8 * ROW32 floats).  Gathering them one query per thread touches 32
// different cache lines per warp-wide load and the L1 tag stage (one line per clock) becomes the bound,
// so for 16-float rows (dim 12..15) FOUR LANES SHARE ONE QUERY, each reading 32 bytes (256-bit load) per step:
// a warp-wide load then touches 8 whole lines for 8 queries, and the per-query bookkeeping is replicated 4
// times instead of 32 (measured balance between L1 tag rate and instruction count).
template <int DIM>
__global__ void __launch_bounds__(256)
    tc_finalize_kernel(const VecSource src, const float *__restrict__ rows32, const float *__restrict__ state,
                       const float margin_coef, const float *__restrict__ c_max_ptr, uint32_t *__restrict__ assign,
                       uint32_t *__restrict__ flag_list, unsigned int *__restrict__ flag_count) {
  constexpr int ROW32 = TcCfg<DIM>::ROW32;
  const float c_max_norm = *c_max_ptr;
  const int lane = threadIdx.x & 31;
  if constexpr (ROW32 == 16) {
    // FOUR lanes share one query (src.dense is set by the host).  The chunk's 8 rows are 512 contiguous bytes; at
    // step i the group reads 128 of them with one 256-bit load per lane: lane j gets the column half h = j & 1
    // (floats 8h..8h+7) of row 2i + (j >> 1).  One shuffle completes a row's score, one more merges the two row
    // parities.  The query's bytes come as words from the dense copy: lane j loads word j and the two words a
    // lane needs travel by shuffle.
    constexpr int WORDS = (DIM + 3) / 4;
    const int j = lane & 3, h = j & 1;
    const int gbase = lane & 28;
    const unsigned int gmask = 0xFu << gbase;  // the four lanes of this group: a dead group skips the shuffles
    const unsigned int gthread = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n_groups = gridDim.x * blockDim.x / 4;
    const unsigned int n_local = (unsigned int)src.n_local;
    const unsigned int n_round = (n_local + 7u) & ~7u;  // 8 queries per warp per round
    const unsigned int *xwords = reinterpret_cast<const unsigned int *>(src.dense);
    for (unsigned int v = gthread >> 2; v < n_round; v += n_groups) {
      bool flag = false;
      if (v < n_local) {
        const float4 rec = __ldcs(reinterpret_cast<const float4 *>(state) + v);  // streamed: keep the rows in L1
        const int chunk = __float_as_int(rec.z);
        const unsigned int mine = j < WORDS ? __ldg(xwords + (size_t)v * WORDS + j) : 0u;
        const unsigned int w0 = __shfl_sync(gmask, mine, gbase + 2 * h);
        const unsigned int w1 = __shfl_sync(gmask, mine, gbase + 2 * h + 1);
        float xe[8], part = 0.f;
#pragma unroll
        for (int t = 0; t < 8; t++) {
          const unsigned int word = t < 4 ? w0 : w1;
          const float byte = (float)(int)(signed char)(word >> (8 * (t & 3)));
          // element index e = 8h + t: coordinates below DIM, then the constant 1 that picks up |C|^2, then zeros
          const float lo = t < DIM ? byte : (t == DIM ? 1.f : 0.f);              // h == 0
          const float hi = (8 + t) < DIM ? byte : ((8 + t) == DIM ? 1.f : 0.f);  // h == 1
          xe[t] = h ? hi : lo;
          const bool coord = h ? (8 + t) < DIM : t < DIM;
          part = coord ? fmaf(xe[t], xe[t], part) : part;
        }
        part += __shfl_xor_sync(gmask, part, 1);
        const float rr = sqrtf(part) + c_max_norm;
        const float margin = margin_coef * rr * rr;
        const float *rows = rows32 + (size_t)chunk * 8 * ROW32 + j * 8;
        float sb = FLT_MAX, s2 = FLT_MAX;  // best and second-best FP32 score among this lane's four rows
        int rbest = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float c[8];
          asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
              : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3]), "=f"(c[4]), "=f"(c[5]), "=f"(c[6]), "=f"(c[7])
              : "l"(rows + i * 32));
          float p = xe[0] * c[0];
#pragma unroll
          for (int t = 1; t < 8; t++) p = fmaf(xe[t], c[t], p);
          p += __shfl_xor_sync(gmask, p, 1);
          s2 = fminf(s2, fmaxf(sb, p));
          if (p < sb) {
            sb = p;
            rbest = 2 * i + (j >> 1);
          }
        }
        const float so = __shfl_xor_sync(gmask, sb, 2);
        const float s2o = __shfl_xor_sync(gmask, s2, 2);
        const int ro = __shfl_xor_sync(gmask, rbest, 2);
        if (so < sb || (so == sb && ro < rbest)) rbest = ro;
        // decided: the winning chunk beats every other chunk (tensor-core scores) AND its best member beats the
        // other seven (FP32 scores), both by more than the tensor-core margin
        const float in_second = fmin3(fmaxf(sb, so), s2, s2o), in_best = fminf(sb, so);
        flag = !((rec.y - rec.x) > margin && (in_second - in_best) > margin);
        if (j == 0) assign[v] = (uint32_t)(chunk * 8 + rbest) | (flag ? kUndecided : 0u);
        flag = flag && j == 0;
      }
      const unsigned int m = __ballot_sync(0xffffffffu, flag);
      if (m) {
        const int leader = __ffs(m) - 1;
        unsigned int basepos = 0;
        if (lane == leader) basepos = atomicAdd(flag_count, (unsigned int)__popc(m));
        basepos = __shfl_sync(0xffffffffu, basepos, leader);
        if (flag) flag_list[basepos + __popc(m & ((1u << lane) - 1u))] = (uint32_t)v;
      }
    }
  } else {
    const unsigned long long n_round = (src.n_local + 31ull) & ~31ull;  // whole warps stay in the loop for the ballot
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_round;
         v += (unsigned long long)gridDim.x * blockDim.x) {
      const bool live = v < src.n_local;
      bool flag = false;
      if (live) {
        const float4 rec = __ldcs(reinterpret_cast<const float4 *>(state) + v);  // streamed: keep the rows in L1
        const float best = rec.x, second = rec.y;
        const int chunk = __float_as_int(rec.z);
        float x[DIM], xn = 0.f;
        gather_lattice<DIM>(src, v, x);
#pragma unroll
        for (int e = 0; e < DIM; e++) xn = fmaf(x[e], x[e], xn);
        const float rr = sqrtf(xn) + c_max_norm;
        const float margin = margin_coef * rr * rr;
        int bidx = chunk * 8;
        float sb = FLT_MAX, s2 = FLT_MAX;  // best / second-best FP32 score inside the winning chunk
        {
          const float4 *rows = reinterpret_cast<const float4 *>(rows32 + (size_t)chunk * 8 * ROW32);
#pragma unroll
          for (int c = 0; c < 8; c++) {
            float cr[ROW32];
#pragma unroll
            for (int q4 = 0; q4 < ROW32 / 4; q4++) {
              const float4 t = __ldg(rows + c * (ROW32 / 4) + q4);
              cr[4 * q4] = t.x; cr[4 * q4 + 1] = t.y; cr[4 * q4 + 2] = t.z; cr[4 * q4 + 3] = t.w;
            }
            float s = cr[DIM];
#pragma unroll
            for (int e = 0; e < DIM; e++) s = fmaf(x[e], cr[e], s);
            s2 = fminf(s2, fmaxf(sb, s));
            if (s < sb) {
              sb = s;
              bidx = chunk * 8 + c;
            }
          }
        }
        flag = !((second - best) > margin && (s2 - sb) > margin);
        assign[v] = (uint32_t)bidx | (flag ? kUndecided : 0u);
      }
      const unsigned int m = __ballot_sync(0xffffffffu, flag);
      if (m) {
        const int leader = __ffs(m) - 1;
        unsigned int basepos = 0;
        if (lane == leader) basepos = atomicAdd(flag_count, (unsigned int)__popc(m));
        basepos = __shfl_sync(0xffffffffu, basepos, leader);
        if (flag) flag_list[basepos + __popc(m & ((1u << lane) - 1u))] = (uint32_t)v;
      }
    }
  }
}

// Same job with the FP32 rows staged in SHARED memory (used whenever the padded table fits): one thread per
// query, the eight candidate rows are read with LDS.128.  Row r keeps its float4 number q at slot
// (q + (r >> 1)) & 3 (16-float rows) so that 32 lanes reading the same q of 32 unrelated rows spread over all
// eight 16-byte bank groups; odd row lengths spread by themselves.
template <int DIM>
__global__ void __launch_bounds__(1024, 1)
    tc_finalize_smem_kernel(const VecSource src, const float *__restrict__ rows32, const int k_rows,
                            const float *__restrict__ state, const float margin_coef,
                            const float *__restrict__ c_max_ptr, uint32_t *__restrict__ assign,
                            uint32_t *__restrict__ flag_list, unsigned int *__restrict__ flag_count) {
  constexpr int ROW32 = TcCfg<DIM>::ROW32, NF4 = ROW32 / 4;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float4 *s_rows = reinterpret_cast<float4 *>(smem_raw);
  auto slot = [](int r, int q) { return NF4 == 4 ? ((q + (r >> 1)) & 3) : (NF4 == 2 ? ((q + (r >> 2)) & 1) : q); };
  for (int i = threadIdx.x; i < k_rows * NF4; i += blockDim.x) {
    const int r = i / NF4, q = i - r * NF4;
    s_rows[r * NF4 + slot(r, q)] = __ldg(reinterpret_cast<const float4 *>(rows32) + i);
  }
  __syncthreads();
  const float c_max_norm = *c_max_ptr;
  const int lane = threadIdx.x & 31;
  const unsigned long long n_round = (src.n_local + 31ull) & ~31ull;  // whole warps stay in the loop for the ballot
  for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_round;
       v += (unsigned long long)gridDim.x * blockDim.x) {
    bool flag = false;
    if (v < src.n_local) {
      const float4 rec = __ldcs(reinterpret_cast<const float4 *>(state) + v);  // streamed: keep the rows in L1
      const float best = rec.x, second = rec.y;
      const int chunk = __float_as_int(rec.z);
      float x[DIM], xn = 0.f;
      gather_lattice<DIM>(src, v, x);
#pragma unroll
      for (int e = 0; e < DIM; e++) xn = fmaf(x[e], x[e], xn);
      const float rr = sqrtf(xn) + c_max_norm;
      const float margin = margin_coef * rr * rr;
      int bidx = chunk * 8;
      float sb = FLT_MAX, s2 = FLT_MAX;  // best / second-best FP32 score inside the winning chunk
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const int r = chunk * 8 + c;
        float cr[ROW32];
#pragma unroll
        for (int q = 0; q < NF4; q++) {
          const float4 t = s_rows[r * NF4 + slot(r, q)];
          cr[4 * q] = t.x; cr[4 * q + 1] = t.y; cr[4 * q + 2] = t.z; cr[4 * q + 3] = t.w;
        }
        float sc = cr[DIM];
#pragma unroll
        for (int e = 0; e < DIM; e++) sc = fmaf(x[e], cr[e], sc);
        s2 = fminf(s2, fmaxf(sb, sc));
        if (sc < sb) {
          sb = sc;
          bidx = r;
        }
      }
      flag = !((second - best) > margin && (s2 - sb) > margin);
      assign[v] = (uint32_t)bidx | (flag ? kUndecided : 0u);
    }
    const unsigned int m = __ballot_sync(0xffffffffu, flag);
    if (m) {
      const int leader = __ffs(m) - 1;
      unsigned int basepos = 0;
      if (lane == leader) basepos = atomicAdd(flag_count, (unsigned int)__popc(m));
      basepos = __shfl_sync(0xffffffffu, basepos, leader);
      if (flag) flag_list[basepos + __popc(m & ((1u << lane) - 1u))] = (uint32_t)v;
    }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
int tc_kblocks(int dim) { return (dim + 1 + 15) / 16; }
size_t tc_row_bytes(int dim) { return (size_t)tc_kblocks(dim) * 3 * 32; }
bool tc_supported(int dim, int K) {
  switch (dim) {
    case 3: case 6: case 9: case 12: case 24: case 27: case 48: return K >= 64;
    default: return false;
  }
}
int tc_padded_rows(int K) { return ((K + kTileN - 1) / kTileN) * kTileN; }
int tc_n_tile(int K) { (void)K; return kTileN; }
// Rows per launch: what fits in shared memory next to the two A tiles, a multiple of the N tile.
int tc_chunk_rows(int dim, int K) {
  const int kp = tc_padded_rows(K), nt = tc_n_tile(K);
  // 227 KB per CTA: B chunk + A ring + barriers/result ring (sizeof(TcShared<2>), 12 KB) + alignment slack
  const size_t budget = 227 * 1024 - sizeof(TcShared<2>) - 2048 - kAStages * (size_t)tc_kblocks(dim) * 4096;
  int rows = (int)(budget / tc_row_bytes(dim));
  rows = (rows / nt) * nt;
  return rows < kp ? rows : kp;
}

template <int DIM>
static cudaError_t launch_tc_t(const AssignTcLaunch &a) {
  using Cfg = TcCfg<DIM>;
  const int kp = tc_padded_rows(a.K), chunk = tc_chunk_rows(DIM, a.K);
  const unsigned long long tiles = (a.src.n_local + kTileQ - 1) / kTileQ;
  if (tiles == 0) return cudaSuccess;
  // Two epilogue warps per TMEM lane quarter (each scans 128 columns in pieces of 32).  Four (64 columns in pieces of
  // 16) were measured as well: 0.675 against 0.686 ms at K = 1024 on config 2 - the epilogue is bound by the alu pipe,
  // not by tcgen05.ld latency - and their larger result ring costs a third pass at K = 4096, so two it is.
  constexpr int kSplit = 2;
  const size_t b_smem = ((size_t)chunk * Cfg::ROW_BYTES + 1023) & ~(size_t)1023;
  const size_t tail_smem = kAStages * Cfg::A_BYTES + sizeof(TcShared<kSplit>) + 1024;
  const size_t rows_smem = ((size_t)kp * Cfg::ROW32 * 4 + 1023) & ~(size_t)1023;
  const unsigned int grid = (unsigned int)(tiles < (unsigned long long)a.sm_count ? tiles : a.sm_count);
  const int threads = tc_threads(kSplit);
  // FUSED variant (opt-in, QB200_TC_FUSE=1): when the whole codebook is one pass and its FP32 rows fit next to the bf16
  // limbs, the merger warps do the finalise step themselves (margin test, member of the winning chunk, index, flag)
  // and no 16-byte per-query record is written and read back.  MEASURED SLOWER than the separate finalise kernel
  // (config 2: 0.52 / 0.59 / 0.71 / 0.86 ms at K = 128 / 256 / 512 / 1024 against 0.36 / 0.36 / 0.47 / 0.69 ms for
  // kernel + tc_finalize_kernel): the four merger warps handle one tile at a time - 32 LDS.128 and ~150 dependent
  // FP32 operations per query behind the epilogue's hand-off - and become the pipeline's slowest stage, while the
  // separate kernel runs the same work at full occupancy in 0.15 ms.  Kept for the record, off by default.
  static const bool fuse_enabled = [] {
    const char *e = std::getenv("QB200_TC_FUSE");
    return e && e[0] == '1';
  }();
  if (fuse_enabled && chunk == kp && b_smem + rows_smem + tail_smem <= 227 * 1024) {
    const size_t smem = b_smem + rows_smem + tail_smem;
    auto kernel = assign_tc_kernel<DIM, kSplit, true>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const char *dbg = std::getenv("QB200_DEBUG_RECORDS");
    const TcFuse f{a.rows32, a.margin_coef, a.c_max_ptr, a.assign, a.flag_list, a.flag_count, dbg && dbg[0] == '1' ? 1 : 0};
    kernel<<<grid, threads, smem, a.stream>>>(a.src, a.b_staged, kp, 0, 1, a.state, tiles, f);
    count_launch();
    return cudaGetLastError();
  }
  const size_t smem_max = b_smem + tail_smem;
  auto kernel = assign_tc_kernel<DIM, kSplit, false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
  if (e != cudaSuccess) return e;
  const TcFuse no_fuse{};
  for (int k0 = 0; k0 < kp; k0 += chunk) {
    const int rows = kp - k0 < chunk ? kp - k0 : chunk;
    kernel<<<grid, threads, smem_max, a.stream>>>(a.src, a.b_staged + (size_t)k0 * Cfg::ROW_BYTES, rows,
                                                    k0, k0 == 0, a.state, tiles, no_fuse);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const size_t table = (size_t)kp * Cfg::ROW32 * 4;
  if (table <= 200 * 1024 && Cfg::ROW32 != 16) {  // FP32 rows fit in shared memory: one thread per query, LDS gathers
    e = cudaFuncSetAttribute(tc_finalize_smem_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    unsigned long long fb = (a.src.n_local + 1023) / 1024;
    if (fb > (unsigned long long)a.sm_count) fb = (unsigned long long)a.sm_count;
    tc_finalize_smem_kernel<DIM><<<(unsigned int)fb, 1024, table, a.stream>>>(a.src, a.rows32, kp, a.state, a.margin_coef,
                                                                             a.c_max_ptr, a.assign, a.flag_list, a.flag_count);
    count_launch();
    return cudaGetLastError();
  }
  const unsigned long long per_block = Cfg::ROW32 == 16 ? 64 : 256;  // queries per 256-thread block per round
  unsigned long long blocks = (a.src.n_local + per_block - 1) / per_block;
  const unsigned long long cap = (unsigned long long)a.sm_count * 8;
  if (blocks > cap) blocks = cap;
  tc_finalize_kernel<DIM><<<(unsigned int)blocks, 256, 0, a.stream>>>(a.src, a.rows32, a.state, a.margin_coef, a.c_max_ptr,
                                                                      a.assign, a.flag_list, a.flag_count);
  count_launch();
  return cudaGetLastError();
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
