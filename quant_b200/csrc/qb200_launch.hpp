// Host-callable launchers of the kernels in qb200_kernels.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "qb200_device.cuh"

namespace qb {

struct AssignLaunch {
  VecSource src;
  const float *cb_rows;      // K rows of assign_row_floats(dim) floats: [-2*C_k, |C_k|^2, pad]
  int K;
  float margin_coef;         // flag when second - best <= margin_coef * (|X| + c_max_norm)^2
  const float *c_max_ptr;    // device: max_k |C_k| (lattice units), rounded up (written by stage_codebook_kernel)
  uint32_t *assign;          // n_local
  uint32_t *flag_list;       // n_local (capacity)
  unsigned int *flag_count;  // device counter, zeroed by the caller
  int sm_count;
  cudaStream_t stream;
  // optional fusion of the per-cell statistics of the DECIDED queries (flagged ones are left to the resolver):
  unsigned long long *stats;  // K*(dim+2) words, zeroed by the caller; null = no fusion
  int k_real;                 // codevectors without padding rows
  bool *fused_out;            // set to whether the kernel really accumulated (the table must fit in shared memory)
};

int assign_row_floats(int dim);
// Dense byte copy of the training set (stride = dim rounded up to 4); src.dense must still be null.
cudaError_t launch_pack_vectors(const VecSource &src, uint8_t *dense, int stride, int sm_count, cudaStream_t stream);
// FP64 codebook (device) -> FP32 rows (+ bf16 limb tiles when tc_out != null) + max codevector norm; c_max must be zeroed.
cudaError_t launch_stage_codebook(const double *cb, int K, int k_rows32, int k_rows_tc, int dim, int scaled,
                                  float *rows32, unsigned char *tc_out, float *c_max, double *cb_t,
                                  cudaStream_t stream);
cudaError_t launch_assign(const AssignLaunch &a);
// FP32 middle tier of the tensor-core levels: exact top-2 re-rank of the queries in list_in against all K staged rows;
// decided ones get their final index, the rest are appended to list_out (count_out must be zeroed).
cudaError_t launch_refilter(const VecSource &src, const float *cb_rows, int K, float margin_coef, const float *c_max_ptr,
                            uint32_t *assign, const uint32_t *list_in, const unsigned int *count_in, uint32_t *list_out,
                            unsigned int *count_out, int sm_count, cudaStream_t stream);
// Exact re-solve of the flagged queries: brute-force FP64 phase, then the reference's KD walk for
// the (near-)exact ties it leaves in tie_list (capacity: n_local).  Counters are device words.
cudaError_t launch_resolve(const VecSource &src, int scaled, const double *cb, const double *cbt, int K,
                           const KdDevice &tree,
                           const uint32_t *flag_list, const unsigned int *flag_count, uint32_t *assign,
                           uint32_t *tie_list, unsigned int *tie_count, unsigned int *changed,
                           unsigned long long *stats /* null unless the flagged queries' statistics are added here */,
                           uint32_t *result /* null: exact indices go straight to assign; else to result[v], see below */,
                           unsigned int *sensitive /* device counter (may be null): decisions that hinged on (near-)ties */,
                           const unsigned char *cv_exact /* per codevector (may be null): reproduced bit for bit by the integer path */,
                           int tree_robust /* the tree keeps its shape under last-bit changes of the codebook (KdHostTree::min_margin) */,
                           int sm_count, cudaStream_t stream);
// assign[v] = result[v] for the flagged queries: used when a statistics pass read `assign` while the resolver ran.
cudaError_t launch_commit_resolved(const uint32_t *flag_list, const unsigned int *flag_count, const uint32_t *result,
                                   uint32_t *assign, int sm_count, cudaStream_t stream);
// stats must be zeroed by the caller; assign may be null only for K == 1.
cudaError_t launch_accumulate(const VecSource &src, const uint32_t *assign, int K, unsigned long long *stats,
                              int sm_count, cudaStream_t stream);

// Tensor-core (tcgen05) variant of the filter, qb200_assign_tc.cu.  Same contract as launch_assign;
// b_staged holds the codebook as three bf16 limbs in the UMMA shared-memory layout (see
// tc_stage_codebook in qb200_api.cu), rows32 the FP32 rows padded to tc_padded_rows(K) rows.
struct AssignTcLaunch {
  VecSource src;
  const unsigned char *b_staged;
  const float *rows32;
  int K;              // real codevectors
  float margin_coef;
  const float *c_max_ptr;  // device
  float *state;       // n_local float4: per-query (best, second, chunk, -) records between the passes and the finalise kernel
  uint32_t *assign, *flag_list;
  unsigned int *flag_count;
  int sm_count;
  cudaStream_t stream;
};
bool tc_supported(int dim, int K);
int tc_kblocks(int dim);
size_t tc_row_bytes(int dim);
int tc_padded_rows(int K);
int tc_n_tile(int K);
int tc_chunk_rows(int dim, int K);
cudaError_t launch_assign_tc(const AssignTcLaunch &a);
void count_launch();

struct DecodeGeom {
  int xSize, ySize, w, h;
  unsigned int wB, hB;
  unsigned long long n_pixels;  // per image
  int n_images;
};
cudaError_t launch_decode(const DecodeGeom &g, const uint8_t *orig, const uint32_t *assign, const uint8_t *cb_bytes,
                          uint8_t *out, unsigned long long *sq_err, int sm_count, cudaStream_t stream);
// Device-side fix + distortions + split of one level.  summary: 32 bytes {double dist_pre, dist_post; u32 dead_cells,
// pad; u64 vectors counted}.  cb_pre / cb_next may be null (K = 1 has no previous codebook; the last level no next).
// exact_state (null = centroids from the integer sums): K*dim pairs {sum, c} left by launch_kahan_sums; the
// centroid is then sum / n, the reference's own operation.
cudaError_t launch_finalize_split(const unsigned long long *stats, const double *cb_pre, const double *exact_state, int K,
                                  int dim, int scaled, double n_total, double f_up, double f_dn, double *cb_post,
                                  double *cb_next, void *summary, unsigned char *exact_next /* 2K flags for cb_next, or null */,
                                  const unsigned char *small_flag, const double *small_sums /* small cells, or null */,
                                  cudaStream_t stream);
// Bit-exact centroid sums (qb200_exact.cu): stable sort of the members by cell, then the reference's compensated
// summation in ascending vector order, one warp per cell.
// Stable LSD radix sort of (cell, vector index) by cell (qb200_sort.cu): keys_out ascending, order_out = original positions.
size_t stable_sort_temp_bytes(size_t n);
cudaError_t launch_stable_sort_by_cell(const uint32_t *keys, uint32_t *keys_out, uint32_t *order_out, size_t n, int key_bits, void *tmp,
                                       size_t tmp_bytes, cudaStream_t stream);
cudaError_t launch_kahan_sums(const VecSource &src, const uint32_t *keys_sorted, const uint32_t *order, int K, int scaled,
                              double *state, unsigned long long *counts, cudaStream_t stream);
// The same sums evaluated in parallel (SCALED lattice sources only; method in qb200_exact_fast.cuh): bit-identical
// to launch_kahan_sums at a small fraction of its latency-bound cost.  workspace: exact_fast_workspace_bytes(n, K, dim).
size_t exact_fast_workspace_bytes(size_t n, int K, int dim);
cudaError_t launch_kahan_sums_fast(const VecSource &src, const uint32_t *keys_sorted, const uint32_t *order, int K, double *state,
                                   unsigned long long *counts, void *workspace, size_t workspace_bytes, int sm_count,
                                   cudaStream_t stream);
// Small cells (<= kSmallCellMax members): the reference's compensated sums next to the integer statistics, so that the
// auto centroid mode can treat their centroids as reproduced bit for bit (qb200_exact.cu, "small cells").
constexpr int kSmallCellMax = 8;
size_t small_cells_workspace_bytes(int K, int dim);
size_t small_cells_table_words(int K, int dim);
unsigned long long *small_cells_table(void *ws, int K, int dim);
const unsigned char *small_cells_flags(void *ws, int K, int dim);
const double *small_cells_sums(void *ws, int K, int dim);
cudaError_t launch_small_cells_reset(void *ws, cudaStream_t stream);
cudaError_t launch_small_cells_count(const unsigned long long *stats_local, int K, int dim, int rank, unsigned long long *packed,
                                     cudaStream_t stream);
cudaError_t launch_small_cells_collect(const VecSource &src, const uint32_t *assign, const unsigned long long *stats,
                                       const unsigned long long *packed, int K, int rank, int level, void *ws, int sm_count,
                                       cudaStream_t stream);
cudaError_t launch_small_cells_sums(const unsigned long long *stats, int K, int dim, void *ws, cudaStream_t stream);
// General FP64 training vectors (qb200_generic.cu).
// CIE1931 colour space (src/ColorSpace.cpp:31-39): the image's block vectors as doubles, n_local x dim.
cudaError_t launch_cie_vectors(const VecSource &src, double *out, int sm_count, cudaStream_t stream);
// updateDistortion (src/Quantizer.cpp:9-22) for FP64 vectors: partials[b] = sum over block b's vectors of
// |x - cb[assign]|^2; the caller adds the `blocks` partials in order and divides by N*dim.  Returns blocks.
int distortion_blocks(int sm_count);
cudaError_t launch_distortion_f64(const VecSource &src, const uint32_t *assign, const double *cb, double *partials,
                                  int sm_count, cudaStream_t stream);
// slots[0..world) = 0 except slots[rank] = partials[0] + partials[1] + ... (in that order): what a rank contributes
// to the sum all-reduce so that every rank ends up with every rank's partial sum, bit patterns intact.
cudaError_t launch_sum_partials(const double *partials, int n_partials, double *slots, int rank, int world,
                                cudaStream_t stream);
// Empty-cell repair (QB200_MODE_FULL_REPAIR): smallest (hash, global index) key per donor cell; member bytes.
cudaError_t launch_pick_members(const VecSource &src, const uint32_t *assign, const int *slot_of_cell,
                                unsigned long long seed, unsigned long long *keys, int sm_count, cudaStream_t stream);
cudaError_t launch_fetch_members(const VecSource &src, const long long *local_idx, int count, unsigned long long *out,
                                 cudaStream_t stream);
// Indices as a bit stream (qb200_get_assign_packed): out_words 32-bit words, zero padded.
cudaError_t launch_pack_indices(const uint32_t *assign, unsigned long long n, int bits, uint32_t *out,
                                unsigned long long out_words, int sm_count, cudaStream_t stream);
cudaError_t launch_ffma_probe(float *out, int blocks, int iters, float m, float c, cudaStream_t stream);

// Peer-memory sum all-reduce (qb200_comm.cu).  Every rank owns an exchange block: an array of epoch flags (one slot
// per rank) and a two-halved word buffer; `flags[p]` / `bufs[p]` are rank p's, mapped into this rank's address space.
constexpr int kCommMaxWorld = 16;
struct CommPeers {
  unsigned long long *flags[kCommMaxWorld];
  unsigned long long *bufs[kCommMaxWorld];
};
cudaError_t launch_comm_allreduce(unsigned long long *dev_words, size_t count, const CommPeers &peers, size_t cap_words, int rank,
                                  int world, unsigned long long epoch, unsigned int *ticket, int sm_count, cudaStream_t stream);

constexpr int kResolveDepthCap = 512;

int launch_count();
void reset_launch_count();

}  // namespace qb
