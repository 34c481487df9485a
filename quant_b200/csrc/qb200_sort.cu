// Stable sort of (cell, vector index) by cell - the member lists the compensated sums walk in ascending vector
// order (Solution::fixCodeVectors buckets the indices per cell in that order, /root/reference/src/Quantizer.cpp:75-78).
//
// Least-significant-digit radix sort with 8-bit digits: ceil(log2 K / 8) passes (one for K <= 256, two up to 65536).
// The unit of work is a WARP TILE of kTile consecutive items, processed in rows of 32: lanes holding the same digit
// find each other with __match_any_sync, so the rank of an item inside its row is a popcount and one lane per
// distinct digit updates the warp's private 256-bin table in shared memory - no atomics, no conflicts, and the
// order of equal digits is the order of the rows: stable by construction.
//   sort_hist_kernel     per tile: digit histogram -> hist[digit][tile]
//   sort_scan*_kernel    exclusive prefix over hist in (digit, tile) order (three small kernels)
//   sort_scatter_kernel  per tile: same row-by-row ranking, items written to out[prefix[digit][tile] + running rank]
// In the first pass the values are the item indices themselves and are not read.
#include "qb200_launch.hpp"

namespace qb {

namespace {

constexpr int kTile = 2048;       // items per warp tile
constexpr int kWarpsPerBlock = 8;
constexpr int kScanBlock = 1024;  // elements per block of the prefix kernels (one per thread)

__global__ void __launch_bounds__(32 * kWarpsPerBlock)
    sort_hist_kernel(const uint32_t *__restrict__ keys, const unsigned int n, const int shift, const unsigned int n_tiles,
                     uint32_t *__restrict__ hist) {
  __shared__ unsigned int s_hist[kWarpsPerBlock][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned int tile = blockIdx.x * kWarpsPerBlock + warp;
  for (int i = lane; i < 256; i += 32) s_hist[warp][i] = 0;
  __syncwarp();
  if (tile < n_tiles) {
    const unsigned int base = tile * kTile, end = min(n, base + kTile);
    for (unsigned int i = base + lane; i - lane < end; i += 32) {  // whole rows: every lane stays in the loop
      const bool live = i < end;
      const unsigned int d = live ? (__ldg(keys + i) >> shift) & 255u : 256u + lane;  // dead lanes: unique pseudo digits
      const unsigned int peers = __match_any_sync(0xffffffffu, d);
      if (live && lane == __ffs(peers) - 1) s_hist[warp][d] += __popc(peers);
      __syncwarp();
    }
    for (int i = lane; i < 256; i += 32) hist[(size_t)i * n_tiles + tile] = s_hist[warp][i];
  }
}

// exclusive prefix of `data` (len elements) in three steps: per-block scan + block totals, scan of the totals, add
__global__ void __launch_bounds__(kScanBlock) sort_scan1_kernel(uint32_t *__restrict__ data, const unsigned int len,
                                                                uint32_t *__restrict__ block_sums) {
  __shared__ unsigned int s[kScanBlock];
  const unsigned int i = blockIdx.x * kScanBlock + threadIdx.x;
  const unsigned int v = i < len ? data[i] : 0u;
  s[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < kScanBlock; o <<= 1) {
    const unsigned int add = (int)threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
    __syncthreads();
    s[threadIdx.x] += add;
    __syncthreads();
  }
  if (i < len) data[i] = s[threadIdx.x] - v;
  if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = s[threadIdx.x];
}
__global__ void __launch_bounds__(kScanBlock) sort_scan2_kernel(uint32_t *__restrict__ block_sums, const unsigned int n_blocks) {
  __shared__ unsigned int s[kScanBlock];
  __shared__ unsigned int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (unsigned int base = 0; base < n_blocks; base += kScanBlock) {
    const unsigned int i = base + threadIdx.x;
    const unsigned int v = i < n_blocks ? block_sums[i] : 0u;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < kScanBlock; o <<= 1) {
      const unsigned int add = (int)threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
      __syncthreads();
      s[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < n_blocks) block_sums[i] = carry + s[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == kScanBlock - 1) carry += s[threadIdx.x];
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kScanBlock) sort_scan3_kernel(uint32_t *__restrict__ data, const unsigned int len,
                                                                const uint32_t *__restrict__ block_sums) {
  const unsigned int i = blockIdx.x * kScanBlock + threadIdx.x;
  if (i < len) data[i] += block_sums[blockIdx.x];
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock)
    sort_scatter_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, const unsigned int n, const int shift,
                        const unsigned int n_tiles, const uint32_t *__restrict__ offsets, uint32_t *__restrict__ keys_out,
                        uint32_t *__restrict__ vals_out) {
  __shared__ unsigned int s_off[kWarpsPerBlock][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned int tile = blockIdx.x * kWarpsPerBlock + warp;
  if (tile >= n_tiles) return;
  for (int i = lane; i < 256; i += 32) s_off[warp][i] = offsets[(size_t)i * n_tiles + tile];
  __syncwarp();
  const unsigned int base = tile * kTile, end = min(n, base + kTile);
  for (unsigned int i = base + lane; i - lane < end; i += 32) {
    const bool live = i < end;
    const uint32_t key = live ? __ldg(keys + i) : 0u;
    const unsigned int d = live ? (key >> shift) & 255u : 256u + lane;
    const unsigned int peers = __match_any_sync(0xffffffffu, d);
    unsigned int pos = 0;
    if (live) pos = s_off[warp][d] + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();  // everyone has read the running offset of its digit
    if (live && lane == __ffs(peers) - 1) s_off[warp][d] += __popc(peers);
    __syncwarp();
    if (live) {
      keys_out[pos] = key;
      vals_out[pos] = vals ? __ldg(vals + i) : i;  // first pass: the value is the item's own index
    }
  }
}

}  // namespace

// Scratch: histogram (256 x tiles words) + block sums of the prefix + one ping-pong pair of key/value arrays.
size_t stable_sort_temp_bytes(size_t n) {
  const size_t tiles = (n + kTile - 1) / kTile, len = 256 * tiles, blocks = (len + kScanBlock - 1) / kScanBlock;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  return up(len * 4) + up(blocks * 4 + 4) + 2 * up(n * 4) + 256;
}

// keys_out ascending (stable), order_out = the items' original positions.  key_bits: significant bits of the keys.
cudaError_t launch_stable_sort_by_cell(const uint32_t *keys, uint32_t *keys_out, uint32_t *order_out, size_t n, int key_bits, void *tmp,
                                       size_t tmp_bytes, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  if (tmp_bytes < stable_sort_temp_bytes(n) || n > 0xffffffffull) return cudaErrorInvalidValue;
  const unsigned int tiles = (unsigned int)((n + kTile - 1) / kTile);
  const unsigned int len = 256u * tiles, scan_blocks = (len + kScanBlock - 1) / kScanBlock;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  char *p = (char *)tmp;
  uint32_t *hist = (uint32_t *)p;
  p += up((size_t)len * 4);
  uint32_t *block_sums = (uint32_t *)p;
  p += up((size_t)scan_blocks * 4 + 4);
  uint32_t *k_tmp = (uint32_t *)p;
  p += up(n * 4);
  uint32_t *v_tmp = (uint32_t *)p;
  const int passes = key_bits <= 8 ? 1 : key_bits <= 16 ? 2 : key_bits <= 24 ? 3 : 4;
  const unsigned int tile_blocks = (tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
  // ping-pong so that the last pass lands in (keys_out, order_out)
  const uint32_t *k_in = keys, *v_in = nullptr;
  for (int pass = 0; pass < passes; pass++) {
    const bool to_final = ((passes - 1 - pass) & 1) == 0;
    uint32_t *k_dst = to_final ? keys_out : k_tmp, *v_dst = to_final ? order_out : v_tmp;
    sort_hist_kernel<<<tile_blocks, 32 * kWarpsPerBlock, 0, stream>>>(k_in, (unsigned int)n, 8 * pass, tiles, hist);
    sort_scan1_kernel<<<scan_blocks, kScanBlock, 0, stream>>>(hist, len, block_sums);
    sort_scan2_kernel<<<1, kScanBlock, 0, stream>>>(block_sums, scan_blocks);
    sort_scan3_kernel<<<scan_blocks, kScanBlock, 0, stream>>>(hist, len, block_sums);
    sort_scatter_kernel<<<tile_blocks, 32 * kWarpsPerBlock, 0, stream>>>(k_in, v_in, (unsigned int)n, 8 * pass, tiles, hist, k_dst, v_dst);
    for (int i = 0; i < 5; i++) count_launch();
    k_in = k_dst;
    v_in = v_dst;
  }
  return cudaGetLastError();
}

}  // namespace qb
