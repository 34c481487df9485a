// SECOND DRAFT (structure-of-arrays records, two speculative rounds; see fast_exact_draft.cu for the first, which is the
// one checked in profiles/r1_fast_exact_draft_check.txt).
// DRAFT - NOT PART OF libqb200.  Checked once on a B200 by fast_exact_test.cu (bit-identical to the floating-point loop,
// profiles/r1_fast_exact_draft_check.txt); untuned.  Compiles with
//   nvcc -gencode arch=compute_100a,code=sm_100a -c tools/research/fast_exact_draft.cu -o /dev/null
// Kernel-level restatement of tools/research/kahan_segments.c (validated on the CPU): the reference's compensated
// member sums (Solution::sumInArea, /root/reference/src/Quantizer.cpp:59-70) of SCALED lattice vectors, bit for bit,
// without the sequential chain of qb200_exact.cu.  One CHAIN per (cell, dimension); members of a cell are the
// positions [beg, end) of the stably sorted list `order` (null: the natural order, K = 1).
//
//   fx_head_kernel      per chain: the floating-point loop until sum >= 4, then integer steps up to the first anchor
//                       (an addend 128 <= t <= 254); short chains are finished here.
//   fx_bounds_kernel    per (chain, nominal segment q): the first anchor at or after head + q*C -> segment boundary,
//                       and the exact 128-bit sum of X over the segment.
//   fx_prefix_kernel    per chain: E_q = A_head + sum of the earlier segments (no rounding corrections).
//   fx_runs_kernel      per (chain, segment): the four speculative runs (start residue 0/128/256/384 mod 512) from E_q,
//                       with the margin of their t = 255 decisions.
//   fx_chain_kernel     per chain: apply the summaries in order, re-run a segment when the true start is farther from
//                       the speculative one than its margin; RN53 of the result is the reference's sum.
#include <cstdint>
#include <cuda_runtime.h>

namespace fx2 {

typedef unsigned __int128 u128;
typedef __int128 i128;

struct Tables { unsigned long long X[256], U[256]; };  // X_t = fl(t/255) * 2^60, U_t = ulp(fl(t/255)) * 2^60 (host-filled)
__constant__ Tables c_tab;

__device__ __forceinline__ int bitlen(u128 a) {
  const unsigned long long hi = (unsigned long long)(a >> 64), lo = (unsigned long long)a;
  return hi ? 128 - __clzll((long long)hi) : (lo ? 64 - __clzll((long long)lo) : 0);
}
__device__ __forceinline__ u128 rn53(u128 A, u128 &U) {
  const int n = bitlen(A);
  if (n <= 53) { U = 1; return A; }
  const int sh = n - 53;
  U = (u128)1 << sh;
  u128 q = A >> sh;
  const u128 r = A & (U - 1), h = U >> 1;
  if (r > h || (r == h && (q & 1))) q++;
  return q << sh;
}
// v to a multiple of the power of two u; ties so that (base + result) / u is even
__device__ __forceinline__ u128 round_even(u128 v, unsigned long long u, u128 base) {
  const u128 r = v & (u128)(u - 1), lo = v - r;
  if (2 * r < u) return lo;
  if (2 * r > u) return lo + u;
  return ((((base + lo) / u) & 1) == 0) ? lo : lo + u;
}
__device__ __forceinline__ u128 step_int(u128 A, int t) {
  if (t == 0) return A;
  const unsigned long long X = c_tab.X[t];
  if (t < 255) return X + round_even(A, c_tab.U[t], X);
  u128 U;
  const u128 s = rn53(A, U);
  if (A == s) return A + X;
  if (A > s) return s + round_even(X + (A - s), 256, 0);
  return s + round_even(X - (s - A), 128, 0);
}
__device__ __forceinline__ u128 decision_margin(u128 A) {
  const int n = bitlen(A);
  u128 d = 0;
  if (n > 53) {
    const u128 U = (u128)1 << (n - 53), r = A & (U - 1), h = U >> 1, a = r > h ? r - h : h - r, b = U - r;
    d = r < a ? r : a;
    if (b < d) d = b;
  }
  const u128 lo = A - ((u128)1 << (n - 1)), hi = ((u128)1 << n) - A;
  if (lo < d) d = lo;
  if (hi < d) d = hi;
  return d;
}

struct Chains {
  const uint8_t *dense;         // local vector v: dense + v * stride, raw image bytes (t = byte ^ 0x80)
  unsigned int stride;
  const uint32_t *order;        // sorted member list (null: identity)
  const uint32_t *cell_beg;     // K + 1 positions into `order`
  int K, dim;
  int seg_len;                  // nominal segment length C
  const uint32_t *seg_off;      // K + 1: first nominal segment of every cell (per dimension the same count)
};
__device__ __forceinline__ int addend(const Chains &c, unsigned int pos, int e) {
  const unsigned int v = c.order ? c.order[pos] : pos;
  return (int)(c.dense[(size_t)v * c.stride + e] ^ 0x80u);
}

// Segment records, structure of arrays.  Record index of (cell k, dimension e, nominal segment q):
//   rec = seg_off[k] * dim + e * (seg_off[k+1] - seg_off[k]) + q        (contiguous along a chain)
struct Segs {
  unsigned int *begin, *end;    // members [begin, end)
  unsigned char *has255, *invalid;
  u128 *sumX, *est;             // exact sum of X; start estimate the runs are made from (round 1: the prefix E)
  u128 *start, *margin;         // [rec * 4 + residue]
  i128 *delta;                  // [rec * 4 + residue]
};
__device__ __forceinline__ long long rec_of(const Chains &c, int k, int e, unsigned int q) {
  return (long long)c.seg_off[k] * c.dim + (long long)e * (c.seg_off[k + 1] - c.seg_off[k]) + q;
}

// per chain: A after the head, position where the segments start; chains that never reach an anchor are done here
__global__ void fx_head_kernel(const Chains c, u128 *A_head, unsigned int *head_end, double *sum_out) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= c.K * c.dim) return;
  const int k = chain / c.dim, e = chain - k * c.dim;
  unsigned int pos = c.cell_beg[k];
  const unsigned int end = c.cell_beg[k + 1];
  double s = 0.0, cc = 0.0;
  while (pos < end && s < 4.0) {  // the reference's loop itself (no contraction: explicit rn intrinsics)
    const double x = __ddiv_rn((double)addend(c, pos, e), 255.0), y = __dsub_rn(x, cc), t = __dadd_rn(s, y);
    cc = __dsub_rn(__dsub_rn(t, s), y);
    s = t;
    pos++;
  }
  if (pos == end) {               // short chain: finished in floating point
    sum_out[chain] = s;
    head_end[chain] = end;
    A_head[chain] = 0;
    return;
  }
  // A = (s - c) * 2^60: s (< 8, ulp 2^-50) and c are multiples of 2^-60, so both products are integers below 2^63
  u128 A = (u128)((i128)__double2ll_rn(ldexp(s, 60)) - (i128)__double2ll_rn(ldexp(cc, 60)));
  while (pos < end) {             // up to and including the first anchor
    const int t = addend(c, pos, e);
    A = step_int(A, t);
    pos++;
    if (t >= 128 && t <= 254) break;
  }
  A_head[chain] = A;
  head_end[chain] = pos;
  if (pos == end) {
    u128 U;
    const u128 r = rn53(A, U);
    sum_out[chain] = ldexp((double)(unsigned long long)(r >> 32), 32 - 60) + ldexp((double)(unsigned long long)(r & 0xffffffffu), -60);
  }
}

// first anchor boundary at or after nominal position p (returns the position AFTER the anchor, or `end`)
__device__ __forceinline__ unsigned int next_boundary(const Chains &c, unsigned int p, unsigned int end, int e) {
  while (p < end) {
    const int t = addend(c, p, e);
    p++;
    if (t >= 128 && t <= 254) break;
  }
  return p;
}

__global__ void fx_bounds_kernel(const Chains c, const unsigned int *head_end, Segs sg) {
  const long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)c.seg_off[c.K] * c.dim;
  if (id >= total) return;
  const unsigned int sq = (unsigned int)(id / c.dim);
  const int e = (int)(id - (long long)sq * c.dim);
  int lo = 0, hi = c.K;           // cell of nominal segment sq
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (c.seg_off[mid] <= sq) lo = mid; else hi = mid; }
  const int k = lo, q = (int)(sq - c.seg_off[k]);
  const unsigned int end = c.cell_beg[k + 1], head = head_end[k * c.dim + e];
  const long long rec = rec_of(c, k, e, (unsigned int)q);
  const unsigned long long nominal = (unsigned long long)head + (unsigned long long)q * c.seg_len;
  unsigned int b = q == 0 ? head : (nominal - 1 < end ? next_boundary(c, (unsigned int)(nominal - 1), end, e) : end);
  unsigned int f = nominal + c.seg_len - 1 < end ? next_boundary(c, (unsigned int)(nominal + c.seg_len - 1), end, e) : end;
  if (b > f) b = f;
  u128 sx = 0;
  int h = 0;
  for (unsigned int p = b; p < f; p++) {
    const int t = addend(c, p, e);
    sx += c_tab.X[t];
    h |= t == 255;
  }
  sg.begin[rec] = b; sg.end[rec] = f; sg.sumX[rec] = sx; sg.has255[rec] = (unsigned char)h; sg.invalid[rec] = 1;
}

__global__ void fx_prefix_kernel(const Chains c, const u128 *A_head, Segs sg) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= c.K * c.dim) return;
  const int k = chain / c.dim, e = chain - k * c.dim;
  u128 E = A_head[chain];
  const unsigned int cnt = c.seg_off[k + 1] - c.seg_off[k];
  const long long r0 = rec_of(c, k, e, 0);
  for (unsigned int q = 0; q < cnt; q++) {
    sg.est[r0 + q] = E;
    E += sg.sumX[r0 + q];
  }
}

// the four speculative runs of every segment still marked invalid, from its current start estimate
__global__ void fx_runs_kernel(const Chains c, Segs sg) {
  const long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // (record, residue)
  const long long total = (long long)c.seg_off[c.K] * c.dim * 4;
  if (id >= total) return;
  const int r = (int)(id & 3);
  const long long rec = id >> 2;
  if (!sg.invalid[rec] || sg.begin[rec] >= sg.end[rec]) return;
  // dimension of this record: records of cell k are dim blocks of cnt_k
  const unsigned int sq = (unsigned int)(rec / c.dim);  // some nominal segment of the same cell
  int lo = 0, hi = c.K;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (c.seg_off[mid] <= sq) lo = mid; else hi = mid; }
  const unsigned int cnt = c.seg_off[lo + 1] - c.seg_off[lo];
  const int e = (int)((rec - (long long)c.seg_off[lo] * c.dim) / cnt);
  const u128 E = sg.est[rec];
  int corr = (int)((128u * r + 512u - (unsigned int)(E & 511)) & 511u);
  if (corr > 256) corr -= 512;
  u128 a = (u128)((i128)E + corr), mg = ~(u128)0;
  const u128 a0 = a;
  for (unsigned int p = sg.begin[rec]; p < sg.end[rec]; p++) {
    const int t = addend(c, p, e);
    if (t == 255) { const u128 d = decision_margin(a); if (d < mg) mg = d; }
    a = step_int(a, t);
  }
  sg.start[rec * 4 + r] = a0; sg.delta[rec * 4 + r] = (i128)a - (i128)a0; sg.margin[rec * 4 + r] = mg;
}

// final == 0: apply every summary (valid or not), leave the running A as the next start estimate of segments whose
//             margin did not cover it, and mark them for a second round of runs;
// final != 0: exact - a segment that is still not covered is executed sequentially.
__global__ void fx_chain_kernel(const Chains c, const u128 *A_head, const unsigned int *head_end, Segs sg, const int final,
                                double *sum_out, unsigned int *reruns) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= c.K * c.dim) return;
  const int k = chain / c.dim, e = chain - k * c.dim;
  if (head_end[chain] == c.cell_beg[k + 1]) return;  // finished by the head kernel
  u128 A = A_head[chain];
  const unsigned int cnt = c.seg_off[k + 1] - c.seg_off[k];
  const long long r0 = rec_of(c, k, e, 0);
  for (unsigned int q = 0; q < cnt; q++) {
    const long long rec = r0 + q;
    if (sg.begin[rec] >= sg.end[rec]) { sg.invalid[rec] = 0; continue; }
    const int r = (int)((A & 511) >> 7);
    const i128 shift = (i128)A - (i128)sg.start[rec * 4 + r];
    const u128 mag = shift < 0 ? (u128)(-shift) : (u128)shift;
    const bool ok = !sg.has255[rec] || mag < sg.margin[rec * 4 + r];
    if (!final) {
      sg.invalid[rec] = ok ? 0 : 1;
      if (!ok) sg.est[rec] = A;
      A = (u128)((i128)A + sg.delta[rec * 4 + r]);
      A &= ~(u128)127;               // a wrong t = 255 branch may leave an estimate off the 128 grid
    } else if (ok) {
      A = (u128)((i128)A + sg.delta[rec * 4 + r]);
    } else {
      for (unsigned int p = sg.begin[rec]; p < sg.end[rec]; p++) A = step_int(A, addend(c, p, e));
      atomicAdd(reruns, 1u);
    }
  }
  if (final) {
    u128 U;
    const u128 rr = rn53(A, U);
    sum_out[chain] = ldexp((double)(unsigned long long)(rr >> 32), 32 - 60) + ldexp((double)(unsigned long long)(rr & 0xffffffffu), -60);
  }
}

}  // namespace fx2
