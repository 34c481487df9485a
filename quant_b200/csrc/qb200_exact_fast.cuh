// Parallel, bit-exact evaluation of the reference's compensated member sums on SCALED lattice vectors.
//
// The reference sums a cell's members (values t/255.0, t = byte ^ 0x80 in 0..255) with a Kahan loop in ascending
// vector order (Solution::sumInArea / trainingSetSum, /root/reference/src/Quantizer.cpp:46-70):
//     y = x - c;  t = sum + y;  c = (t - sum) - y;  sum = t
// and the centroid is that sum divided by the member count.  The loop is a dependent chain of four FP64
// operations per member; executed literally (kahan_sums_kernel, qb200_exact.cu) a 4.2 M-vector train costs ~270 ms.
// This file evaluates the SAME chain in parallel and returns the SAME bits.
//
// Integer model (validated against the floating-point loop step by step, tools/research/kahan_automaton.py).
// With g = 2^-60, X_t = fl(t/255)/g is an integer multiple of u_t = 2^lev(t), lev(t) = floor(log2 t), and once
// sum >= 4 (Fast2Sum exact, ulp(sum) >= 2 ulp(y)) the pair (sum, c) is a function of ONE integer
//     A = (sum - c)/g,      sum = RN53(A),  c = sum - A,
// and an addend t maps
//     t = 0        A unchanged
//     1..254       A <- X_t + round(A to a multiple of u_t; ties so that the result/u_t is even)
//     255          x = 1.0 sits at a binade edge: y = fl(1 - c) is on the grid 2^8 when c < 0, 2^7 when c > 0:
//                  A <- RN53(A) + round(X_255 + (A - RN53(A)) to that grid, ties to even)
// So a step looks at A only through its low bits (bits 0..lev(t); bits 0..8 for t = 255) - except for the SIGN of
// c at t = 255, which depends on where A sits inside ulp(sum).  After a step with t >= 1 the state is a multiple of
// 2^tz(t), tz = lev(t) (7 for t = 255).
//
// Decomposition.  A chain (one cell, one dimension; members in sorted order) is cut into nominal windows of C
// members.  A segment starts right after an ANCHOR: the first member, among the first Wn of its window, with the
// largest tz found there (je).  Entering the segment the state is a multiple of 2^je; if `top` is the highest bit
// any member of the segment looks at, the segment's effect depends on the entry state only through bits je..top -
// ncls = 2^(top - je + 1) classes (1 when top < je; typically 2 or 4: natural data has members with t >= 128 all
// over, bright areas have none of them and a small `top`).  Every segment is run speculatively for each class from
// E = (exact sum of the X before it, 128-bit) adjusted to the class, recording
//     dcorr[cls]  = (A_end - A_start) - sum of the segment's X      (the rounding corrections: a small integer)
//     margin[cls] = how far the start may move (in multiples of 2^(top+1)) before a t = 255 decision flips.
// A short sequential pass per chain then applies the summaries (W = A - E is a small signed integer; the class is
// read off (E + W) mod 2^(top+1)) and re-runs, exactly, the rare segment whose margin does not cover W, and segments
// with more than kFxMaxCls classes.  The head of a chain (until sum >= 4 and up to the first boundary after that)
// runs the floating-point loop / the integer steps sequentially.
//
// Everything here is __host__ __device__ so that the same code is checked on the CPU against the floating-point
// loop (tests/cpp/exact_fast_test.cpp) and runs unchanged in the kernels of qb200_exact.cu.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define QB_HD __host__ __device__ __forceinline__
#else
#define QB_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define QB_UNROLL _Pragma("unroll")
#else
#define QB_UNROLL
#endif

namespace qb {
namespace fx {

typedef unsigned __int128 u128;
typedef __int128 i128;

constexpr int kFxMaxCls = 8;   // speculative classes per segment; beyond that the segment is run sequentially
// members searched for an anchor at the start of a nominal window (16 was tried: a fifth less staging, but cells whose
// high-level members are sparse then enter with a low je, their batches stop being composable and chaining slows 3-5x)
constexpr int kFxAnchorWin = 64;

struct Tables {
  unsigned long long X[256];  // X_t = fl(t/255) * 2^60 (host-filled: fx_fill_tables)
};

inline void fx_fill_tables(Tables &tab) {
  tab.X[0] = 0;
  for (int t = 1; t < 256; t++) {
    const double x = (double)t / 255.0;  // correctly rounded IEEE division, as the reference computes it
    // x * 2^60 is an integer below 2^61 (x has 53 significant bits, x >= 2^-8)
    double scaled = x;
    for (int i = 0; i < 60; i++) scaled *= 2.0;
    tab.X[t] = (unsigned long long)scaled;
  }
}

QB_HD int fx_clz64(unsigned long long v) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return v ? __builtin_clzll(v) : 64;
#endif
}
QB_HD int fx_lev(int t) {  // floor(log2 t), t >= 1
#if defined(__CUDA_ARCH__)
  return 31 - __clz(t);
#else
  return 31 - __builtin_clz((unsigned)t);
#endif
}
// highest bit of A the step with addend t looks at (-1: none)
QB_HD int fx_need(int t) { return t == 0 ? -1 : (t == 255 ? 8 : fx_lev(t)); }
// trailing zero bits of A guaranteed after the step (-1: the step changes nothing)
QB_HD int fx_tz(int t) { return t == 0 ? -1 : (t == 255 ? 7 : fx_lev(t)); }

QB_HD int fx_bitlen(u128 a) {
  const unsigned long long hi = (unsigned long long)(a >> 64), lo = (unsigned long long)a;
  return hi ? 128 - fx_clz64(hi) : (lo ? 64 - fx_clz64(lo) : 0);
}

// RN53 of A (A >= 2^62, so the shift is at least 10 and at most 75): returns s and the signed remainder d = A - s.
QB_HD u128 fx_rn53(u128 A, long long &d) {
  const int sh = fx_bitlen(A) - 53;
  if (sh <= 0) { d = 0; return A; }
  const u128 U = (u128)1 << sh, r = A & (U - 1), h = U >> 1;
  const bool up = r > h || (r == h && ((A >> sh) & 1));
  d = up ? -(long long)(unsigned long long)(U - r) : (long long)(unsigned long long)r;
  return up ? A - r + U : A - r;
}

// v >= 0 rounded to a multiple of the power of two u, ties to an even multiple
QB_HD long long fx_round_even(long long v, long long u) {
  const long long r = v & (u - 1), lo = v - r;
  if (2 * r < u) return lo;
  if (2 * r > u) return lo + u;
  return ((lo / u) & 1) ? lo + u : lo;
}

QB_HD u128 fx_step(u128 A, int t, const Tables &tab) {
  if (t == 0) return A;
  const unsigned long long X = tab.X[t];
  if (t < 255) {
    const int lev = fx_lev(t);
    const unsigned long long u = 1ull << lev, lo = (unsigned long long)A, r = lo & (u - 1);
    // ties: the result (X + rounded A)/u must be even; X and the rounded-down A are multiples of u
    const bool up = 2 * r > u || (2 * r == u && (((X >> lev) ^ (lo >> lev)) & 1));
    return A + (X - r + (up ? u : 0ull));
  }
  long long d;
  const u128 s = fx_rn53(A, d);
  if (d == 0) return A + X;
  // y = fl(1 - c), c = -d g: above 1 the grid is 2^-52 (2^8 g), below 1 it is 2^-53 (2^7 g)
  const long long y = fx_round_even((long long)X + d, d > 0 ? 256 : 128);
  return s + (u128)(unsigned long long)y;
}

// Distance from A to the nearest point where the t = 255 branch changes (sign of A - RN53(A), or ulp(A)), saturated.
QB_HD unsigned long long fx_decision_margin(u128 A) {
  const int n = fx_bitlen(A), sh = n - 53;
  u128 dist = 0;
  if (sh > 0) {
    const u128 U = (u128)1 << sh, r = A & (U - 1), h = U >> 1, a = r > h ? r - h : h - r, b = U - r;
    dist = r < a ? r : a;
    if (b < dist) dist = b;
  }
  const u128 lo = A - ((u128)1 << (n - 1)), hi = ((u128)1 << n) - A;
  if (lo < dist) dist = lo;
  if (hi < dist) dist = hi;
  return dist > (u128)0xffffffffffffffffull ? 0xffffffffffffffffull : (unsigned long long)dist;
}

// double (>= 0, a multiple of 2^-60, below 2^67) -> integer multiple of g = 2^-60
QB_HD u128 fx_from_double(double v) {
  union { double d; unsigned long long u; } cv;
  cv.d = v;
  const unsigned long long bits = cv.u;
  const int e = (int)((bits >> 52) & 0x7ff);
  if (e == 0) return 0;  // zero (denormals are not multiples of 2^-60)
  const unsigned long long m = (bits & 0xfffffffffffffull) | (1ull << 52);
  const int sh = e - 1075 + 60;  // v = m * 2^(e - 1075)
  return sh >= 0 ? (u128)m << sh : (u128)(m >> (-sh));
}
// integer with at most 53 significant bits -> double (exact)
QB_HD double fx_to_double(u128 A) {
  const double two32 = 4294967296.0, g = 1.0 / (1024.0 * 1024.0 * 1024.0) / (1024.0 * 1024.0 * 1024.0);  // 2^-60
  return ((double)(unsigned long long)(A >> 32) * two32 + (double)(unsigned long long)(A & 0xffffffffu)) * g;
}

// The pair (sum, c) of a state A in the regime.
QB_HD void fx_state_to_pair(u128 A, double &sum, double &c) {
  long long d;
  const u128 s = fx_rn53(A, d);
  const double g = 1.0 / (1024.0 * 1024.0 * 1024.0) / (1024.0 * 1024.0 * 1024.0);
  sum = fx_to_double(s);
  c = d == 0 ? 0.0 : -(double)d * g;  // |d| < 2^40: exact; the loop's (t - sum) - y is +0 when both agree
}
// A = (sum - c)/g for a pair in the regime (sum >= 4).
QB_HD u128 fx_pair_to_state(double sum, double c) {
  const u128 s = fx_from_double(sum);
  return c >= 0 ? s - fx_from_double(c) : s + fx_from_double(-c);
}

// One iteration of the reference's loop itself (src/Quantizer.cpp:64-67), no contraction.
QB_HD void fx_fp_step(double &sum, double &c, int t) {
#if defined(__CUDA_ARCH__)
  const double x = __ddiv_rn((double)t, 255.0), y = __dsub_rn(x, c), s2 = __dadd_rn(sum, y);
  c = __dsub_rn(__dsub_rn(s2, sum), y);
  sum = s2;
#else
  volatile double x = (double)t / 255.0;
  volatile double y = x - c;
  volatile double s2 = sum + y;
  volatile double d1 = s2 - sum;
  c = d1 - y;
  sum = s2;
#endif
}

// ---- segmentation --------------------------------------------------------------------------------------------
// Boundary of nominal window starting at position P (P < end): the position right after the first member, among
// the first kFxAnchorWin of the window, with the largest tz found there; je = that tz (-1: only zeros, the boundary
// is P itself and nothing is known about the entry state).  xsum = sum of X over [P, boundary).
struct Anchor {
  unsigned int b;
  int je;
  unsigned long long xsum_lo;  // at most 64 * 2^60 < 2^67: low 64 bits and the carry bits
  unsigned int xsum_hi;
};
template <typename Acc>
QB_HD Anchor fx_anchor(const Acc &acc, unsigned int P, unsigned int end, const Tables &tab) {
  Anchor a;
  a.b = P;
  a.je = -1;
  a.xsum_lo = 0;
  a.xsum_hi = 0;
  u128 run = 0;
  const unsigned int stop = end - P > (unsigned)kFxAnchorWin ? P + kFxAnchorWin : end;
  for (unsigned int p = P; p < stop; p++) {
    const int t = acc(p);
    run += tab.X[t];
    const int z = fx_tz(t);
    if (z > a.je) {
      a.je = z;
      a.b = p + 1;
      a.xsum_lo = (unsigned long long)run;
      a.xsum_hi = (unsigned int)(run >> 64);
      if (z == 7) break;
    }
  }
  return a;
}

// Per (window, dimension) record written by the speculative runs and read by the chaining pass.
struct SegRecord {
  u128 Eb;                        // E at the segment's first member, plus w_base
  unsigned int margin[kFxMaxCls]; // per class: |W - w_base - cc| must stay below this (0xffffffff: no t = 255 in the segment)
  short dcorr[kFxMaxCls];         // per class: (A_end - A_start) - sum of X over the segment (|.| <= 128 per member)
  unsigned int begin, end;        // members [begin, end)
  int w_base;                     // the speculative runs were made around E + w_base (0 in the first round; the refinement
                                  // round uses the first chaining pass's estimate of W, see kFxRefine)
  unsigned short xs_low;          // (sum of X over the segment) mod 2^16
  signed char je, top;            // entry trailing zeros (>= 0) and highest bit looked at (-1: nothing but zeros)
  signed char ncls;               // 1..kFxMaxCls, or 0: run sequentially
  unsigned char flags;            // kFxRefine: the first chaining pass found W outside this segment's margin
};
constexpr unsigned char kFxRefine = 1;
constexpr long long kFxWLimitHost = 1ll << 29;
static_assert(sizeof(SegRecord) == 96, "SegRecord is copied as six 16-byte words");

// centred representative of v modulo m (a power of two): in (-m/2, m/2]
QB_HD long long fx_centered(long long v, long long m) {
  v &= m - 1;
  return v > m / 2 ? v - m : v;
}

// Geometry of a segment's classes: the entry state is a multiple of 2^je and its members look at bits <= top, so
// it matters modulo `mod` = 2^max(top + 1, je) only, i.e. through ncls = mod >> je classes (bits je..top).
QB_HD long long fx_class_mod(int je, int top) { return 1ll << (top + 1 > je ? top + 1 : je); }

// Speculative runs of the segment [begin, end) entered with trailing zeros je, whose members look at bits <= top.
template <typename Acc>
QB_HD void fx_run_segment(const Acc &acc, const Tables &tab, SegRecord &rec) {
  const int je = rec.je;
  const long long mod = fx_class_mod(je, rec.top);
  if ((mod >> je) > kFxMaxCls) {
    rec.ncls = 0;
    return;
  }
  rec.ncls = (signed char)(mod >> je);
  const long long e_low = (long long)((unsigned long long)rec.Eb & (unsigned long long)(mod - 1));
  for (int cls = 0; cls < rec.ncls; cls++) {
    const long long cc = fx_centered(((long long)cls << je) - e_low, mod);  // nearest start of this class to E
    u128 a = (u128)((i128)rec.Eb + cc);
    const u128 a0 = a;
    u128 xs = 0;
    unsigned long long mg = 0xffffffffffffffffull;
    for (unsigned int p = rec.begin; p < rec.end; p++) {
      const int t = acc(p);
      if (t == 255) {
        const unsigned long long dm = fx_decision_margin(a);
        if (dm < mg) mg = dm;
      }
      xs += tab.X[t];
      a = fx_step(a, t, tab);
    }
    rec.dcorr[cls] = (short)(long long)((i128)(a - a0) - (i128)xs);
    rec.margin[cls] = mg == 0xffffffffffffffffull ? 0xffffffffu : (mg > 0xfffffffeull ? 0xfffffffeu : (unsigned int)mg);
  }
}

// Chaining step: applies a segment's summary to W = A - E (A: true state at the segment's first member).  Returns
// false when the segment has to be run sequentially (fx_rerun).
QB_HD bool fx_apply(const SegRecord &rec, long long &W) {
  if (rec.begin >= rec.end) return true;
  if (rec.ncls == 0) return false;
  const long long mod = fx_class_mod(rec.je, rec.top);
  const long long e_low = (long long)((unsigned long long)rec.Eb & (unsigned long long)(mod - 1));
  const long long Wrel = W - rec.w_base;              // rec.Eb already contains w_base
  const long long a_low = (e_low + Wrel) & (mod - 1); // true state modulo `mod` (two's complement for negative values)
  if ((a_low & ((1ll << rec.je) - 1)) != 0) return false;  // cannot happen; run exactly rather than trust it
  const int cls = (int)(a_low >> rec.je);
  const long long cc = fx_centered(((long long)cls << rec.je) - e_low, mod);
  const long long shift = Wrel - cc;                  // a multiple of `mod`: the run's decisions at t < 255 carry over
  const unsigned long long mag = (unsigned long long)(shift < 0 ? -shift : shift);
  // shift == 0: the speculative run started at the true state itself (flat data, where no step ever rounds)
  if (rec.margin[cls] != 0xffffffffu && mag != 0 && mag >= (unsigned long long)rec.margin[cls]) return false;
  W += rec.dcorr[cls];
  return true;
}
// The same speculative runs, all classes in ONE pass over the members.  Steps with t < 255 look at the low bits of
// the state only, so per class it is enough to carry the state modulo 2^16 and the accumulated rounding
// correction; the addend's table lookup, level and the (128-bit) sum of X are shared by the classes.  A t = 255
// step (binade edge) needs where the state sits inside ulp(sum): the exponent comes from the shared E + sum of X
// (the classes are within 2^16 of it; next to a power of two the general step is taken instead), the position
// from the class's low 64 bits.  NC = unrolled class slots (4 covers the usual je = 7 segments).
template <int NC, typename Acc>
QB_HD void fx_run_classes(const Acc &acc, const Tables &tab, SegRecord &rec, const int ncls, const long long mod) {
  const int je = rec.je;
  const long long e_low = (long long)((unsigned long long)rec.Eb & (unsigned long long)(mod - 1));
  unsigned int lo16[NC];              // state mod 2^16
  int cc[NC];                         // start offset of the class relative to E
  int corr[NC];                       // accumulated rounding corrections
  unsigned long long mg[NC];
QB_UNROLL
  for (int cls = 0; cls < NC; cls++) {
    cc[cls] = cls < ncls ? (int)fx_centered(((long long)cls << je) - e_low, mod) : 0;
    lo16[cls] = (unsigned int)((unsigned long long)((i128)rec.Eb + cc[cls])) & 0xffffu;
    corr[cls] = 0;
    mg[cls] = 0xffffffffffffffffull;
  }
  u128 xs = 0;
  for (unsigned int p = rec.begin; p < rec.end; p++) {
    const int t = acc(p);
    if (t == 0) continue;
    const unsigned long long X = tab.X[t];
    if (t < 255) {
      const int lev = fx_lev(t);
      const unsigned int u = 1u << lev, xl = (unsigned int)X & 0xffffu, xpar = (unsigned int)(X >> lev);
QB_UNROLL
      for (int cls = 0; cls < NC; cls++) {
        if (cls < ncls) {
          const unsigned int a = lo16[cls], r = a & (u - 1);
          const bool up = 2 * r > u || (2 * r == u && ((xpar ^ (a >> lev)) & 1u));
          const int d = (int)(up ? u : 0u) - (int)r;
          corr[cls] += d;
          lo16[cls] = (a + xl + (unsigned int)d) & 0xffffu;
        }
      }
    } else {
      const u128 full = rec.Eb + xs;
      const int n = fx_bitlen(full), sh = n - 53;
      const u128 lo_edge = full - ((u128)1 << (n - 1)), hi_edge = ((u128)1 << n) - full;
      const u128 edge128 = lo_edge < hi_edge ? lo_edge : hi_edge;
      if (sh < 10 || sh > 48 || edge128 < ((u128)1 << 17)) {  // next to a power of two (or outside the regime): general step
QB_UNROLL
        for (int cls = 0; cls < NC; cls++) {
          if (cls < ncls) {
            const u128 a = (u128)((i128)full + cc[cls] + corr[cls]);
            const unsigned long long dm = fx_decision_margin(a);
            if (dm < mg[cls]) mg[cls] = dm;
            const u128 a2 = fx_step(a, 255, tab);
            corr[cls] += (int)(long long)((i128)(a2 - a) - (i128)X);
            lo16[cls] = (unsigned int)(unsigned long long)a2 & 0xffffu;
          }
        }
      } else {
        const unsigned long long U = 1ull << sh, h = U >> 1, full_lo = (unsigned long long)full;
        const long long lo_e = lo_edge > (u128)0x3fffffffffffffffull ? 0x3fffffffffffffffll : (long long)lo_edge;
        const long long hi_e = hi_edge > (u128)0x3fffffffffffffffull ? 0x3fffffffffffffffll : (long long)hi_edge;
QB_UNROLL
        for (int cls = 0; cls < NC; cls++) {
          if (cls < ncls) {
            const long long delta = (long long)cc[cls] + corr[cls];
            const unsigned long long a64 = full_lo + (unsigned long long)delta;
            const unsigned long long r = a64 & (U - 1);
            // fx_decision_margin: distance to 0, U/2, U inside the ulp, and to the binade edges
            unsigned long long dist = r < U - r ? r : U - r;
            const unsigned long long dh = r > h ? r - h : h - r;
            dist = dh < dist ? dh : dist;
            const long long le = lo_e == 0x3fffffffffffffffll ? lo_e : lo_e + delta, he = hi_e == 0x3fffffffffffffffll ? hi_e : hi_e - delta;
            if ((unsigned long long)le < dist) dist = (unsigned long long)le;
            if ((unsigned long long)he < dist) dist = (unsigned long long)he;
            if (dist < mg[cls]) mg[cls] = dist;
            // fx_step(t = 255): d = A - RN53(A); y = (X + d) rounded to 2^8 (d > 0) or 2^7 (d < 0), ties to even
            const bool up = r > h || (r == h && ((a64 >> sh) & 1ull));
            const long long d = up ? (long long)r - (long long)U : (long long)r;
            long long step = 0;  // A' - A - X
            if (d != 0) step = fx_round_even((long long)X + d, d > 0 ? 256 : 128) - d - (long long)X;
            corr[cls] += (int)step;
            lo16[cls] = (unsigned int)(a64 + (unsigned long long)X + (unsigned long long)step) & 0xffffu;
          }
        }
      }
    }
    xs += X;
  }
  rec.xs_low = (unsigned short)((unsigned long long)xs & 0xffffu);
QB_UNROLL
  for (int cls = 0; cls < NC; cls++) {
    if (cls < ncls) {
      rec.dcorr[cls] = (short)corr[cls];
      rec.margin[cls] = mg[cls] == 0xffffffffffffffffull ? 0xffffffffu : (mg[cls] > 0xfffffffeull ? 0xfffffffeu : (unsigned int)mg[cls]);
    }
  }
}
template <typename Acc>
QB_HD void fx_run_segment_multi(const Acc &acc, const Tables &tab, SegRecord &rec) {
  const long long mod = fx_class_mod(rec.je, rec.top);
  const int ncls = (int)(mod >> rec.je);
  if (ncls > kFxMaxCls) {
    rec.ncls = 0;
    return;
  }
  rec.ncls = (signed char)ncls;
  if (ncls <= 4)
    fx_run_classes<4>(acc, tab, rec, ncls, mod);
  else
    fx_run_classes<kFxMaxCls>(acc, tab, rec, ncls, mod);
}

// ---- chaining 32 segments at a time -----------------------------------------------------------------------
// When the class bits of a batch of consecutive segments all lie in two adjacent bit positions [jb, jb + 1] (noise
// and natural images: jb = 7, the bits an addend t >= 128 or t = 255 looks at), every segment is a map on the FOUR
// states s = (A >> jb) & 3 of the true state A at its first member:
//     s -> (state after the segment, dW = its rounding corrections, [lo, hi] = the entry W it is valid for).
// Maps compose associatively, so a warp combines 32 of them with a parallel prefix (fx_compose) instead of 32
// dependent steps.  An empty interval means "cannot happen / must be run exactly".
struct Map4 {
  int s[4], dW[4], lo[4], hi[4];
};
constexpr int kFxWLimit = 1 << 29;  // |W| beyond this is handled by the sequential path (32-bit intervals)

QB_HD Map4 fx_map_identity() {
  Map4 m;
  for (int i = 0; i < 4; i++) {
    m.s[i] = i;
    m.dW[i] = 0;
    m.lo[i] = -kFxWLimit;
    m.hi[i] = kFxWLimit;
  }
  return m;
}
// Map of one segment for batch bit position jb; requires rec.je >= jb and max(rec.top, rec.je - 1) <= jb + 1.
QB_HD Map4 fx_map_of(const SegRecord &rec, int jb) {
  if (rec.begin >= rec.end) return fx_map_identity();
  Map4 m;
  const long long mod = fx_class_mod(rec.je, rec.top);
  const long long e_low = (long long)((unsigned long long)rec.Eb & (unsigned long long)(mod - 1));
  for (int s = 0; s < 4; s++) {
    const long long a_low = (long long)s << jb;          // the true state modulo 2^(jb + 2)
    m.s[s] = s;
    m.dW[s] = 0;
    m.lo[s] = 1;                                         // empty until proven usable
    m.hi[s] = 0;
    if (rec.ncls == 0 || (a_low & ((1ll << rec.je) - 1)) != 0) continue;  // sequential segment / impossible entry state
    const int cls = (int)((a_low & (mod - 1)) >> rec.je);
    const long long cc = fx_centered(((long long)cls << rec.je) - e_low, mod);
    const unsigned int mgn = rec.margin[cls];
    long long lo, hi;
    if (mgn == 0xffffffffu) {
      lo = -kFxWLimit;
      hi = kFxWLimit;
    } else {
      const long long rad = mgn > 0 ? (long long)mgn - 1 : 0;  // |W - w_base - cc| < margin, or equal
      lo = cc - rad + rec.w_base;
      hi = cc + rad + rec.w_base;
      if (lo < -kFxWLimit) lo = -kFxWLimit;
      if (hi > kFxWLimit) hi = kFxWLimit;
    }
    m.lo[s] = (int)lo;
    m.hi[s] = (int)hi;
    m.dW[s] = rec.dcorr[cls];
    m.s[s] = (int)(((a_low + (long long)rec.xs_low + (long long)rec.dcorr[cls]) >> jb) & 3);
  }
  return m;
}
// f first, then g
QB_HD Map4 fx_compose(const Map4 &f, const Map4 &g) {
  Map4 h;
  for (int s = 0; s < 4; s++) {
    const int mid = f.s[s];
    h.s[s] = g.s[mid];
    const long long dW = (long long)f.dW[s] + g.dW[mid];
    long long lo = (long long)g.lo[mid] - f.dW[s], hi = (long long)g.hi[mid] - f.dW[s];
    if (f.lo[s] > lo) lo = f.lo[s];
    if (f.hi[s] < hi) hi = f.hi[s];
    if (f.lo[s] > f.hi[s] || g.lo[mid] > g.hi[mid] || dW > kFxWLimit || dW < -kFxWLimit) {  // keep emptiness
      lo = 1;
      hi = 0;
    }
    if (lo < -kFxWLimit) lo = -kFxWLimit;
    if (hi > kFxWLimit) hi = kFxWLimit;
    h.dW[s] = (int)(dW > kFxWLimit ? kFxWLimit : (dW < -kFxWLimit ? -kFxWLimit : dW));
    h.lo[s] = (int)lo;
    h.hi[s] = (int)hi;
  }
  return h;
}
// Can a batch with these extremes (over its non-empty, non-sequential segments) be chained through 4-state maps?
QB_HD bool fx_batch_composable(int je_min, int je_max_unused, int top_max, int &jb) {
  (void)je_max_unused;
  jb = je_min;
  const int hi_bit = top_max > je_min ? top_max : je_min;  // highest class bit of any segment
  return hi_bit <= jb + 1;
}

// First chaining pass, segment not covered by its summary: note W as the point around which the refinement round
// re-runs it, then move on with the summary anyway - W is then only an ESTIMATE (a flipped t = 255 decision is worth a
// few hundred units), which is all the later marks of this pass need; the second pass is exact again.
QB_HD void fx_mark_and_estimate(SegRecord &rec, long long &W) {
  if (rec.begin >= rec.end || rec.ncls == 0) return;  // sequential segments are run exactly in the second pass
  const long long mod = fx_class_mod(rec.je, rec.top);
  const long long e_low = (long long)((unsigned long long)rec.Eb & (unsigned long long)(mod - 1));
  const long long a_low = (e_low + (W - rec.w_base)) & (mod - 1);
  const int cls = (int)(a_low >> rec.je) & (rec.ncls - 1);
  const long long w_new = W;
  W += rec.dcorr[cls];
  if (w_new > -kFxWLimitHost && w_new < kFxWLimitHost) {
    rec.Eb = (u128)((i128)rec.Eb + (w_new - rec.w_base));
    rec.w_base = (int)w_new;
    rec.flags |= kFxRefine;
  }
}

// Exact sequential run of a segment from the true state Eb + W (summary not usable): steps with t < 255 only need
// the state modulo 2^16 and add their rounding correction to W; a t = 255 step rebuilds the full state.
template <typename Acc>
QB_HD void fx_rerun(const Acc &acc, const Tables &tab, const SegRecord &rec, long long &W) {
  const u128 base = (u128)((i128)rec.Eb + (W - rec.w_base));
  unsigned int lo16 = (unsigned int)(unsigned long long)base & 0xffffu;
  long long corr = 0;
  u128 xs = 0;
  for (unsigned int p = rec.begin; p < rec.end; p++) {
    const int t = acc(p);
    if (t == 0) continue;
    const unsigned long long X = tab.X[t];
    if (t < 255) {
      const int lev = fx_lev(t);
      const unsigned int u = 1u << lev, r = lo16 & (u - 1);
      const bool up = 2 * r > u || (2 * r == u && ((((unsigned int)(X >> lev)) ^ (lo16 >> lev)) & 1u));
      const int d = (int)(up ? u : 0u) - (int)r;
      corr += d;
      lo16 = (lo16 + ((unsigned int)X & 0xffffu) + (unsigned int)d) & 0xffffu;
    } else {
      const u128 a = (u128)((i128)base + (i128)xs + corr);
      const u128 a2 = fx_step(a, 255, tab);
      corr += (long long)((i128)(a2 - a) - (i128)X);
      lo16 = (unsigned int)(unsigned long long)a2 & 0xffffu;
    }
    xs += X;
  }
  W += corr;
}

// ---- per-thread bodies of the kernels (shared with the CPU check) ---------------------------------------------
// Window summary: sum of X and the highest bit looked at over members [P, P_end).
template <typename Acc>
QB_HD void fx_window(const Acc &acc, const Tables &tab, unsigned int P, unsigned int P_end, u128 &sumX, int &need_max) {
  u128 s = 0;
  int m = -1;
  for (unsigned int p = P; p < P_end; p++) {
    const int t = acc(p);
    s += tab.X[t];
    const int nd = fx_need(t);
    m = nd > m ? nd : m;
  }
  sumX = s;
  need_max = m;
}

struct HeadOut {
  i128 B;               // E(p) = B + (sum of X over [beg, p)): the trajectory without rounding corrections, E(pos_end) = A
  unsigned int pos_end; // first member the segments cover (right after the head's anchor); the state there is exact
  unsigned int q_start; // window that contains pos_end: the first segment of the chaining pass, entered with W = 0
  int je;               // trailing zeros of the state at pos_end (the head's anchor)
  double sum, c;        // final pair when the head finished the chain
  int done;
};
// Head of a chain [beg, end) with nominal windows of C members: the reference's loop itself from the incoming pair
// (sum, c) until sum >= 4, then integer steps up to the first anchor found after that point (fx_anchor).
template <typename Acc>
QB_HD HeadOut fx_head(const Acc &acc, const Tables &tab, unsigned int beg, unsigned int end, unsigned int C, double sum,
                      double c) {
  HeadOut o;
  o.done = 0;
  unsigned int pos = beg;
  u128 px = 0;
  while (pos < end && sum < 4.0) {
    const int t = acc(pos);
    fx_fp_step(sum, c, t);
    px += tab.X[t];
    pos++;
  }
  o.sum = sum;
  o.c = c;
  o.B = 0;
  o.pos_end = end;
  o.q_start = 0;
  o.je = 0;
  if (sum < 4.0) {
    o.done = 1;
    return o;
  }
  u128 A = fx_pair_to_state(sum, c);
  if (pos < end) {
    const Anchor a = fx_anchor(acc, pos, end, tab);
    for (; pos < a.b; pos++) {
      const int t = acc(pos);
      px += tab.X[t];
      A = fx_step(A, t, tab);
    }
    o.je = a.je < 0 ? 0 : a.je;
  }
  if (pos >= end) {
    fx_state_to_pair(A, o.sum, o.c);
    o.done = 1;
    return o;
  }
  o.B = (i128)A - (i128)px;
  o.pos_end = pos;
  o.q_start = (pos - beg) / C;
  return o;
}

// Geometry of segment q of a chain (windows of C members from beg): members [begin, end), entry trailing zeros je and
// sum of X over [P_q, begin).  The head's segment starts at the head's end; the others right after their window's anchor.
template <typename Acc>
QB_HD void fx_segment_bounds(const Acc &acc, const Tables &tab, unsigned int beg, unsigned int end, unsigned int C, unsigned int q,
                             unsigned int n_win, unsigned int q_start, unsigned int head_end, int head_je, unsigned int &seg_begin,
                             unsigned int &seg_end, int &je, u128 &xsum_before) {
  const unsigned int P = beg + q * C;
  if (q == q_start) {
    u128 xs = 0;
    for (unsigned int p = P; p < head_end; p++) xs += tab.X[acc(p)];
    seg_begin = head_end;
    je = head_je;
    xsum_before = xs;
  } else {
    const Anchor a = fx_anchor(acc, P, end, tab);
    seg_begin = a.b;
    je = a.je < 0 ? 0 : a.je;
    xsum_before = ((u128)a.xsum_hi << 64) | a.xsum_lo;
  }
  seg_end = q + 1 < n_win ? fx_anchor(acc, beg + (q + 1) * C, end, tab).b : end;
}

}  // namespace fx
}  // namespace qb
