// CPU check of quant_b200/csrc/qb200_exact_fast.cuh: the per-thread bodies the exact-sum kernels run (window sums,
// head, anchors, speculative class runs, chaining) are executed here in plain loops over chains of lattice values and
// compared, bit for bit, with the reference's compensated loop (Solution::sumInArea,
// /root/reference/src/Quantizer.cpp:59-70) run in IEEE double.
//   g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -I quant_b200/csrc tests/cpp/exact_fast_test.cpp -o /tmp/exact_fast_test
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "qb200_exact_fast.cuh"

using namespace qb::fx;

static Tables g_tab;

struct Acc {
  const uint8_t *ts;
  int operator()(unsigned int p) const { return ts[p]; }
};

static void kahan_fp(const uint8_t *ts, size_t n, double &sum, double &c) {
  volatile double s = sum, cc = c;
  for (size_t i = 0; i < n; i++) {
    volatile double x = (double)ts[i] / 255.0;
    volatile double y = x - cc;
    volatile double t = s + y;
    volatile double d = t - s;
    cc = d - y;
    s = t;
  }
  sum = s;
  c = cc;
}

struct Stats {
  long segments = 0, reruns = 0, sequential = 0, classes = 0, batches = 0, refined = 0;
};

// The kernels' pipeline on one chain [0, n) with window length C, from the incoming pair (sum, c).
static void fast_chain(const uint8_t *ts, unsigned int n, unsigned int C, double &sum, double &c, Stats &st) {
  Acc acc{ts};
  const unsigned int n_win = (n + C - 1) / C;
  // pre-pass: window sums and need maxima; exclusive prefix
  std::vector<u128> sumX(n_win), pref(n_win + 1);
  std::vector<int> need(n_win + 1, -1);
  for (unsigned int q = 0; q < n_win; q++) fx_window(acc, g_tab, q * C, std::min(n, (q + 1) * C), sumX[q], need[q]);
  pref[0] = 0;
  for (unsigned int q = 0; q < n_win; q++) pref[q + 1] = pref[q] + sumX[q];
  // head
  const HeadOut h = fx_head(acc, g_tab, 0, n, C, sum, c);
  if (h.done) {
    sum = h.sum;
    c = h.c;
    return;
  }
  // speculative runs, every window independently
  std::vector<SegRecord> rec(n_win);
  for (unsigned int q = h.q_start; q < n_win; q++) {
    SegRecord &r = rec[q];
    int je;
    u128 before;
    fx_segment_bounds(acc, g_tab, 0, n, C, q, n_win, h.q_start, h.pos_end, h.je, r.begin, r.end, je, before);
    r.je = (signed char)je;
    r.top = (signed char)std::max(need[q], need[q + 1]);
    r.Eb = (u128)(h.B + (i128)pref[q] + (i128)before);
    r.ncls = 0;
    r.w_base = 0;
    r.flags = 0;
    if (r.begin < r.end) {
      fx_run_segment_multi(acc, g_tab, r);
      SegRecord chk = r;   // the one-pass evaluation must agree with the class-by-class one
      fx_run_segment(acc, g_tab, chk);
      if (chk.ncls != r.ncls || std::memcmp(chk.dcorr, r.dcorr, sizeof(short) * (chk.ncls > 0 ? chk.ncls : 0)) ||
          std::memcmp(chk.margin, r.margin, sizeof(unsigned) * (chk.ncls > 0 ? chk.ncls : 0))) {
        std::printf("multi-class run differs from the class-by-class run (window %u)\n", q);
        std::exit(1);
      }
    }
  }
  // chaining: the head's state is exact, so the first segment is entered with W = 0.  Batches of 32 segments, as
  // the chain kernel walks them: 4-state maps and their prefix composition where the batch allows it.  Two passes:
  // the first only marks the segments whose summary does not cover W (and estimates W behind them), the marked ones
  // are re-run around that estimate, the second pass is exact.
  long long W = 0;
  for (int pass = 0; pass < 2; pass++) {
    W = 0;
    auto sequential = [&](unsigned int q) {
      if (rec[q].begin >= rec[q].end) return;
      if (!fx_apply(rec[q], W)) {
        if (pass == 0) {
          fx_mark_and_estimate(rec[q], W);
        } else {
          fx_rerun(acc, g_tab, rec[q], W);
          if (rec[q].ncls == 0) st.sequential++; else st.reruns++;
        }
      }
    };
    for (unsigned int q0 = h.q_start; q0 < n_win; q0 += 32) {
      const unsigned int nb = std::min(32u, n_win - q0);
      int je_min = 99, top_max = -1, live = 0;
      for (unsigned int j = 0; j < nb; j++) {
        const SegRecord &r = rec[q0 + j];
        if (r.begin >= r.end) continue;
        if (pass == 1) {
          st.segments++;
          st.classes += r.ncls;
        }
        live++;
        if (r.ncls == 0) continue;
        je_min = std::min(je_min, (int)r.je);
        top_max = std::max(top_max, (int)r.top);
      }
      int jb = 0;
      if (!live) continue;
      if (je_min == 99 || !fx_batch_composable(je_min, 0, top_max, jb) || W > kFxWLimit / 2 || W < -kFxWLimit / 2) {
        for (unsigned int j = 0; j < nb; j++) sequential(q0 + j);
        continue;
      }
      if (pass == 1) st.batches++;
      unsigned int j = 0;
      while (j < nb) {
        // prefix maps of records j .. nb-1 (the warp computes them with a parallel prefix)
        const SegRecord &first = rec[q0 + j];
        const int s0 = (int)(((((long long)((unsigned long long)first.Eb & 0xffffu)) + (W - first.w_base)) >> jb) & 3);
        Map4 pre = fx_map_identity();
        unsigned int applied = j;
        long long W_after = W;
        for (unsigned int i = j; i < nb; i++) {
          pre = fx_compose(pre, fx_map_of(rec[q0 + i], jb));
          if (!(pre.lo[s0] <= pre.hi[s0] && W >= pre.lo[s0] && W <= pre.hi[s0])) break;
          applied = i + 1;
          W_after = W + pre.dW[s0];
        }
        W = W_after;
        j = applied;
        if (j < nb) {   // record j does not cover W (or is sequential): mark / exact, then go on behind it
          sequential(q0 + j);
          j++;
        }
      }
    }
    if (pass == 0) {  // refinement round: the marked segments again, around the first pass's estimate of W
      for (unsigned int q = h.q_start; q < n_win; q++)
        if (rec[q].flags & kFxRefine) {
          fx_run_segment_multi(acc, g_tab, rec[q]);
          rec[q].flags = 0;
          st.refined++;
        }
    }
  }
  const u128 A = (u128)(h.B + (i128)pref[n_win] + W);
  fx_state_to_pair(A, sum, c);
}

static uint32_t rng_state = 12345;
static uint32_t rnd() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 17;
  rng_state ^= rng_state << 5;
  return rng_state;
}

static void fill(std::vector<uint8_t> &ts, int kind) {
  const size_t n = ts.size();
  for (size_t i = 0; i < n; i++) {
    const uint32_t r = rnd();
    uint8_t t;
    switch (kind) {
      case 0: t = (uint8_t)(r & 255); break;                                                  // noise
      case 1: t = (uint8_t)((r % 16 == 0) ? 255 : (r >> 8) & 255); break;                    // 6 % of t = 255
      case 2: t = (uint8_t)((r % 50 == 0) ? 128 + (r >> 8) % 127 : (r >> 8) % 40); break;    // small values, rare large
      case 3: { const uint8_t v[5] = {255, 128, 127, 64, 63}; t = v[r % 5]; break; }
      case 4: t = (uint8_t)((r >> 8) % 64); break;                                             // bright image bytes: t < 64 only
      case 5: t = 255; break;                                                                  // flat mid-grey
      case 6: t = (uint8_t)((r % 3 == 0) ? 255 : 1); break;                                    // 1s and 255s: many classes
      case 7: t = (uint8_t)((r % 7 == 0) ? (r >> 8) % 8 : 0); break;                           // mostly zeros
      case 8: t = (uint8_t)(((i / 5000) & 1) ? (r >> 8) % 32 : 128 + (r >> 8) % 128); break;  // alternating regions
      case 9: t = (uint8_t)((r % 4 == 0) ? 255 : 254); break;
      case 10: t = (uint8_t)(100 + (r >> 8) % 3); break;                                       // near-flat patch
      case 11: t = (uint8_t)((r % 97 == 0) ? 255 : ((r >> 8) % 16)); break;                   // dark with rare 255s
      case 12: t = 0; break;
      default: t = (uint8_t)((r % 2) ? 3 : 200); break;
    }
    ts[i] = t;
  }
}

int main(int argc, char **argv) {
  fx_fill_tables(g_tab);
  // tables as the research prototypes define them
  for (int t = 1; t < 256; t++) {
    const double x = (double)t / 255.0;
    if (std::ldexp(x, 60) != (double)g_tab.X[t]) { std::printf("table mismatch at %d\n", t); return 1; }
  }
  const long scale = argc > 1 ? std::atol(argv[1]) : 1;
  long bad = 0, chains = 0;
  Stats total;
  const unsigned int Cs[4] = {64, 128, 512, 1024};
  for (int kind = 0; kind < 14; kind++) {
    for (int rep = 0; rep < 6; rep++) {
      const size_t n = (size_t)(scale * (rep == 0 ? 200000 : (1000 + rnd() % 40000)));
      const unsigned int C = Cs[(kind + rep) % 4];
      std::vector<uint8_t> ts(n);
      fill(ts, kind);
      double s_ref = 0, c_ref = 0;
      kahan_fp(ts.data(), n, s_ref, c_ref);
      double s = 0, c = 0;
      Stats st;
      fast_chain(ts.data(), (unsigned int)n, C, s, c, st);
      chains++;
      if (std::memcmp(&s, &s_ref, 8) || std::memcmp(&c, &c_ref, 8)) {
        bad++;
        std::printf("kind %d rep %d n %zu C %u: MISMATCH sum %.17g vs %.17g, c %.17g vs %.17g\n", kind, rep, n, C, s, s_ref, c, c_ref);
      }
      // the same chain cut at a random point and continued from the first part's pair (two ranks)
      const size_t cut = 1 + rnd() % (n - 1);
      double s2 = 0, c2 = 0;
      Stats st2;
      fast_chain(ts.data(), (unsigned int)cut, C, s2, c2, st2);
      fast_chain(ts.data() + cut, (unsigned int)(n - cut), C, s2, c2, st2);
      chains++;
      if (std::memcmp(&s2, &s_ref, 8) || std::memcmp(&c2, &c_ref, 8)) {
        bad++;
        std::printf("kind %d rep %d n %zu C %u cut %zu: MISMATCH (continued chain) sum %.17g vs %.17g, c %.17g vs %.17g\n", kind, rep, n,
                    C, cut, s2, s_ref, c2, c_ref);
      }
      total.segments += st.segments; total.reruns += st.reruns; total.sequential += st.sequential; total.classes += st.classes; total.batches += st.batches; total.refined += st.refined;
      if (rep == 0)
        std::printf("kind %2d n %7zu C %4u: %6ld segments, %5.2f classes each, %4ld margin re-runs, %4ld sequential   sum %.17g\n", kind, n, C,
                    st.segments, st.segments ? (double)st.classes / st.segments : 0.0, st.reruns, st.sequential, s_ref);
    }
  }
  std::printf("%ld chains, %ld mismatches; %ld segments (%ld batches chained through composed maps), %ld refined around the first pass's estimate, %ld re-run exactly for their margin, %ld run sequentially\n",
              chains, bad, total.segments, total.batches, total.refined, total.reruns, total.sequential);
  return bad != 0;
}
