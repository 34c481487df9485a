/* Research prototype (CPU, C): the segment scheme of kahan_segments.py in the decomposition a GPU would use, on one
 * chain: (1) anchors, (2) exact 128-bit segment sums and their prefix E, (3) four speculative runs per segment with
 * the margin of their t = 255 decisions - every segment independently, from E alone, (4) one sequential chaining
 * loop that re-runs a segment only when the true start is farther from the speculative one than the margin.
 *   gcc -O2 -std=gnu11 -fno-fast-math -ffp-contract=off -o kahan_segments kahan_segments.c -lm && ./kahan_segments
 * Reports exactness against the floating-point loop and how many segments had to be re-run. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef __int128 i128;

static uint64_t X_T[256], U_T[256];
static void tables(void) {
  for (int t = 1; t < 256; t++) {
    double x = (double)t / 255.0;
    int e;
    frexp(x, &e);
    X_T[t] = (uint64_t)ldexp(x, 60);
    U_T[t] = (uint64_t)1 << (e - 1 - 52 + 60);
  }
}
static int bitlen(u128 a) {
  uint64_t hi = (uint64_t)(a >> 64), lo = (uint64_t)a;
  return hi ? 128 - __builtin_clzll(hi) : (lo ? 64 - __builtin_clzll(lo) : 0);
}
static u128 rn53(u128 A, u128 *U) {
  int n = bitlen(A);
  if (n <= 53) { *U = 1; return A; }
  u128 u = (u128)1 << (n - 53), q = A / u, r = A % u;
  if (r * 2 > u || (r * 2 == u && (q & 1))) q++;
  *U = u;
  return q * u;
}
static u128 round_even(u128 v, uint64_t u, u128 base) {
  u128 r = v % u, lo = v - r;
  if (r * 2 < u) return lo;
  if (r * 2 > u) return lo + u;
  return (((base + lo) / u) & 1) == 0 ? lo : lo + u;
}
static u128 step_int(u128 A, int t) {
  if (t == 0) return A;
  uint64_t X = X_T[t];
  if (t < 255) return X + round_even(A, U_T[t], X);
  u128 U, s = rn53(A, &U);
  if (A == s) return A + X;
  if (A > s) return s + round_even(X + (A - s), 256, 0);
  return s + round_even(X - (s - A), 128, 0);
}
/* distance from A to the nearest point where the t = 255 branch (sign of A - RN53(A), or ulp(A)) changes */
static u128 decision_margin(u128 A) {
  int n = bitlen(A);
  u128 U = n > 53 ? (u128)1 << (n - 53) : 1, d = 0;
  if (U > 1) {
    u128 r = A % U, h = U / 2, a = r > h ? r - h : h - r, b = U - r;
    d = r < a ? r : a;
    if (b < d) d = b;
  }
  u128 lo = A - ((u128)1 << (n - 1)), hi = ((u128)1 << n) - A;
  if (lo < d) d = lo;
  if (hi < d) d = hi;
  return d;
}
static double to_double(u128 A) {
  return ldexp((double)(uint64_t)(A >> 32), 32 - 60) + ldexp((double)(uint64_t)(A & 0xffffffffu), -60);
}

typedef struct { long begin, end; u128 sumX; i128 delta[4]; u128 start[4], margin[4]; int has255; } Segment;

int main(void) {
  tables();
  srand(777);
  long bad = 0, reruns_total = 0, segs_total = 0;
  for (int trial = 0; trial < 12; trial++) {
    const long n = 3000000, spacing = 512;
    const int kind = trial % 4;
    uint8_t *ts = malloc(n);
    for (long i = 0; i < n; i++) {
      int r = rand();
      ts[i] = kind == 0 ? (uint8_t)(r & 255)
            : kind == 1 ? (uint8_t)((r % 16 == 0) ? 255 : (r >> 8) & 255)          /* 6 % saturated pixels */
            : kind == 2 ? (uint8_t)((r % 50 == 0) ? 128 + (r >> 8) % 127 : (r >> 8) % 40) /* dark, rare anchors */
                        : (uint8_t)((r % 5 == 0) ? 255 : (r % 5 == 1) ? 128 : (r % 5 == 2) ? 127 : (r % 5 == 3) ? 64 : 63);
    }
    /* reference: the floating-point loop over everything */
    volatile double s = 0.0, c = 0.0;
    long k = 0;
    double s_at_k = 0, c_at_k = 0;
    for (long i = 0; i < n; i++) {
      if (k == 0 && s >= 4.0) { k = i; s_at_k = s; c_at_k = c; }
      double x = (double)ts[i] / 255.0, y = x - c, t2 = s + y;
      c = (t2 - s) - y;
      s = t2;
    }
    u128 A = (u128)((i128)ldexp(s_at_k, 60) - (i128)ldexp(c_at_k, 60));
    /* head: sequentially up to and including the first anchor */
    long i = k;
    while (i < n && !(ts[i] >= 128 && ts[i] <= 254)) A = step_int(A, ts[i++]);
    if (i < n) A = step_int(A, ts[i++]);
    /* (1) segments (begin, end] ending at anchors >= spacing apart; the last one runs to the end of the chain */
    Segment *seg = malloc(sizeof(Segment) * (size_t)(n / spacing + 2));
    long ns = 0, b = i, last = i;
    for (long j = i; j < n; j++)
      if ((ts[j] >= 128 && ts[j] <= 254 && j + 1 - last >= spacing) || j == n - 1) {
        seg[ns].begin = b; seg[ns].end = j + 1; ns++;
        b = j + 1; last = j + 1;
      }
    /* (2) exact sums, then their prefix E (the trajectory without rounding corrections) */
    for (long q = 0; q < ns; q++) {
      u128 sx = 0; int h = 0;
      for (long j = seg[q].begin; j < seg[q].end; j++) { sx += X_T[ts[j]]; h |= ts[j] == 255; }
      seg[q].sumX = sx; seg[q].has255 = h;
    }
    /* (3) every segment, independently, for the residues 0, 128, 256, 384 of its start */
    u128 E = A;
    for (long q = 0; q < ns; q++) {
      for (int r = 0; r < 4; r++) {
        i128 corr = (i128)((128u * r + 512u - (unsigned)(E % 512)) % 512u);
        if (corr > 256) corr -= 512;
        u128 a = (u128)((i128)E + corr), a0 = a, mg = ~(u128)0;
        for (long j = seg[q].begin; j < seg[q].end; j++) {
          if (ts[j] == 255) { u128 d = decision_margin(a); if (d < mg) mg = d; }
          a = step_int(a, ts[j]);
        }
        seg[q].start[r] = a0; seg[q].delta[r] = (i128)a - (i128)a0; seg[q].margin[r] = mg;
      }
      E += seg[q].sumX;
    }
    /* (4) chaining */
    long reruns = 0;
    for (long q = 0; q < ns; q++) {
      int r = (int)((A % 512) / 128);
      if (A % 128 != 0) { printf("start of a segment is not a multiple of 128\n"); bad++; break; }
      i128 shift = (i128)A - (i128)seg[q].start[r];
      u128 mag = shift < 0 ? (u128)(-shift) : (u128)shift;
      if (!seg[q].has255 || mag < seg[q].margin[r]) {
        A = (u128)((i128)A + seg[q].delta[r]);
      } else {
        for (long j = seg[q].begin; j < seg[q].end; j++) A = step_int(A, ts[j]);
        reruns++;
      }
    }
    u128 U, rr = rn53(A, &U);
    const int ok = to_double(rr) == s;
    bad += !ok;
    reruns_total += reruns; segs_total += ns;
    printf("trial %2d kind %d: %ld segments, %ld re-run, sum %.17g %s\n", trial, kind, ns, reruns, (double)s, ok ? "exact" : "MISMATCH");
    free(seg); free(ts);
  }
  printf("%ld mismatches; %ld of %ld segments re-run\n", bad, reruns_total, segs_total);
  return bad != 0;
}
