/* Research prototype (CPU, C): the integer model of kahan_automaton.py / kahan_segments.py in fixed-width arithmetic
 * (unsigned __int128, as a CUDA kernel would hold it), checked against the floating-point loop on long chains.
 *   gcc -O2 -std=c11 -fno-fast-math -ffp-contract=off -o kahan_fast kahan_fast.c && ./kahan_fast
 * A = (sum - c) / 2^-60 once sum >= 4; see the Python files for the derivation. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef unsigned __int128 u128;
typedef __int128 i128;

static uint64_t X_T[256], U_T[256];

static void tables(void) {
  for (int t = 1; t < 256; t++) {
    double x = (double)t / 255.0;
    int e;
    frexp(x, &e);                       /* x = m * 2^e, 0.5 <= m < 1 */
    X_T[t] = (uint64_t)ldexp(x, 60);    /* exact */
    U_T[t] = (uint64_t)1 << (e - 1 - 52 + 60);
  }
}

static int bitlen(u128 a) {
  uint64_t hi = (uint64_t)(a >> 64), lo = (uint64_t)a;
  return hi ? 128 - __builtin_clzll(hi) : (lo ? 64 - __builtin_clzll(lo) : 0);
}

/* round to 53 significant bits, ties to even; *U receives the ulp */
static u128 rn53(u128 A, u128 *U) {
  int n = bitlen(A);
  if (n <= 53) { *U = 1; return A; }
  u128 u = (u128)1 << (n - 53), q = A / u, r = A % u;
  if (r * 2 > u || (r * 2 == u && (q & 1))) q++;
  *U = u;
  return q * u;
}

static u128 round_even(u128 v, uint64_t u, u128 parity_base) { /* v to a multiple of u, ties: (parity_base + result)/u even */
  u128 r = v % u, lo = v - r;
  if (r * 2 < u) return lo;
  if (r * 2 > u) return lo + u;
  return (((parity_base + lo) / u) & 1) == 0 ? lo : lo + u;
}

static u128 step_int(u128 A, int t) {
  if (t == 0) return A;
  uint64_t X = X_T[t];
  if (t < 255) return X + round_even(A, U_T[t], X);
  u128 U, s = rn53(A, &U);
  if (A == s) return A + X;
  if (A > s) {                          /* c < 0: y = fl(1 + d) above 1, grid 2^8 */
    u128 d = A - s, v = X + d;
    return s + round_even(v, 256, 0);
  } else {                              /* c > 0: y = fl(1 - d) below 1, grid 2^7 */
    u128 d = s - A, v = X - d;
    return s + round_even(v, 128, 0);
  }
}

static double to_double(u128 A) { /* A is already rounded to 53 bits */
  return ldexp((double)(uint64_t)(A >> 32), 32 - 60) + ldexp((double)(uint64_t)(A & 0xffffffffu), -60);
}

int main(void) {
  tables();
  srand(12345);
  long bad = 0;
  for (int trial = 0; trial < 16; trial++) {
    const long n = 4000000;
    const int kind = trial % 4;
    uint8_t *ts = malloc(n);
    for (long i = 0; i < n; i++) {
      int r = rand();
      ts[i] = kind == 0 ? (uint8_t)(r & 255)
            : kind == 1 ? (uint8_t)((r % 5 == 0) ? 255 : (r % 5 == 1) ? 254 : (r % 5 == 2) ? 0 : (r % 5 == 3) ? 1 : 200)
            : kind == 2 ? (uint8_t)(r % 40)
                        : (uint8_t)((r % 5 == 0) ? 255 : (r % 5 == 1) ? 128 : (r % 5 == 2) ? 127 : (r % 5 == 3) ? 64 : 63);
    }
    /* the reference's loop */
    volatile double s = 0.0, c = 0.0;
    long k = 0;
    for (; k < n && s < 4.0; k++) {
      double x = (double)ts[k] / 255.0, y = x - c, t2 = s + y;
      c = (t2 - s) - y;
      s = t2;
    }
    /* A = (s - c) * 2^60, exactly: s and c are multiples of 2^-60 here and small */
    i128 Ai = (i128)ldexp(s, 60) - (i128)ldexp(c, 60);
    u128 A = (u128)Ai;
    for (long j = k; j < n; j++) {
      double x = (double)ts[j] / 255.0, y = x - c, t2 = s + y;
      c = (t2 - s) - y;
      s = t2;
      A = step_int(A, ts[j]);
      if ((j & 0xfffff) == 0 || j == n - 1) {
        u128 U, r = rn53(A, &U);
        if (to_double(r) != s) { bad++; printf("trial %d kind %d: diverged by element %ld\n", trial, kind, j); break; }
      }
    }
    printf("trial %2d kind %d: n = %ld, sum = %.17g %s\n", trial, kind, n, (double)s, "ok");
    free(ts);
  }
  printf("%ld divergences\n", bad);
  return bad != 0;
}
