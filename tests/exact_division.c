/* CPU restatement of the division sequences of the strict CUDA kernels
 * (xline_b200/csrc/track_impl.cuh: div_small_int, div_known_recip), checked against the IEEE
 * division the reference performs (xline/elements.py:130-134 `/ ii`, :143-144 `/ length`,
 * :436 `/ (a*a)`).  Test infrastructure: built and run by tests/test_exact_division.py.
 *
 *   usage: exact_division <mode> <samples per divisor> <seed> [divisor ...]
 *   mode 0: integer divisors 1..255 (all of them when none are listed), three-instruction form
 *   mode 1: arbitrary divisors (random ones when none are listed), five-instruction form
 * Prints "<samples> <mismatches>".
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

static inline double div_small_int(double a, double b, double y) {
  const double q0 = a * y;
  const double e = fma(q0, b, -a);
  return fma(-e, y, q0);
}
static inline double div_known_recip(double a, double b, double y) {
  const double q0 = a * y;
  const double e0 = fma(q0, b, -a);
  const double q1 = fma(-e0, y, q0);
  const double e1 = fma(q1, b, -a);
  return fma(-e1, y, q1);
}

static long check(int mode, double b, long n, long *bad) {
  const double y = 1.0 / b;
  long done = 0;
  for (long k = 0; k < n; ++k) {
    uint64_t m = rnd() & 0x800fffffffffffffULL;
    /* exponents 2^-300 .. 2^300; the kernels guard |a| < 2^-959 and non-finite values */
    uint64_t e = 1023 + (int)(rnd() % 601) - 300;
    double a = from_bits(m | (e << 52));
    if ((k & 3) == 0) { /* next to a rounding midpoint of the quotient */
      const double q = fabs(a);
      const double half_ulp = from_bits(bits(q) & 0x7ff0000000000000ULL) * 1.1102230246251565e-16;
      a = fma(q, b, half_ulp * b);
      if (k & 4) a = -a;
      if (k & 8) a = nextafter(a, (k & 16) ? 0.0 : INFINITY);
    }
    const double want = a / b;
    const double got = mode == 0 ? div_small_int(a, b, y) : div_known_recip(a, b, y);
    if (bits(want) != bits(got)) {
      if (*bad < 5) fprintf(stderr, "mismatch: %a / %a = %a, sequence gives %a\n", a, b, want, got);
      ++*bad;
    }
    ++done;
  }
  /* signed zeros keep their sign */
  const double z[2] = {0.0, -0.0};
  for (int i = 0; i < 2; ++i) {
    const double got = mode == 0 ? div_small_int(z[i], b, y) : div_known_recip(z[i], b, y);
    if (bits(got) != bits(z[i] / b)) ++*bad;
    ++done;
  }
  return done;
}

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  const int mode = atoi(argv[1]);
  const long n = atol(argv[2]);
  s = strtoull(argv[3], 0, 10) * 2654435761ULL + 88172645463325252ULL;
  long bad = 0, done = 0;
  if (argc > 4) {
    for (int i = 4; i < argc; ++i) done += check(mode, strtod(argv[i], 0), n, &bad);
  } else if (mode == 0) {
    for (int ii = 1; ii <= 255; ++ii) done += check(0, (double)ii, n, &bad);
  } else {
    for (int i = 0; i < 256; ++i) {
      uint64_t m = rnd() & 0x000fffffffffffffULL;
      uint64_t e = 1023 + (int)(rnd() % 81) - 40;
      if ((i & 7) == 0) m = 0x000fffffffffffffULL - (rnd() & 3); /* significands next to all ones */
      done += check(1, from_bits(m | (e << 52)), n, &bad);
    }
  }
  printf("%ld %ld\n", done, bad);
  return 0;
}
