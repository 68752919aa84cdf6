/* TEST INFRASTRUCTURE. Exhaustive-style check of the reciprocal division used by the CUDA kernels
 * (hommexx_b200/csrc/hxx.cuh div_rcp): q = RN(x r), e = fma(-d, q, x), result = fma(e, r, q) with r = RN(1/d)
 * must equal the IEEE quotient RN(x / d) for every operand inside the kernels' exponent window. Operands are
 * drawn to hit the hard cases: random mantissas, mantissas next to a power of two, divisors with long runs
 * of ones, quotients next to a representable value (x = q0 d +- a few ulps) and — the cases that decide correct
 * rounding — quotients next to a TIE, the midpoint m = q0 + ulp(q0)/2 of two neighbouring doubles: x = RN(m d)
 * formed in binary128 and perturbed by a few ulps, so that x / d lies within ~2^-53 ulp of the midpoint.
 *   div_rcp_check <millions of pairs> <seed>   ->  prints "<pairs> <mismatches>"                            */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[2];
static inline uint64_t rnd(void) { /* xorshift128+ */
  uint64_t a = s[0], b = s[1];
  s[0] = b;
  a ^= a << 23;
  s[1] = a ^ b ^ (a >> 17) ^ (b >> 26);
  return s[1] + b;
}
static inline double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t to_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

static double mant(int kind) { /* a double in [1, 2) */
  uint64_t m = rnd() & 0xfffffffffffffULL;
  switch (kind & 7) {
    case 1: m &= 0xffULL; break;                        /* just above 1 */
    case 2: m |= 0xfffffffffff00ULL; break;             /* just below 2 */
    case 3: m = (m & 0xfff) | ((rnd() & 0xfff) << 40); break; /* sparse */
    case 4: m |= 0x000ffffffff000ULL; break;            /* long run of ones */
    case 5: m &= 0xffffff0000000ULL; break;             /* few significant bits */
    default: break;
  }
  return from_bits(0x3ff0000000000000ULL | m);
}

static inline double div_rcp(double x, double d, double r) {
  const double q = x * r;
  const double e = fma(-d, q, x);
  return fma(e, r, q);
}

int main(int argc, char** argv) {
  const long n = (argc > 1 ? atol(argv[1]) : 10) * 1000000L;
  s[0] = 0x9e3779b97f4a7c15ULL ^ (uint64_t)(argc > 2 ? atol(argv[2]) : 1);
  s[1] = 0xbf58476d1ce4e5b9ULL;
  long bad = 0;
  for (long i = 0; i < n; ++i) {
    const int kind = (int)(rnd() & 63);
    double d = ldexp(mant(kind), (int)(rnd() % 80) - 40);           /* divisors: 2^-40 .. 2^40 */
    double x;
    if ((kind & 24) == 24) { /* a quotient next to a tie: x ~ (q0 + ulp/2) d */
      const double q0 = ldexp(mant(kind >> 3), (int)(rnd() % 60) - 30);
      const __float128 m = (__float128)q0 + (__float128)(nextafter(q0, INFINITY) - q0) / 2;
      x = (double)(m * (__float128)d);
      x = from_bits(to_bits(x) + (rnd() % 5) - 2);
    } else if (kind & 8) { /* a quotient next to a representable value: x = q0 * d perturbed by a few ulps */
      const double q0 = ldexp(mant(kind >> 3), (int)(rnd() % 60) - 30);
      x = q0 * d;
      x = from_bits(to_bits(x) + (rnd() % 5) - 2);
    } else {
      x = ldexp(mant(kind >> 3), (int)(rnd() % 400) - 200);        /* the kernels' window is 2^-600 .. 2^600 */
    }
    if (rnd() & 1) x = -x;
    if ((rnd() & 7) == 0) d = -d;
    const double r = 1.0 / d;
    if (div_rcp(x, d, r) != x / d) {
      if (bad < 5) fprintf(stderr, "mismatch: x=%a d=%a  rcp=%a ieee=%a\n", x, d, div_rcp(x, d, r), x / d);
      ++bad;
    }
  }
  printf("%ld %ld\n", n, bad);
  return 0;
}
