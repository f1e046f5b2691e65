// log_unit_check.c — host model of the device log_unit() (nlsolver_b200/csrc/pso_impl.cuh): only IEEE operations (fma, +,
// *), integer bit manipulation and the shared table nlsolver_b200/csrc/log_table.h, so the device result is bit-identical.  Measures the error against glibc log in ulps
// over unit-interval inputs of the draw tape's form.   gcc -O2 -ffp-contract=off tools/log_unit_check.c -lm && ./a.out
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "../nlsolver_b200/csrc/log_table.h"
static const double kLogTab[256] = {NLS_LOG_TABLE_ROWS};
static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
  P2 = -5.0e-01, P3 = 3.3333333333333331483e-01, P4 = -2.5e-01, P5 = 2.0000000000000001110e-01,
  P6 = -1.6666666666666665741e-01, P7 = 1.4285714285714284921e-01;
// log(x), x = raw * 2^-64 in [0, 1]: x = 2^k m with m in [sqrt(1/2), sqrt(2)) (the fdlibm re-biasing), the top seven bits
// of the re-biased mantissa select a cell {rc, -log(rc)}, r = m rc - 1 (one fma, |r| <= 2^-8), log m = -log(rc) + log1p(r)
// with the degree-7 Taylor polynomial (the cell around m = 1 returns log1p(r) itself, so the
// truncation r^8 / 8 has to be small RELATIVE to r), and k ln2 added with the compensated split.
static double log_unit(double x) {
  uint64_t b; memcpy(&b, &x, 8);
  int32_t hx = (int32_t)(b >> 32); uint32_t lx = (uint32_t)b;
  if (x == 0.0) return -INFINITY;
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int a = hx + 0x95f64;
  const int i = a & 0x100000;
  const int j = (a >> 13) & 0x7f;
  hx |= (i ^ 0x3ff00000); k += (i >> 20);
  b = ((uint64_t)(uint32_t)hx << 32) | lx;
  double m; memcpy(&m, &b, 8);
  const double r = fma(m, kLogTab[2 * j], -1.0);
  const double dk = (double)k;
  double q = fma(r, P7, P6);
  q = fma(r, q, P5);
  q = fma(r, q, P4);
  q = fma(r, q, P3);
  q = fma(r, q, P2);
  const double z = fma(r * r, q, r);
  return fma(dk, ln2_hi, kLogTab[2 * j + 1]) + fma(dk, ln2_lo, z);
}
static uint64_t mix64(uint64_t z){ z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31);}
int main(int argc, char **argv) {
  double maxulp = 0; uint64_t worst = 0; long n = 0, exact = 0;
  const uint64_t count = argc > 1 ? strtoull(argv[1], 0, 10) : 40000000ull;
  for (uint64_t c = 1; c <= count; c++) {
    uint64_t raw = mix64(c * 0x9E3779B97F4A7C15ull);
    if (c % 7 == 0) raw >>= (c % 60);          // small values too
    if (raw == 0) continue;
    double x = (double)raw * 0x1p-64;
    double a = log_unit(x), r = log(x);
    long double t = logl((long double)x);
    double ulp = fabs(nextafter(r, INFINITY) - r);
    double e = fabsl((long double)a - t) / ulp;
    if (e > maxulp) { maxulp = e; worst = raw; }
    n++; exact += (a == r);
  }
  printf("n=%ld  max error %.4f ulp (raw=%llu)  bit-equal to glibc in %.4f%%\n", n, maxulp, (unsigned long long)worst, 100.0*exact/n);
  double xs[] = {1.0, 0x1p-64, 0.5, 0.70710678118654752, 0.70710678118654757, 1.0 - 0x1p-53, 0x1p-1 + 0x1p-54};
  for (unsigned i = 0; i < sizeof(xs)/sizeof(xs[0]); i++) printf("x=%.17g  ours=%.17g  glibc=%.17g\n", xs[i], log_unit(xs[i]), log(xs[i]));
  return 0;
}
