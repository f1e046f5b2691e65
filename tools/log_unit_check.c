// log_unit_check.c — host model of the device log_unit() (nlsolver_b200/csrc/pso_impl.cuh): only IEEE operations (fma, /,
// -, *) and integer bit manipulation, so the device result is bit-identical.  Measures the error against glibc log in ulps
// over unit-interval inputs of the draw tape's form.   gcc -O2 -ffp-contract=off tools/log_unit_check.c -lm && ./a.out
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
  Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
  Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
  Lg7 = 1.479819860511658591e-01;
static double log_unit(double x) {
  uint64_t b; memcpy(&b, &x, 8);
  int32_t hx = (int32_t)(b >> 32); uint32_t lx = (uint32_t)b;
  if (x == 0.0) return -INFINITY;
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  int i = (hx + 0x95f64) & 0x100000;
  hx |= (i ^ 0x3ff00000); k += (i >> 20);
  b = ((uint64_t)(uint32_t)hx << 32) | lx;
  double m; memcpy(&m, &b, 8);
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double dk = (double)k;
  double z = s * s, w = z * z;
  double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
  double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
  double R = t2 + t1;
  double hfsq = 0.5 * f * f;
  return fma(dk, ln2_hi, -((hfsq - fma(s, hfsq + R, dk * ln2_lo)) - f));
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
