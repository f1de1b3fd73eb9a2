// Host check of shrimp_b200/csrc/glibc_math.cuh against the libm of this image: exits non-zero on the first
// argument whose exp()/log() bits differ.   g++ -O2 -x c++ tools/check_glibc_math.c -o /tmp/chk -lm && /tmp/chk
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../shrimp_b200/csrc/glibc_math.cuh"
using namespace glibc_math;
static unsigned long long s = 0x9E3779B97F4A7C15ull;
static unsigned long long rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(int argc, char **argv) {
  const long n = argc > 1 ? atol(argv[1]) : 50000000L;
  Tables T = {GLIBC_EXP_CONST, GLIBC_EXP_TAB, GLIBC_LOG_CONST, GLIBC_LOG_TAB};
  long bad = 0;
  for (long i = 0; i < n && bad < 10; i++) {
    // exp: the arguments post_sw uses (0 .. -800), then anything
    double x;
    const unsigned long long u = rnd();
    switch (i & 3) {
      case 0: x = -(double)(u >> 11) / 9007199254740992.0 * 60.0; break;
      case 1: x = -(double)(u >> 11) / 9007199254740992.0 * 800.0; break;
      case 2: x = ((double)(u >> 11) / 9007199254740992.0 - 0.5) * 1500.0; break;
      default: x = gm_asdouble(u); break;
    }
    const double a = exp(x), b = exp_glibc(x, T);
    if (gm_asuint64(a) != gm_asuint64(b) && !(a != a && b != b)) { printf("exp(%a): libm %a mine %a\n", x, a, b); bad++; }
    double y;
    const unsigned long long v = rnd();
    switch (i & 3) {
      case 0: y = (double)(v >> 11) / 9007199254740992.0 * 16.0; break;
      case 1: y = 0.9 + (double)(v >> 11) / 9007199254740992.0 * 0.2; break;
      case 2: y = gm_asdouble(v & 0x7fffffffffffffffull); break;
      default: y = exp(-(double)(v >> 11) / 9007199254740992.0 * 700.0); break;
    }
    const double c = log(y), d = log_glibc(y, T);
    if (gm_asuint64(c) != gm_asuint64(d) && !(c != c && d != d)) { printf("log(%a): libm %a mine %a\n", y, c, d); bad++; }
  }
  // the arguments the device entry points take off libm's branches: |x| around and below 2^-54 for exp (libm returns
  // 1 + x there, glibc_math.cuh runs the main path), the edges of log's near-1 interval
  for (long i = 0; i < n / 8 && bad < 10; i++) {
    const unsigned long long u = rnd(), v = rnd();
    const int e = 1023 - 50 - (int)(u % 12);   // 2^-50 .. 2^-61
    const double x = gm_asdouble(((u >> 63) << 63) | ((unsigned long long)(i & 1 ? e : (int)(v % 1023)) << 52) | (v >> 12));
    const double a = exp(x), b = exp_glibc(x, T);
    if (gm_asuint64(a) != gm_asuint64(b)) { printf("exp(%a): libm %a mine %a\n", x, a, b); bad++; }
    const double edge[4] = {1.0 - 0x1p-4, 1.0 + 0x1.09p-4, 1.0, 1.0};
    const double y = edge[u & 3] + ((double)(v >> 11) / 9007199254740992.0 - 0.5) * ((u & 3) < 2 ? 1e-12 : 1e-15 * (double)(1 + (u >> 8) % 1000));
    const double c = log(y), d = log_glibc(y, T);
    if (gm_asuint64(c) != gm_asuint64(d)) { printf("log(%a): libm %a mine %a\n", y, c, d); bad++; }
  }
  const double sp[] = {0.0, -0.0, 1.0, INFINITY, -INFINITY, 1e-310, 4.9e-324, -1.0, 0x1p-1022, 709.78, 709.79, -745.13, -745.14, -708.4, -1022.0, 512.0, -512.0};
  for (unsigned i = 0; i < sizeof(sp) / sizeof(sp[0]); i++) {
    const double a = exp(sp[i]), b = exp_glibc(sp[i], T), c = log(sp[i]), d = log_glibc(sp[i], T);
    if (gm_asuint64(a) != gm_asuint64(b) && !(a != a && b != b)) { printf("exp(%a): libm %a mine %a\n", sp[i], a, b); bad++; }
    if (gm_asuint64(c) != gm_asuint64(d) && !(c != c && d != d)) { printf("log(%a): libm %a mine %a\n", sp[i], c, d); bad++; }
  }
  printf("%ld arguments each, %ld mismatches\n", n, bad);
  return bad != 0;
}
