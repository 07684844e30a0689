// CPU check of quant_fast_bits / quant4_pack (csrc/ot_rowmath.cuh): for random and adversarial (y, s) the byte taken from the bit
// pattern of q1 + 1.5*2^23 -- or the exact fallback when the element is flagged -- equals (int8) rint(y / s), the reference
// RowQuant (quant_linear.py:31-43).  Build: gcc -O2 -ffp-contract=off -o tools/bin/check_quant_bits tools/check_quant_bits.c -lm
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint64_t rng = 88172645463325252ull;
static uint32_t xr(void) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (uint32_t)(rng >> 16); }

static int check(float y, float s, long* flagged) {
  const float r = 1.0f / s;                       // __frcp_rn
  const float q0 = y * r;
  const float rem = fmaf(-q0, s, y);
  const float q1 = fmaf(rem, r, q0);
  const float t = q1 + 12582912.0f;
  const float n = t - 12582912.0f;
  const int slow = (fabsf(fabsf(q1 - n) - 0.5f) < 1.52587890625e-05f) || !(fabsf(q1) < 1024.0f);
  const float exact = rintf(y / s);
  if (slow) { ++*flagged; return 1; }             // the kernel then evaluates rintf(y / s) itself
  const int8_t fast = (int8_t)(bits(t) & 0xFFu);
  long e = (long)exact;
  if (e != (long)n) { printf("n mismatch y=%a s=%a n=%g exact=%g\n", y, s, n, exact); return 0; }
  if (e >= -128 && e <= 127 && fast != (int8_t)e) { printf("byte mismatch y=%a s=%a\n", y, s); return 0; }
  if (((uint32_t)(int32_t)e & 0xFFu) != (bits(t) & 0xFFu)) { printf("low byte mismatch y=%a s=%a\n", y, s); return 0; }
  return 1;
}

int main(int argc, char** argv) {
  const long N = argc > 1 ? atol(argv[1]) : 400000000L;
  long flagged = 0, done = 0;
  for (long i = 0; i < N; ++i) {
    // scale: amax / 127 with amax log-uniform in [1e-5, 1e6]; y: |y| <= amax (the RowQuant contract), biased towards half-integers
    const float amax = expf(logf(1e-5f) + (xr() / 4294967296.0f) * (logf(1e6f) - logf(1e-5f)));
    const float s = fmaxf(amax, 1e-5f) / 127.0f;
    float y;
    const uint32_t k = xr();
    if ((k & 3) == 0) {                            // near a rounding boundary: (m + 0.5) * s perturbed by a few ulps
      const int m = (int)(xr() % 255) - 127;
      y = ((float)m + 0.5f) * s;
      y = from_bits(bits(y) + (xr() % 9) - 4);
    } else {
      y = ((xr() / 2147483648.0f) - 1.0f) * amax;
    }
    if (fabsf(y) > amax) y = copysignf(amax, y);
    if (!check(y, s, &flagged)) return 1;
    ++done;
  }
  printf("%ld cases, %ld flagged for the exact path (%.4f %%), 0 mismatches\n", done, flagged, 100.0 * flagged / done);
  return 0;
}
