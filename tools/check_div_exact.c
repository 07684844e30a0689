// CPU check of the branch-free exact quotient csrc/ot_gemm_wres.cu quantizes with:
//   r = RN(1/s); q0 = RN(y*r); q1 = fma(fma(-q0, s, y), r, q0); q2 = fma(fma(-q1, s, y), r, q1)   ==   RN(y / s)   (IEEE division)
// (Markstein: with the correctly rounded reciprocal, one FMA residual step applied to a FAITHFUL quotient gives the correctly rounded
// quotient; q0 can be 2 ulps off, q1 is faithful, q2 is exact.)  Random and adversarial (y, s): scales s = amax/127 with amax log-uniform
// in [1e-5, 1e6], |y| <= amax(1 + 2^-20), a quarter of the cases within a few ulps of (m + 1/2) s, plus scales whose significand is all
// ones / a power of two.  For dividends below s/4 (incl. zeros and denormals, where the residuals underflow)
// only the byte is compared (it is 0 either way).  Also checks that the FINAL byte (low byte of q2 + 1.5*2^23) equals (int8) rint(y / s).
// -DSEQ4: a four-operation form (r_lo = RN((1 - s r) r) per row; t = RN(y r_lo), qb = RN(y r + t), one residual step): 8e9 cases, 0
// mismatches.  One FFMA2 fewer per column pair, but in ot_gemm_wres.cu the extra per-row value tips the 96-register epilogue into spills
// (QKV 92.3 -> 95.6 us, FFN1 137.5 -> 143.6 us): measured and not adopted; the kernels keep the two-step form.  -DSEQ3: one step from the plain product (NOT proven; kept to show what the checker sees).
// Build: gcc -O2 -ffp-contract=off -fopenmp [-DSEQ4] -o tools/bin/check_div_exact tools/check_div_exact.c -lm
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static int check(float y, float s, int small) {
  const float r = 1.0f / s;
#ifdef SEQ4
  // four operations: r_lo = RN((1 - s r) r) per ROW; per element t = RN(y r_lo), qb = RN(y r + t) (one rounding of the almost exact
  // quotient: faithful), then ONE Markstein step
  const float r_lo = fmaf(-s, r, 1.0f) * r;
  const float t0 = y * r_lo;
  const float qb = fmaf(y, r, t0);
  const float q2 = fmaf(fmaf(-qb, s, y), r, qb);
#elif defined(SEQ3)
  const float q0 = y * r;
  const float q2 = fmaf(fmaf(-q0, s, y), r, q0);       // one step from the plain product: NOT exact (the checker must find cases)
#else
  const float q0 = y * r;
  const float q1 = fmaf(fmaf(-q0, s, y), r, q0);
  const float q2 = fmaf(fmaf(-q1, s, y), r, q1);
#endif
  const float x = y / s;
  if (!small && bits(q2) != bits(x) && !(q2 == 0.0f && x == 0.0f)) { printf("quotient mismatch y=%a s=%a q2=%a y/s=%a\n", y, s, q2, x); return 0; }
  const float t = q2 + 12582912.0f;
  const long e = (long)rintf(x);
  if (((uint32_t)(int32_t)e & 0xFFu) != (bits(t) & 0xFFu)) { printf("byte mismatch y=%a s=%a\n", y, s); return 0; }
  return 1;
}

int main(int argc, char** argv) {
  const long N = argc > 1 ? atol(argv[1]) : 2000000000L;
  long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
  for (long blk = 0; blk < 4096; ++blk) {
    uint64_t rng = 88172645463325252ull ^ (0x9E3779B97F4A7C15ull * (uint64_t)(blk + 1));
#define XR() (rng ^= rng << 13, rng ^= rng >> 7, rng ^= rng << 17, (uint32_t)(rng >> 16))
    for (long i = 0; i < N / 4096 && !bad; ++i) {
      float amax = expf(logf(1e-5f) + (XR() / 4294967296.0f) * (logf(1e6f) - logf(1e-5f)));
      float s = fmaxf(amax, 1e-5f) / 127.0f;
      const uint32_t k = XR();
      if ((k & 0xF0) == 0x10) s = from_bits(bits(s) | 0x7FFFFFu);          // significand all ones
      if ((k & 0xF0) == 0x20) s = from_bits(bits(s) & 0xFF800000u);        // power of two
      if ((k & 0xF0) == 0x30) s = from_bits((bits(s) & 0xFF800000u) | (XR() & 0x7FFFFFu));   // any significand
      amax = s * 127.0f;
      float y;
      if ((k & 3) == 0) {
        const int m = (int)(XR() % 256) - 128;
        y = ((float)m + 0.5f) * s;
        y = from_bits(bits(y) + (XR() % 17) - 8);
      } else if ((k & 3) == 1) {
        const int m = (int)(XR() % 255) - 127;                             // near an integer multiple (exact quotients)
        y = (float)m * s;
        y = from_bits(bits(y) + (XR() % 5) - 2);
      } else {
        y = ((XR() / 2147483648.0f) - 1.0f) * amax;
      }
      if (!(fabsf(y) <= amax * 1.000001f)) y = copysignf(amax, (k & 4) ? 1.0f : -1.0f);    // also: a zero dividend whose perturbed bits are a NaN
      // |y| < s/4 (incl. zeros and denormals, where the residuals underflow): only the quantized byte (0) is compared
      if (!check(y, s, fabsf(y) < 0.25f * s)) ++bad;
    }
  }
  if (bad) { printf("FAILED\n"); return 1; }
  printf("%ld cases: q2 == y / s bit for bit, low byte of q2 + 1.5*2^23 == (int8) rint(y / s): 0 mismatches\n", N / 4096 * 4096);
  return 0;
}
