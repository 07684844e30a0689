#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
// exhaustive check: for every float x in [lo, hi], does the 3-FMA sequence with r = RN(1/D) reproduce x / D (IEEE RN)?
//   gcc -O2 -ffp-contract=off -o check tools/check_div127.c -lm && ./check            (D = 127, [1e-5, 1e30]:   975,318,303 floats, 0 mismatches)
//   sed s/127.0f/511.0f/g and [1e-37, 1e37] for the LayerNorm N-1 divisor             (2,062,065,881 floats, 0 mismatches)
static inline float seq(float x) {
  const float r = 1.0f / 127.0f;            // RN(1/127)
  const float e = fmaf(-127.0f, r, 1.0f);
  const float r2 = fmaf(e, r, r);
  const float q0 = x * r2;
  const float rem = fmaf(-127.0f, q0, x);
  return fmaf(rem, r2, q0);
}
int main() {
  float lo = 1e-5f, hi = 1e30f;
  uint32_t a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4);
  uint64_t bad = 0, n = 0;
  for (uint32_t u = a; u <= b; ++u) {
    float x; memcpy(&x, &u, 4);
    float t = x / 127.0f, s = seq(x);
    if (memcmp(&t, &s, 4) != 0) { if (bad < 5) printf("mismatch x=%a true=%a seq=%a\n", x, t, s); ++bad; }
    ++n;
  }
  printf("checked %llu floats, mismatches %llu\n", (unsigned long long)n, (unsigned long long)bad);
  return bad != 0;
}
