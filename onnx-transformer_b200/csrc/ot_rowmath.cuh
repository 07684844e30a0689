// Row-local fp32 math shared by the row kernels (ot_rowops.cu) and the persistent decoder kernel (ot_decoder.cu): both
// execute exactly these instruction sequences, so their results are bit-identical.  Op order follows SURVEY.md App. A.
#pragma once
#include "ot_common.h"

namespace ot {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Max of a NON-NEGATIVE quantity over the warp as one redux.sync: IEEE floats >= 0 order like their bit patterns.  A NaN lane
// counts as 0 -- fmaxf ignores NaN operands the same way, and every user clamps the result with fmaxf(., 1e-5f), so the all-NaN
// corner gives the same scale as the shuffle tree of warp_max.
__device__ __forceinline__ float warp_max_nonneg(float v) {
  return __uint_as_float(__reduce_max_sync(0xffffffffu, (v == v) ? __float_as_uint(v) : 0u));
}

// Max of ANY floats over the warp as one redux.sync through the order-preserving map float -> uint (flip all bits of negatives,
// the sign bit of non-negatives).  NaN lanes count as -inf, as fmaxf ignores them.
__device__ __forceinline__ float warp_max_any(float v) {
  uint32_t u = (v == v) ? __float_as_uint(v) : 0xff800000u;
  u ^= (u & 0x80000000u) ? 0xffffffffu : 0x80000000u;
  u = __reduce_max_sync(0xffffffffu, u);
  u ^= (u & 0x80000000u) ? 0x80000000u : 0xffffffffu;
  return __uint_as_float(u);
}

// x / 127 rounded exactly like IEEE division, for 0 or 1e-5 <= x <= 1e30, without the div.rn expansion (range check + branch to a
// subroutine): r = RN(1/127) refined by one Newton step, quotient corrected by one exact FMA residual.  tools/check_div127.c
// compares the sequence with x / 127.0f for EVERY float in [1e-5, 1e30] (975,318,303 values): 0 mismatches.  Outside that range the
// true division is used.
__device__ __forceinline__ float div127_exact(float x) {
  const float r = 0.007874015718698502f;                 // RN(1/127)
  const float r2 = __fmaf_rn(__fmaf_rn(-127.0f, r, 1.0f), r, r);
  const float q0 = __fmul_rn(x, r2);
  const float q1 = __fmaf_rn(__fmaf_rn(-127.0f, q0, x), r2, q0);
  if (__builtin_expect(!(x == 0.0f || (x >= 1e-5f && x <= 1e30f)), 0)) return __fdiv_rn(x, 127.0f);
  return q1;
}
__device__ __forceinline__ float quant_scale_x(float amax) { return div127_exact(fmaxf(amax, 1e-5f)); }
// x / 511 (the N-1 of the 512-feature LayerNorm), same construction, checked for every float in [1e-37, 1e37] (2,062,065,881 values).
__device__ __forceinline__ float div511_exact(float x) {
  const float r = 1.0f / 511.0f;                         // RN(1/511), folded at compile time
  const float r2 = __fmaf_rn(__fmaf_rn(-511.0f, r, 1.0f), r, r);
  const float q0 = __fmul_rn(x, r2);
  const float q1 = __fmaf_rn(__fmaf_rn(-511.0f, q0, x), r2, q0);
  if (__builtin_expect(!(x == 0.0f || (x >= 1e-37f && x <= 1e37f)), 0)) return __fdiv_rn(x, 511.0f);
  return q1;
}

// RowQuant of quant_linear.py:31-43: s = max(amax, 1e-5) / 127 ; q = rint(x / s).
__device__ __forceinline__ float quant_scale(float amax) { return __fdiv_rn(fmaxf(amax, 1e-5f), 127.0f); }
__device__ __forceinline__ int quant_one(float x, float s) { return __float2int_rn(rintf(__fdiv_rn(x, s))); }
// rint(y / s) with the quotient rounded exactly like IEEE division, without the div.rn expansion (whose special-case path is taken
// for every zero dividend -- half of a ReLU output -- and whose branch serialises independent elements) and without the conversion
// pipe (FRND / F2I are quarter rate and share the MIO queue with the shared-memory instructions around them):
//   q1 = y*r corrected by one FMA residual step (r = RN(1/s)) is within 1 ulp of RN(y/s);
//   t = q1 + 1.5*2^23 rounds q1 to an integer, ties to even, exactly as rintf does (|q1| < 2^22): n = t - 1.5*2^23, and the low byte
//   of t's bit pattern is n mod 256 -- the int8 two's-complement byte;
//   rint(q1) can only differ from rint(RN(y/s)) when q1 sits within 2^-16 of a half-integer (|y/s| <= 127, so 1 ulp <= 2^-17):
//   exactly then -- or when q1 is not a small finite number -- the true division decides (flagged, redone by the caller).
__device__ __forceinline__ uint32_t quant_fast_bits(float y, float s, float r, bool& slow) {
  const float q0 = __fmul_rn(y, r);
  const float rem = __fmaf_rn(-q0, s, y);
  const float q1 = __fmaf_rn(rem, r, q0);
  const float t = __fadd_rn(q1, 12582912.0f);
  const float n = __fsub_rn(t, 12582912.0f);
  slow = slow || (fabsf(fabsf(q1 - n) - 0.5f) < 1.52587890625e-05f) || !(fabsf(q1) < 1024.0f);
  return __float_as_uint(t);
}
// The exact form of one element, out of line: rint(y / s) with the IEEE division, as the low byte of the result.
static __device__ __noinline__ uint32_t quant_exact_bits(float y, float s) {
  return static_cast<uint32_t>(__float2int_rn(rintf(__fdiv_rn(y, s))));
}
// Redo of a chunk whose fast pass raised `slow`: the test of quant_fast_bits is re-evaluated per element (same instructions, same
// bits) and ONLY the elements that need it take the division.  A warp-chunk (512 elements) raises `slow` with probability ~1.5 %,
// nearly always for a single element: the redo costs ~150 issue slots instead of 16 divisions per lane through local memory
// (measured in ot_gemm_stream.cu: the old whole-chunk fallback made one warp in ~15 a 3-6 us straggler that the whole cluster then
// waited for at the row-maximum exchange).
template <int CW>
__device__ __forceinline__ void quant_redo_chunk(const float (&y)[CW], float s, float r, bool all, uint32_t (&tb)[CW]) {
#pragma unroll
  for (int j = 0; j < CW; ++j) {
    bool f = all;
    (void)quant_fast_bits(y[j], s, r, f);
    if (f) tb[j] = quant_exact_bits(y[j], s);
  }
}

// Leaner fast pass for callers that have already established that y, s and r are plain finite numbers with |y / s| <= 127 + tiny
// (the row's scale comes from the row's own maximum): q = y*r is within 2^-16 of y/s (r = RN(1/s): relative error 2^-24, the product
// another 2^-24, |q| <= 127), so rint(q) == rint(RN(y/s)) unless q sits within 2^-15 of a half-integer -- exactly then `slow` is
// raised and the element takes the true division (quant_redo_chunk2).  3 FP32 instructions fewer per element than quant_fast_bits.
__device__ __forceinline__ uint32_t quant_fast2_bits(float y, float r, bool& slow) {
  const float q = __fmul_rn(y, r);
  const float t = __fadd_rn(q, 12582912.0f);
  const float n = __fsub_rn(t, 12582912.0f);
  slow = slow || (fabsf(fabsf(q - n) - 0.5f) < 3.0517578125e-05f);
  return __float_as_uint(t);
}
template <int CW>
__device__ __forceinline__ void quant_redo_chunk2(const float (&y)[CW], float s, float r, bool all, uint32_t (&tb)[CW]) {
#pragma unroll
  for (int j = 0; j < CW; ++j) {
    bool f = all;
    (void)quant_fast2_bits(y[j], r, f);
    if (f) tb[j] = quant_exact_bits(y[j], s);
  }
}

// pack4(quant_one(v.x, s), ...) of four values, same bits as the division form
__device__ __forceinline__ uint32_t quant4_pack(float4 v, float s, float r) {
  bool slow = false;
  const uint32_t a = quant_fast_bits(v.x, s, r, slow), b = quant_fast_bits(v.y, s, r, slow);
  const uint32_t c = quant_fast_bits(v.z, s, r, slow), d = quant_fast_bits(v.w, s, r, slow);
  uint32_t w = __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
  if (__builtin_expect(slow, 0)) {
    const int ia = quant_one(v.x, s), ib = quant_one(v.y, s), ic = quant_one(v.z, s), id = quant_one(v.w, s);
    w = (static_cast<uint32_t>(ia) & 0xFFu) | ((static_cast<uint32_t>(ib) & 0xFFu) << 8) | ((static_cast<uint32_t>(ic) & 0xFFu) << 16) |
        ((static_cast<uint32_t>(id) & 0xFFu) << 24);
  }
  return w;
}
__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
  return (static_cast<uint32_t>(a) & 0xFFu) | ((static_cast<uint32_t>(b) & 0xFFu) << 8) |
         ((static_cast<uint32_t>(c) & 0xFFu) << 16) | ((static_cast<uint32_t>(d) & 0xFFu) << 24);
}

// LayerNorm (layer_norm.py:12-15 as exported) of one row held by a warp: lane owns float4 i*32+lane, i < VEC (n = 128*VEC).
//   mu = mean(x); d = x-mu; v = mean(d*d)*n/(n-1); y = (a*d)/(sqrt(v)+eps) + b
// On return v holds y; the result is this lane's max |y| (warp_max of it = the row's abs-max for RowQuant).
template <int VEC>
__device__ __forceinline__ float layernorm_row(float4 (&v)[VEC], int lane, int n, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps) {
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float nf = static_cast<float>(n);
  // n = 512 (every LayerNorm of this model): the divisions by N and N-1 as exact equivalents (a power-of-two scaling; the FMA
  // sequence of div511_exact, verified for every float) -- same bits, no div.rn expansion; any other n takes the IEEE divisions
  const bool n512 = (n == 512);
  const float tot = warp_sum(sum);
  const float mu = n512 ? __fmul_rn(tot, 0.001953125f) : __fdiv_rn(tot, nf);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i].x = __fsub_rn(v[i].x, mu);
    v[i].y = __fsub_rn(v[i].y, mu);
    v[i].z = __fsub_rn(v[i].z, mu);
    v[i].w = __fsub_rn(v[i].w, mu);
    sq += (__fmul_rn(v[i].x, v[i].x) + __fmul_rn(v[i].y, v[i].y)) + (__fmul_rn(v[i].z, v[i].z) + __fmul_rn(v[i].w, v[i].w));
  }
  const float tsq = warp_sum(sq);
  float var = n512 ? __fmul_rn(tsq, 0.001953125f) : __fdiv_rn(tsq, nf);            // ReduceMean(d*d)
  var = n512 ? div511_exact(__fmul_rn(var, nf)) : __fdiv_rn(__fmul_rn(var, nf), nf - 1.0f);            // * N / (N-1)
  const float denom = __fadd_rn(__fsqrt_rn(var), eps);       // sqrt + eps (eps added to std)
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 g = g4[i * 32 + lane];   // plain loads: gamma/beta may be staged in shared memory
    const float4 b = b4[i * 32 + lane];
    v[i].x = __fadd_rn(__fdiv_rn(__fmul_rn(g.x, v[i].x), denom), b.x);
    v[i].y = __fadd_rn(__fdiv_rn(__fmul_rn(g.y, v[i].y), denom), b.y);
    v[i].z = __fadd_rn(__fdiv_rn(__fmul_rn(g.z, v[i].z), denom), b.z);
    v[i].w = __fadd_rn(__fdiv_rn(__fmul_rn(g.w, v[i].w), denom), b.w);
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
  }
  return amax;
}

}  // namespace ot
