// ot_generator_argmax: generator.py:14-15  log_softmax(Linear(512 -> vocab)(h)) followed by the greedy arg-max of
// parallelized_inject_onnx_transformer.py:753-756 (torch.max: first index on ties).
// fp32 FMA GEMM (the reference runs this layer un-quantized in fp32 on the host) + one row-reduction kernel.
// Bound by the 9.1 MB fp32 weight stream per greedy step.
#include "ot_common.h"

namespace ot {

constexpr int kGenVT = 16;    // vocab rows per CTA
constexpr int kGenKC = 128;   // K chunk
constexpr int kGenRows = 64;  // h rows per CTA pass

// logits[r, v] = bias[v] + sum_k h[r,k] * W[v,k];  CTA: 64 rows x 16 vocab entries, 256 threads, 4 outputs each.
__global__ void __launch_bounds__(256) generator_logits_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ W,
                                                               const float* __restrict__ bias, int rows, int d, int vocab,
                                                               float* __restrict__ logits) {
  __shared__ float hs[kGenRows][kGenKC + 1];
  __shared__ float ws[kGenVT][kGenKC + 1];
  const unsigned int tl = tl_begin(7);
  pdl_wait();      // upstream results are complete and visible from here on
  pdl_trigger();   // now let exactly one successor start its launch + prologue (look-ahead depth 1)
  tl_mark(tl, 2);
  const int v0 = blockIdx.x * kGenVT;
  const int r0 = blockIdx.y * kGenRows;
  const int tid = threadIdx.x;
  const int vl = tid & 15;        // vocab entry within tile
  const int rl = tid >> 4;        // 0..15 -> rows rl, rl+16, rl+32, rl+48
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < d; k0 += kGenKC) {
    for (int idx = tid; idx < kGenRows * (kGenKC / 4); idx += 256) {
      const int r = idx / (kGenKC / 4), c = idx % (kGenKC / 4);
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < rows && k0 + c * 4 < d) t = __ldg(reinterpret_cast<const float4*>(h + static_cast<int64_t>(r0 + r) * ldh + k0) + c);
      hs[r][c * 4] = t.x; hs[r][c * 4 + 1] = t.y; hs[r][c * 4 + 2] = t.z; hs[r][c * 4 + 3] = t.w;
    }
    for (int idx = tid; idx < kGenVT * (kGenKC / 4); idx += 256) {
      const int r = idx / (kGenKC / 4), c = idx % (kGenKC / 4);
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v0 + r < vocab && k0 + c * 4 < d) t = __ldg(reinterpret_cast<const float4*>(W + static_cast<int64_t>(v0 + r) * d + k0) + c);
      ws[r][c * 4] = t.x; ws[r][c * 4 + 1] = t.y; ws[r][c * 4 + 2] = t.z; ws[r][c * 4 + 3] = t.w;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kGenKC; ++k) {
      const float w = ws[vl][k];
      acc[0] = fmaf(hs[rl][k], w, acc[0]);
      acc[1] = fmaf(hs[rl + 16][k], w, acc[1]);
      acc[2] = fmaf(hs[rl + 32][k], w, acc[2]);
      acc[3] = fmaf(hs[rl + 48][k], w, acc[3]);
    }
    __syncthreads();
  }
  if (v0 + vl < vocab) {
    const float bv = bias ? __ldg(bias + v0 + vl) : 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int r = r0 + rl + 16 * t;
      if (r < rows) logits[static_cast<int64_t>(r) * vocab + v0 + vl] = __fadd_rn(acc[t], bv);
    }
  }
  tl_mark(tl, 3);
}

// Per row: first arg-max, top1-top2 margin, optional log-softmax in place of the logits.
__global__ void __launch_bounds__(256) generator_reduce_kernel(float* __restrict__ logits, int vocab, int64_t* __restrict__ next_ids,
                                                               float* __restrict__ logp, float* __restrict__ margin) {
  __shared__ float s_val[8], s_second[8], s_sum[8];
  __shared__ int s_idx[8];
  const unsigned int tl = tl_begin(8);
  pdl_wait();      // upstream results are complete and visible from here on
  pdl_trigger();   // now let exactly one successor start its launch + prologue (look-ahead depth 1)
  tl_mark(tl, 2);
  const int row = blockIdx.x;
  const float* x = logits + static_cast<int64_t>(row) * vocab;
  // torch.max / np.argmax semantics: a NaN logit (a fault can produce one) ranks above every number, the first one wins;
  // NaNs are mapped to +inf for the comparison so the returned index is always a valid token id.
  float best = -INFINITY, second = -INFINITY;
  int bidx = 0x7fffffff;
  for (int v = threadIdx.x; v < vocab; v += blockDim.x) {
    float t = x[v];
    if (t != t) t = INFINITY;
    if (t > best || (t == best && v < bidx)) { second = best; best = t; bidx = v; }
    else if (t > second) second = t;
  }
  // warp then block reduction of (best, first index, second)
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    const float os = __shfl_xor_sync(0xffffffffu, second, o);
    if (ob > best || (ob == best && oi < bidx)) { second = fmaxf(best, os); best = ob; bidx = oi; }
    else second = fmaxf(second, ob);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_val[warp] = best; s_idx[warp] = bidx; s_second[warp] = second; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = s_val[0], s2 = s_second[0];
    int bi = s_idx[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) {
      const float ob = s_val[w]; const int oi = s_idx[w]; const float os = s_second[w];
      if (ob > b || (ob == b && oi < bi)) { s2 = fmaxf(b, os); b = ob; bi = oi; }
      else s2 = fmaxf(s2, ob);
    }
    s_val[0] = b; s_idx[0] = bi; s_second[0] = s2;
    next_ids[row] = (bi >= 0 && bi < vocab) ? bi : 0;
    if (margin) margin[row] = b - s2;
  }
  __syncthreads();
  if (logp) {
    const float mx = s_val[0];
    float sum = 0.f;
    for (int v = threadIdx.x; v < vocab; v += blockDim.x) sum += expf(x[v] - mx);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_sum[warp] = sum;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += s_sum[w];
    const float lse = mx + logf(tot);
    float* out = logp + static_cast<int64_t>(row) * vocab;
    for (int v = threadIdx.x; v < vocab; v += blockDim.x) out[v] = x[v] - lse;
  }
  tl_mark(tl, 3);
}

OT_DEFINE_TL_SETTER(tl_set_generator)

}  // namespace ot

using namespace ot;

// `logits` doubles as scratch: the caller passes a [rows, vocab] fp32 buffer in `logp` or (when it does not want
// log-probabilities) in `scratch`.
extern "C" int ot_generator_argmax(const float* h, int64_t ldh, const float* Wg, const float* bg, int rows, int d, int vocab,
                                   int64_t* next_ids, float* scratch_logits, float* logp, float* margin, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(h && Wg && next_ids && scratch_logits, "null operand (scratch_logits [rows, vocab] is required)");
  OT_REQUIRE(rows > 0 && d % 4 == 0 && ldh % 4 == 0 && vocab > 1, "bad generator shape");
  cudaStream_t s = as_stream(stream);
  dim3 grid((vocab + kGenVT - 1) / kGenVT, (rows + kGenRows - 1) / kGenRows, 1);
  OT_CHECK_CUDA(launch_kernel(generator_logits_kernel, grid, dim3(256), 0, s, 1, h, ldh, Wg, bg, rows, d, vocab, scratch_logits));
  OT_CHECK_CUDA(launch_kernel(generator_reduce_kernel, dim3(rows), dim3(256), 0, s, 1, scratch_logits, vocab, next_ids, logp, margin));
  count_launch(2);
  return OT_OK;
}
