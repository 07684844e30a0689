// ot_generator_argmax: generator.py:14-15  log_softmax(Linear(512 -> vocab)(h)) followed by the greedy arg-max of
// parallelized_inject_onnx_transformer.py:753-756 (torch.max: first index on ties).
// fp32 FMA GEMM (the reference runs this layer un-quantized in fp32 on the host) + one row-reduction kernel.
// Bound by the 9.1 MB fp32 weight stream per greedy step.
#include "ot_common.h"

namespace ot {

constexpr int kGenVT = 32;    // vocab entries per CTA
constexpr int kGenKC = 64;    // K chunk (one chunk of FMAs ~ one L2 round trip: the register prefetch of the next chunk is hidden)
constexpr int kGenRows = 64;  // h rows per CTA pass
constexpr int kGenPH = kGenRows + 4;   // shared pitches (floats): rows of 16-byte aligned float4, conflict-free broadcasts
constexpr int kGenPW = kGenVT + 4;

// logits[r, v] = bias[v] + sum_k h[r,k] * W[v,k]  (k ascending, fmaf: same order as a plain sequential dot product).
// CTA: 64 rows x 32 vocab entries, 128 threads, 4x4 register tile per thread; the operand chunks are stored k-major in
// shared memory so that a thread's 4 rows / 4 vocab entries are one 128-bit load and a warp touches 64 B + 128 B per k
// (the previous 1x4 tiling was shared-memory-bandwidth bound: 25 us); the next chunk's global loads are prefetched into
// registers while the current one is multiplied.
__global__ void __launch_bounds__(128) generator_logits_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ W,
                                                               const float* __restrict__ bias, int rows, int d, int vocab,
                                                               float* __restrict__ logits) {
  __shared__ __align__(16) float hs[kGenKC][kGenPH];
  __shared__ __align__(16) float ws[kGenKC][kGenPW];
  const unsigned int tl = tl_begin(7);
  pdl_wait();      // upstream results are complete and visible from here on
  pdl_trigger();   // now let exactly one successor start its launch + prologue (look-ahead depth 1)
  tl_mark(tl, 2);
  const int v0 = blockIdx.x * kGenVT;
  const int r0 = blockIdx.y * kGenRows;
  const int tid = threadIdx.x;
  const int tx = tid & 7;         // vocab group: entries 4*tx .. 4*tx+3
  const int ty = tid >> 3;        // row group:   rows    4*ty .. 4*ty+3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // global -> register staging: h chunk = 64 rows x 16 float4 (8 per thread), W chunk = 32 rows x 16 float4 (4 per thread)
  float4 hreg[8], wreg[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 128, r = idx >> 4, c = idx & 15;
      hreg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < rows && k0 + c * 4 < d) hreg[i] = __ldg(reinterpret_cast<const float4*>(h + static_cast<int64_t>(r0 + r) * ldh + k0) + c);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 128, r = idx >> 4, c = idx & 15;
      wreg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v0 + r < vocab && k0 + c * 4 < d) wreg[i] = __ldg(reinterpret_cast<const float4*>(W + static_cast<int64_t>(v0 + r) * d + k0) + c);
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 128, r = idx >> 4, c = idx & 15;
      hs[c * 4][r] = hreg[i].x; hs[c * 4 + 1][r] = hreg[i].y; hs[c * 4 + 2][r] = hreg[i].z; hs[c * 4 + 3][r] = hreg[i].w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 128, r = idx >> 4, c = idx & 15;
      ws[c * 4][r] = wreg[i].x; ws[c * 4 + 1][r] = wreg[i].y; ws[c * 4 + 2][r] = wreg[i].z; ws[c * 4 + 3][r] = wreg[i].w;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < d; k0 += kGenKC) {
    __syncthreads();            // previous chunk fully consumed
    stash();
    __syncthreads();
    if (k0 + kGenKC < d) fetch(k0 + kGenKC);
#pragma unroll 8
    for (int k = 0; k < kGenKC; ++k) {
      const float4 hv = *reinterpret_cast<const float4*>(&hs[k][ty * 4]);
      const float4 wv = *reinterpret_cast<const float4*>(&ws[k][tx * 4]);
      const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
      const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(hh[i], ww[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int v = v0 + tx * 4 + j;
    if (v < vocab) {
      const float bv = bias ? __ldg(bias + v) : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty * 4 + i;
        if (r < rows) logits[static_cast<int64_t>(r) * vocab + v] = __fadd_rn(acc[i][j], bv);
      }
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
  OT_CHECK_CUDA(launch_kernel(generator_logits_kernel, grid, dim3(128), 0, s, 1, h, ldh, Wg, bg, rows, d, vocab, scratch_logits));
  OT_CHECK_CUDA(launch_kernel(generator_reduce_kernel, dim3(rows), dim3(256), 0, s, 1, scratch_logits, vocab, next_ids, logp, margin));
  count_launch(2);
  return OT_OK;
}
