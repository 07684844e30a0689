// ot_decoder_run: the KV-cached greedy decoder as ONE persistent kernel (fault-free fast path of greedy_decode,
// parallelized_inject_onnx_transformer.py:616-758 / batch_output.py:659-672).
//
// A greedy step at batch 64 is a chain of ~70 dependent operations of a few hundred KB each: as separate launches it is
// bound by launch/prologue/drain latency (~5.8 us per link), not by the tensor cores or by HBM.  Here one CTA per SM stays
// resident for all requested steps; the links of the chain become phases separated by a grid-wide barrier (one L2 atomic +
// one polled load, ~0.7 us).  Barriers are only placed where a whole-row dependency forces them:
//
//   LN      warp per sentence row: [dequant + bias + residual of the previous GEMM's int32 accumulators | embedding+PE]
//           -> residual stream x (fp32) -> LayerNorm -> RowQuant -> xq int8 + scale          (a6, a7, a17/a18, a20)
//   GEMM    one 64 x 16 output tile x 512-deep contraction slice per CTA: TMA (128B swizzle) -> tcgen05.mma kind::i8 with
//           int32 accumulators in TMEM -> tcgen05.ld -> raw int32 partials to L2 (split-K slices in separate planes:
//           integer partial sums are exact in any order).  The fp32 epilogue runs in the consumer phase, on complete rows,
//           in the canonical order y = fl(fl(float(acc)*sx[m])*sw[n]) + b[n] -- the same instructions as ot_linear_w8a8.
//   FFN1    as GEMM, but the consumer is another GEMM (needs int8 operand tiles), so its epilogue (bias, ReLU, RowQuant over
//           2048 features) runs in place: row abs-max through one atomicMax per row, a second grid barrier, then quantize
//           from registers.                                                                   (a9, a10, a18)
//   ATTN    CTA per sentence: Q/K/V (or cross-Q) epilogue + RowQuant of its own row, KV-cache append, then the decode
//           attention body shared with attention_decode_kernel (ot_attention_decode.cuh)       (a10-a17, a19)
//   GEN     fp32 generator GEMM on CUDA cores, 64 x 32 logits per CTA, cp.async 3-stage pipeline, per-tile first-arg-max;
//           then one warp per sentence reduces the tile maxima and appends the token          (a21, a22)
//
// Everything a phase reads that another CTA wrote in an earlier phase is read with ld.global.cg (L2) or by TMA.
// Arithmetic is instruction-for-instruction that of the stand-alone kernels, so tokens and KV caches are bit-identical
// to the graph-replay engine path (tests/test_decoder_gpu.py).
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include "ot_attention_decode.cuh"
#include "ot_common.h"
#include "ot_ptx.cuh"
#include "ot_rowmath.cuh"

namespace ot {

int get_tensor_map(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols,
                   bool swizzle128);

constexpr int kD = 512;
constexpr int kFF = 2048;
constexpr int kMkThreads = 256;
constexpr int kMkBN = 16;        // output columns per GEMM tile (tcgen05 N)
constexpr int kMkKB = 4;         // 128-byte k-blocks per tile: contraction slice of 512
constexpr int kMkRows = 64;      // sentence rows per launch (TMA box rows); the MMA is M = 128, rows 64..127 are never read back
constexpr int kMkMaxLayers = 8;
constexpr int kGenVT = 32;       // vocab entries per generator tile
constexpr int kGenKC = 64;       // generator K chunk
constexpr int kGenPitch = kGenKC + 4;
constexpr int kGenStages = 3;

// shared-memory map (dynamic, base aligned to 1024)
constexpr int kSmemA = 0;                                   // [kMkKB][64 x 128 B]
constexpr int kSmemB = kMkKB * kMkRows * 128;               // [kMkKB][16 x 128 B]  (also the phantom rows 64..127 of the last A block)
constexpr int kSmemGemmEnd = kSmemB + kMkKB * kMkBN * 128;  // 40960
constexpr int kSmemVh = 0;                                  // attention: [8][96][64] int8 = 49152
constexpr int kSmemRow = 49152;                             // attention: this sentence's quantized q | k | v row (1536 B)
constexpr int kSmemRed = kSmemRow + 1536;                   // 64 floats of reduction scratch
constexpr int kSmemGen = 0;                                 // generator: kGenStages x (64 + 32) x kGenPitch floats = 78336
constexpr int kSmemBars = 96 * 1024;                        // mbarriers + TMEM slot
constexpr int kSmemHot = kSmemBars + 128;                   // MkHot copy (<= 8 KB)
constexpr int kSmemTotal = 128 * 1024;                      // > half an SM: exactly one CTA per SM
static_assert(kGenStages * (64 + kGenVT) * kGenPitch * 4 <= kSmemBars, "generator stages overflow");
static_assert(kSmemGemmEnd + 0 <= kSmemBars && kSmemRed + 256 <= kSmemBars, "smem map overflow");

struct MkLayer {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
  const float *qkv_sw, *qkv_b, *o_sw, *o_b, *cq_sw, *cq_b, *co_sw, *co_b, *w1_sw, *w1_b, *w2_sw, *w2_b;
  int8_t *kc, *vc;     // self-attention KV cache [B, cap, 512]
  float *skc, *svc;    // [B, cap]
};

// Pointers and sizes: copied to shared memory at kernel start (the grid barrier's fences invalidate L1, and a phase must not
// begin with a dependent L2 round trip just to learn where its operands are).
struct MkHot {
  MkLayer layer[kMkMaxLayers];
  int n_layers, B, S, cap, vocab;
  float emb_scale;
  float* x;              // [B, 512] residual stream
  int8_t* xq; float* sx;
  int32_t* acc;          // raw accumulators: [ksplit][64][N]
  int8_t* cq; float* cs;
  int8_t* hq; float* sh;
  unsigned int* rowmax;  // [n_layers][64] FFN1 row abs-max (float bits)
  const int8_t* ckv; const float* sckv;       // cross K/V projections [B*S, 2*512*n_layers], scales [B*S, 2*n_layers]
  const uint8_t* mask;   // [B, S]
  const float *fin_g, *fin_b;
  float* hout;           // [B, 512]
  const float *gen_w, *gen_b;
  float* gen_pv; int* gen_pi;                  // per generator tile, per row: best logit / its index
  const float *tgt_lut, *pe;
  int64_t* ys; int64_t ys_ld;
  unsigned int* bar;     // grid barrier counter (zeroed by the host before every launch)
  unsigned long long* trace;   // optional: %globaltimer of CTA 0 after every phase of the last step of a launch
};

struct MkPlan {
  CUtensorMap map_xq, map_cq, map_hq;          // A operands (activations); TMA descriptors stay in global memory
  CUtensorMap map_w[kMkMaxLayers][6];          // qkv, o, cq, co, w1, w2
  MkHot hot;
};
static_assert(sizeof(MkHot) % 16 == 0 && sizeof(MkHot) <= 8 * 1024, "MkHot is copied to shared memory in 16-byte pieces");

// ------------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ int4 ldcg_i4(const int32_t* p) { return __ldcg(reinterpret_cast<const int4*>(p)); }
__device__ __forceinline__ float4 ldcg_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct MkCtx {
  const MkHot* P;       // shared-memory copy
  const MkPlan* G;      // global: tensor maps
  uint8_t* smem;
  uint64_t* bars;       // full[kMkKB], tmem_full
  uint32_t tmem_base;
  unsigned int bar_target;
  uint32_t parity;      // of this CTA's GEMM mbarriers
  int trace_slot;
  bool trace_on;
};

// Grid-wide barrier: every thread's earlier global writes (generic proxy) are visible to every thread's later reads, through
// the generic proxy and through TMA.  Bounded spin: a protocol bug traps instead of hanging the box.
__device__ __forceinline__ void grid_sync(MkCtx& c) {
  asm volatile("fence.proxy.async;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  c.bar_target += gridDim.x;
  if (threadIdx.x == 0) {
    // release (cumulative over the CTA's writes ordered by the bar.sync above) + acquire: measured 1.4 us per barrier at 148
    // CTAs vs 1.9 us with explicit __threadfence() pairs (tools/bench_barrier.cu)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(c.P->bar) : "memory");
    unsigned int v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c.P->bar) : "memory");
      if (++spins > (1u << 24)) __trap();
    } while (static_cast<int>(v - c.bar_target) < 0);
    asm volatile("fence.proxy.async;" ::: "memory");
    if (c.trace_on) c.P->trace[c.trace_slot] = tl_now();
  }
  ++c.trace_slot;
  __syncthreads();
  tc_fence_after();
}

// ------------------------------------------------------------------------------------------------ GEMM tile
// Accumulate A[0:64, k0:k0+512] * W[n0:n0+16, k0:k0+512]^T into TMEM columns [0,16).  Warp 7 (one elected lane) issues the
// TMA loads and the MMAs; on return warps 0 and 1 (thread = row) hold the 16 int32 accumulators of their row in r[].
__device__ __forceinline__ void gemm_tile(MkCtx& c, const CUtensorMap* amap, const CUtensorMap* wmap, int k0, int n0, uint32_t (&r)[16]) {
  const int warp = threadIdx.x >> 5;
  uint8_t* sA = c.smem + kSmemA;
  uint8_t* sB = c.smem + kSmemB;
  if (warp == 7) {
    if (elect_one()) {
      for (int kb = 0; kb < kMkKB; ++kb) {
        const uint32_t fb = smem_u32(&c.bars[kb]);
        mbar_arrive_expect_tx(fb, kMkRows * 128 + kMkBN * 128);
        tma_load_2d(smem_u32(sA + kb * kMkRows * 128), amap, fb, k0 + kb * 128, 0);
        tma_load_2d(smem_u32(sB + kb * kMkBN * 128), wmap, fb, k0 + kb * 128, n0);
      }
      constexpr uint32_t idesc = make_idesc_i8(128, kMkBN);
      for (int kb = 0; kb < kMkKB; ++kb) {
        mbar_wait(smem_u32(&c.bars[kb]), c.parity);
        tc_fence_after();
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sA + kb * kMkRows * 128));
        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(sB + kb * kMkBN * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_i8_ss(c.tmem_base, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
      }
      mma_commit(smem_u32(&c.bars[kMkKB]));
    }
    __syncwarp();
  } else if (warp < 2) {
    mbar_wait(smem_u32(&c.bars[kMkKB]), c.parity);
    tc_fence_after();
    tmem_ld_32x16(c.tmem_base + (static_cast<uint32_t>(warp * 32) << 16), r);
    tmem_wait_ld();
  }
  c.parity ^= 1u;
}

// Plain GEMM phase: raw int32 partials -> acc[ks][row][N].
__device__ __forceinline__ void phase_gemm(MkCtx& c, const CUtensorMap* amap, const CUtensorMap* wmap, int N, int ksplit) {
  const MkHot& P = *c.P;
  const int n_tiles = N / kMkBN;
  const int warp = threadIdx.x >> 5;
  bool first = true;
  for (int tile = blockIdx.x; tile < n_tiles * ksplit; tile += gridDim.x) {
    if (!first) {   // smem / TMEM reuse inside one phase (only on GPUs with fewer CTAs than tiles)
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
    first = false;
    const int nt = tile % n_tiles, ks = tile / n_tiles;
    uint32_t r[16];
    gemm_tile(c, amap, wmap, ks * kMkKB * 128, nt * kMkBN, r);
    const int row = threadIdx.x;
    if (warp < 2 && row < P.B) {
      int4* dst = reinterpret_cast<int4*>(P.acc + (static_cast<int64_t>(ks) * kMkRows + row) * N + nt * kMkBN);
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = make_int4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
  }
}

// FFN1: h = RowQuant_2048(ReLU(x_hat W1^T + b1)) written as the int8 operand of FFN2.  Contains one grid barrier.
__device__ __forceinline__ void phase_ffn1(MkCtx& c, int l) {
  const MkHot& P = *c.P;
  const MkLayer& L = P.layer[l];
  const int warp = threadIdx.x >> 5;
  const int tile = blockIdx.x;
  const int row = threadIdx.x;
  const bool active = tile < kFF / kMkBN;
  const bool owner = active && warp < 2 && row < P.B;
  float y[16];
  if (active) {
    uint32_t r[16];
    gemm_tile(c, &c.G->map_xq, &c.G->map_w[l][4], 0, tile * kMkBN, r);
    if (owner) {
      const float sxr = __ldcg(P.sx + row);
      float amax = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float sw = __ldg(L.w1_sw + tile * kMkBN + j), bb = __ldg(L.w1_b + tile * kMkBN + j);
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(static_cast<int>(r[j])), sxr), sw), bb);
        y[j] = fmaxf(v, 0.0f);
        amax = fmaxf(amax, fabsf(y[j]));
      }
      atomicMax(P.rowmax + l * kMkRows + row, __float_as_uint(amax));
    }
  }
  grid_sync(c);
  if (owner) {
    const float s = quant_scale(__uint_as_float(__ldcg(P.rowmax + l * kMkRows + row)));
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = pack4(quant_one(y[4 * j], s), quant_one(y[4 * j + 1], s), quant_one(y[4 * j + 2], s), quant_one(y[4 * j + 3], s));
    *reinterpret_cast<uint4*>(P.hq + static_cast<int64_t>(row) * kFF + tile * kMkBN) = make_uint4(w[0], w[1], w[2], w[3]);
    if (tile == 0) P.sh[row] = s;
  }
}

// ------------------------------------------------------------------------------------------------ LN phase
// SRC 0: x = embedding(ys[:, t]) * sqrt(d) + pe[t]           (embeddings.py:13, positional_encodings.py:24)
// SRC 1: x = x + (fl(fl(float(sum_ks acc)*sa[row])*sw[n]) + b[n])     (epilogue + residual of the previous GEMM)
// then LayerNorm; quant: RowQuant -> xq, sx; else y -> hout.
template <int SRC>
__device__ __forceinline__ void phase_ln(MkCtx& c, int t, const float* gamma, const float* beta, bool quant, const float* a_scale,
                                         const float* sw, const float* bias, int ksplit) {
  const MkHot& P = *c.P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (SRC == 0 && blockIdx.x == gridDim.x - 1)
    for (int i = threadIdx.x; i < P.n_layers * kMkRows; i += blockDim.x) P.rowmax[i] = 0u;
  if (warp != 0) return;
  for (int row = blockIdx.x; row < P.B; row += gridDim.x) {
    float4 v[4];
    float* xr = P.x + static_cast<int64_t>(row) * kD;
    if (SRC == 0) {
      const int64_t id = __ldcg(P.ys + row * P.ys_ld + t);
      const float4* e4 = reinterpret_cast<const float4*>(P.tgt_lut + id * kD);
      const float4* p4 = reinterpret_cast<const float4*>(P.pe + static_cast<int64_t>(t) * kD);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 e = __ldg(e4 + i * 32 + lane), q = __ldg(p4 + i * 32 + lane);
        v[i] = make_float4(__fadd_rn(__fmul_rn(e.x, P.emb_scale), q.x), __fadd_rn(__fmul_rn(e.y, P.emb_scale), q.y),
                           __fadd_rn(__fmul_rn(e.z, P.emb_scale), q.z), __fadd_rn(__fmul_rn(e.w, P.emb_scale), q.w));
      }
    } else {
      const float sa = __ldcg(a_scale + row);
      int4 a[4];
      float4 res[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = ldcg_i4(P.acc + static_cast<int64_t>(row) * kD + (i * 32 + lane) * 4);
        res[i] = ldcg_f4(xr + (i * 32 + lane) * 4);
      }
      if (ksplit == 4) {          // all 12 partial loads in flight together
        int4 p[3][4];
#pragma unroll
        for (int ks = 1; ks < 4; ++ks)
#pragma unroll
          for (int i = 0; i < 4; ++i) p[ks - 1][i] = ldcg_i4(P.acc + (static_cast<int64_t>(ks) * kMkRows + row) * kD + (i * 32 + lane) * 4);
#pragma unroll
        for (int ks = 0; ks < 3; ++ks)
#pragma unroll
          for (int i = 0; i < 4; ++i) { a[i].x += p[ks][i].x; a[i].y += p[ks][i].y; a[i].z += p[ks][i].z; a[i].w += p[ks][i].w; }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(sw) + i * 32 + lane);
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + i * 32 + lane);
        v[i].x = __fadd_rn(res[i].x, __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a[i].x), sa), w4.x), b4.x));
        v[i].y = __fadd_rn(res[i].y, __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a[i].y), sa), w4.y), b4.y));
        v[i].z = __fadd_rn(res[i].z, __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a[i].z), sa), w4.z), b4.z));
        v[i].w = __fadd_rn(res[i].w, __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a[i].w), sa), w4.w), b4.w));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(xr)[i * 32 + lane] = v[i];
    const float amax = layernorm_row<4>(v, lane, kD, gamma, beta, 1e-6f);
    if (quant) {
      const float s = quant_scale(warp_max(amax));
      uint32_t* qr = reinterpret_cast<uint32_t*>(P.xq + static_cast<int64_t>(row) * kD);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        qr[i * 32 + lane] = pack4(quant_one(v[i].x, s), quant_one(v[i].y, s), quant_one(v[i].z, s), quant_one(v[i].w, s));
      if (lane == 0) P.sx[row] = s;
    } else {
      float4* yr = reinterpret_cast<float4*>(P.hout + static_cast<int64_t>(row) * kD);
#pragma unroll
      for (int i = 0; i < 4; ++i) yr[i * 32 + lane] = v[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention phases
// Epilogue + RowQuant (groups of 512 features) of this sentence's projection row: acc -> int8 row in shared memory.
// NG = 3: fused Q|K|V (ksplit 1, N = 1536); NG = 1: cross-attention Q (ksplit given, N = 512).
template <int NG>
__device__ __forceinline__ void project_row(MkCtx& c, int b, const float* a_scale, const float* sw, const float* bias, int ksplit, float (&scale)[3]) {
  const MkHot& P = *c.P;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int8_t* rowbuf = reinterpret_cast<int8_t*>(c.smem + kSmemRow);
  float* red = reinterpret_cast<float*>(c.smem + kSmemRed);
  constexpr int N = NG * kD;
  const float sa = __ldcg(a_scale + b);
  float4 y[2];
  float am[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = tid + 256 * j;           // float4 index within the row
    y[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < N / 4) {
      int4 a = ldcg_i4(P.acc + static_cast<int64_t>(b) * N + i * 4);
      if (ksplit == 4) {
        int4 p[3];
#pragma unroll
        for (int ks = 1; ks < 4; ++ks) p[ks - 1] = ldcg_i4(P.acc + (static_cast<int64_t>(ks) * kMkRows + b) * N + i * 4);
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) { a.x += p[ks].x; a.y += p[ks].y; a.z += p[ks].z; a.w += p[ks].w; }
      }
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(sw) + i);
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + i);
      y[j].x = __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a.x), sa), w4.x), b4.x);
      y[j].y = __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a.y), sa), w4.y), b4.y);
      y[j].z = __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a.z), sa), w4.z), b4.z);
      y[j].w = __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a.w), sa), w4.w), b4.w);
      am[j] = fmaxf(fmaxf(fabsf(y[j].x), fabsf(y[j].y)), fmaxf(fabsf(y[j].z), fabsf(y[j].w)));
    }
  }
  am[0] = warp_max(am[0]);
  am[1] = warp_max(am[1]);
  if (lane == 0) { red[warp] = am[0]; red[8 + warp] = am[1]; }
  __syncthreads();
  // float4 index i = tid (+256): group = i / 128 -> warps 0-3 of j=0: group 0, warps 4-7 of j=0: group 1, warps 0-3 of j=1: group 2
  scale[0] = quant_scale(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])));
  scale[1] = scale[2] = 0.f;
  if (NG == 3) {
    scale[1] = quant_scale(fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7])));
    scale[2] = quant_scale(fmaxf(fmaxf(red[8], red[9]), fmaxf(red[10], red[11])));
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = tid + 256 * j;
    if (i < N / 4) {
      const float s = scale[i >> 7];
      reinterpret_cast<uint32_t*>(rowbuf)[i] = pack4(quant_one(y[j].x, s), quant_one(y[j].y, s), quant_one(y[j].z, s), quant_one(y[j].w, s));
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void phase_self_attention(MkCtx& c, int t, int l) {
  const MkHot& P = *c.P;
  const MkLayer& L = P.layer[l];
  for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
    float sc[3];
    project_row<3>(c, b, P.sx, L.qkv_sw, L.qkv_b, 1, sc);
    const int8_t* rowbuf = reinterpret_cast<const int8_t*>(c.smem + kSmemRow);
    AttnArgs a = {};
    a.k = L.kc; a.v = L.vc; a.ldk = kD; a.sk = L.skc; a.sv = L.svc; a.skv_stride = 1;
    a.B = P.B; a.Tq = 1; a.Tk = t + 1; a.Tk_cap = P.cap; a.mask_kind = 2;
    a.key_mask = nullptr; a.mask_stride = 0; a.q_pos0 = t;
    a.ctx = nullptr; a.ld_ctx = 0; a.ctx_q = P.cq; a.ctx_s = P.cs;
    AttnDecRow r = {rowbuf, sc[0], rowbuf + kD, rowbuf + 2 * kD, sc[1], sc[2]};
    attention_decode_body(a, r, b, t + 1, t, reinterpret_cast<AttnDecVh>(c.smem + kSmemVh));
  }
}

__device__ __forceinline__ void phase_cross_attention(MkCtx& c, int l) {
  const MkHot& P = *c.P;
  const MkLayer& L = P.layer[l];
  for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
    float sc[3];
    project_row<1>(c, b, P.sx, L.cq_sw, L.cq_b, 4, sc);
    const int8_t* rowbuf = reinterpret_cast<const int8_t*>(c.smem + kSmemRow);
    AttnArgs a = {};
    a.k = const_cast<int8_t*>(P.ckv) + 2 * kD * l; a.v = const_cast<int8_t*>(P.ckv) + 2 * kD * l + kD; a.ldk = 2 * kD * P.n_layers;
    a.sk = const_cast<float*>(P.sckv) + 2 * l; a.sv = const_cast<float*>(P.sckv) + 2 * l + 1; a.skv_stride = 2 * P.n_layers;
    a.B = P.B; a.Tq = 1; a.Tk = P.S; a.Tk_cap = P.S; a.mask_kind = 1;
    a.key_mask = P.mask; a.mask_stride = P.S; a.q_pos0 = 0;
    a.ctx = nullptr; a.ld_ctx = 0; a.ctx_q = P.cq; a.ctx_s = P.cs;
    AttnDecRow r = {rowbuf, sc[0], nullptr, nullptr, 0.f, 0.f};
    attention_decode_body(a, r, b, P.S, 0, reinterpret_cast<AttnDecVh>(c.smem + kSmemVh));
  }
}

// ------------------------------------------------------------------------------------------------ generator
// logits[r, v] = bias[v] + sum_k h[r,k] * W[v,k]  (k ascending, fmaf: the order of generator_logits_kernel), 64 rows x 32
// vocab entries per tile; warp w owns rows 8w..8w+7, lane = vocab entry; operand chunks arrive through a 3-stage cp.async
// ring.  Each tile leaves, per row, its best logit and the first index attaining it (NaN ranks as +inf: torch.max).
__device__ __forceinline__ void phase_generator_logits(MkCtx& c) {
  const MkHot& P = *c.P;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* stage = reinterpret_cast<float*>(c.smem + kSmemGen);
  constexpr int kStageFloats = (64 + kGenVT) * kGenPitch;
  const int n_tiles = (P.vocab + kGenVT - 1) / kGenVT;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int v0 = tile * kGenVT;
    auto issue = [&](int chunk) {
      float* hs = stage + (chunk % kGenStages) * kStageFloats;
      float* ws = hs + 64 * kGenPitch;
      const int k0 = chunk * kGenKC;
#pragma unroll
      for (int i = 0; i < 4; ++i) {          // 64 rows x 16 float4
        const int idx = tid + i * 256, r = idx >> 4, q = idx & 15;
        cp_async16(smem_u32(hs + r * kGenPitch + q * 4), P.hout + static_cast<int64_t>(min(r, P.B - 1)) * kD + k0 + q * 4);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {          // 32 vocab rows x 16 float4
        const int idx = tid + i * 256, r = idx >> 4, q = idx & 15;
        cp_async16(smem_u32(ws + r * kGenPitch + q * 4), P.gen_w + static_cast<int64_t>(min(v0 + r, P.vocab - 1)) * kD + k0 + q * 4);
      }
      cp_async_commit();
    };
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    constexpr int kChunks = kD / kGenKC;
    issue(0);
    issue(1);
    for (int chunk = 0; chunk < kChunks; ++chunk) {
      if (chunk + 1 < kChunks) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncthreads();                        // chunk landed for everyone; stage (chunk+2)%3 = (chunk-1)%3 is free again
      if (chunk + 2 < kChunks) issue(chunk + 2);
      const float* hs = stage + (chunk % kGenStages) * kStageFloats + warp * 8 * kGenPitch;
      const float* ws = stage + (chunk % kGenStages) * kStageFloats + 64 * kGenPitch + lane * kGenPitch;
#pragma unroll 4
      for (int k = 0; k < kGenKC; k += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(ws + k);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float4 h4 = *reinterpret_cast<const float4*>(hs + r * kGenPitch + k);
          acc[r] = fmaf(h4.x, w4.x, acc[r]);
          acc[r] = fmaf(h4.y, w4.y, acc[r]);
          acc[r] = fmaf(h4.z, w4.z, acc[r]);
          acc[r] = fmaf(h4.w, w4.w, acc[r]);
        }
      }
    }
    const int v = v0 + lane;
    const bool valid = v < P.vocab;
    const float bv = (valid && P.gen_b) ? __ldg(P.gen_b + v) : 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float best = __fadd_rn(acc[r], bv);
      if (best != best) best = INFINITY;
      int bidx = v;
      if (!valid) { best = -INFINITY; bidx = 0x7fffffff; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
      }
      if (lane == 0) {
        P.gen_pv[tile * kMkRows + warp * 8 + r] = best;
        P.gen_pi[tile * kMkRows + warp * 8 + r] = bidx;
      }
    }
    __syncthreads();   // all warps done with the stages before the next tile's loads
  }
}

// ys[b, t+1] = first arg-max over the tile maxima (greedy_decode: parallelized_inject_onnx_transformer.py:753-758).
__device__ __forceinline__ void phase_generator_reduce(MkCtx& c, int t) {
  const MkHot& P = *c.P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp != 0) return;
  const int n_tiles = (P.vocab + kGenVT - 1) / kGenVT;
  for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    for (int tile = lane; tile < n_tiles; tile += 32) {
      const float ob = __ldcg(P.gen_pv + tile * kMkRows + b);
      const int oi = __ldcg(P.gen_pi + tile * kMkRows + b);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane == 0) P.ys[b * P.ys_ld + t + 1] = (bidx >= 0 && bidx < P.vocab) ? bidx : 0;
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kMkThreads, 1) decoder_steps_kernel(const MkPlan* __restrict__ plan, int t0, int n_steps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  MkCtx c;
  c.G = plan;
  c.smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  c.P = reinterpret_cast<const MkHot*>(c.smem + kSmemHot);
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(MkHot) / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(c.smem + kSmemHot)[i] = reinterpret_cast<const uint4*>(&plan->hot)[i];
  c.bars = reinterpret_cast<uint64_t*>(c.smem + kSmemBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c.smem + kSmemBars + 64);
  c.bar_target = 0;
  c.parity = 0;
  const MkHot& P = *c.P;
  const MkPlan& G = *plan;
  const int warp = threadIdx.x >> 5;

  if (warp == 7) {
    if (elect_one()) {
      for (int i = 0; i <= kMkKB; ++i) mbar_init(smem_u32(&c.bars[i]), 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = *tmem_slot;

  for (int t = t0; t < t0 + n_steps; ++t) {
    c.trace_slot = 0;
    c.trace_on = P.trace != nullptr && blockIdx.x == 0 && t == t0 + n_steps - 1;
    if (c.trace_on && threadIdx.x == 0) P.trace[127] = tl_now();
    for (int l = 0; l < P.n_layers; ++l) {
      const MkLayer& L = P.layer[l];
      // --- masked self-attention over the KV cache
      if (l == 0) phase_ln<0>(c, t, L.ln1_g, L.ln1_b, true, nullptr, nullptr, nullptr, 0);
      else phase_ln<1>(c, t, L.ln1_g, L.ln1_b, true, P.sh, P.layer[l - 1].w2_sw, P.layer[l - 1].w2_b, 4);
      grid_sync(c);
      phase_gemm(c, &G.map_xq, &G.map_w[l][0], 3 * kD, 1);
      grid_sync(c);
      phase_self_attention(c, t, l);
      grid_sync(c);
      phase_gemm(c, &G.map_cq, &G.map_w[l][1], kD, 4);
      grid_sync(c);
      // --- cross-attention over the cached memory projections
      phase_ln<1>(c, t, L.ln2_g, L.ln2_b, true, P.cs, L.o_sw, L.o_b, 4);
      grid_sync(c);
      phase_gemm(c, &G.map_xq, &G.map_w[l][2], kD, 4);
      grid_sync(c);
      phase_cross_attention(c, l);
      grid_sync(c);
      phase_gemm(c, &G.map_cq, &G.map_w[l][3], kD, 4);
      grid_sync(c);
      // --- feed forward
      phase_ln<1>(c, t, L.ln3_g, L.ln3_b, true, P.cs, L.co_sw, L.co_b, 4);
      grid_sync(c);
      phase_ffn1(c, l);          // one barrier inside
      grid_sync(c);
      phase_gemm(c, &G.map_hq, &G.map_w[l][5], kD, 4);
      grid_sync(c);
    }
    phase_ln<1>(c, t, P.fin_g, P.fin_b, false, P.sh, P.layer[P.n_layers - 1].w2_sw, P.layer[P.n_layers - 1].w2_b, 4);
    grid_sync(c);
    phase_generator_logits(c);
    grid_sync(c);
    phase_generator_reduce(c, t);
    grid_sync(c);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 7) tmem_dealloc(c.tmem_base, 32);
}

}  // namespace ot

using namespace ot;

extern "C" int ot_decoder_plan_size(void) { return static_cast<int>(sizeof(MkPlan)); }

// layer_ptrs: n_layers x 28 device pointers in the order
//   ln1_g ln1_b ln2_g ln2_b ln3_g ln3_b | qkv_w qkv_sw qkv_b | o_w o_sw o_b | cq_w cq_sw cq_b | co_w co_sw co_b |
//   w1_w w1_sw w1_b | w2_w w2_sw w2_b | kc vc skc svc
// ws_ptrs: x xq sx acc cq cs hq sh rowmax ckv sckv mask fin_g fin_b hout gen_w gen_b gen_pv gen_pi tgt_lut pe ys bar trace
extern "C" int ot_decoder_plan_build(void* plan_dev, int n_layers, int B, int S, int cap, int vocab, int64_t ys_ld,
                                     const void* const* layer_ptrs, const void* const* ws_ptrs) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(plan_dev && layer_ptrs && ws_ptrs, "null argument");
  OT_REQUIRE(n_layers >= 1 && n_layers <= kMkMaxLayers, "1..8 decoder layers");
  OT_REQUIRE(B >= 1 && B <= kMkRows, "the persistent decoder handles 1..64 sentences per launch");
  OT_REQUIRE(S >= 1 && S <= 32 * kDecKeysPerLane && cap >= 2 && cap <= 32 * kDecKeysPerLane, "source length and cache capacity must be <= 96");
  OT_REQUIRE(vocab > 1, "bad vocab");
  MkPlan plan;
  memset(&plan, 0, sizeof(plan));
  MkHot& h = plan.hot;
  h.n_layers = n_layers; h.B = B; h.S = S; h.cap = cap; h.vocab = vocab;
  h.emb_scale = sqrtf(static_cast<float>(kD));
  int i = 0;
  auto nxt = [&]() { return const_cast<void*>(ws_ptrs[i++]); };
  h.x = static_cast<float*>(nxt()); h.xq = static_cast<int8_t*>(nxt()); h.sx = static_cast<float*>(nxt());
  h.acc = static_cast<int32_t*>(nxt()); h.cq = static_cast<int8_t*>(nxt()); h.cs = static_cast<float*>(nxt());
  h.hq = static_cast<int8_t*>(nxt()); h.sh = static_cast<float*>(nxt()); h.rowmax = static_cast<unsigned int*>(nxt());
  h.ckv = static_cast<const int8_t*>(nxt()); h.sckv = static_cast<const float*>(nxt()); h.mask = static_cast<const uint8_t*>(nxt());
  h.fin_g = static_cast<const float*>(nxt()); h.fin_b = static_cast<const float*>(nxt()); h.hout = static_cast<float*>(nxt());
  h.gen_w = static_cast<const float*>(nxt()); h.gen_b = static_cast<const float*>(nxt());
  h.gen_pv = static_cast<float*>(nxt()); h.gen_pi = static_cast<int*>(nxt());
  h.tgt_lut = static_cast<const float*>(nxt()); h.pe = static_cast<const float*>(nxt());
  h.ys = static_cast<int64_t*>(nxt()); h.bar = static_cast<unsigned int*>(nxt()); h.trace = static_cast<unsigned long long*>(nxt());
  h.ys_ld = ys_ld;
  for (int k = 0; k < 23; ++k) OT_REQUIRE(ws_ptrs[k] != nullptr || k == 16, "null workspace pointer");
  int rc;
  if ((rc = get_tensor_map(&plan.map_xq, h.xq, B, kD, kD, kMkRows, 128, true))) return rc;
  if ((rc = get_tensor_map(&plan.map_cq, h.cq, B, kD, kD, kMkRows, 128, true))) return rc;
  if ((rc = get_tensor_map(&plan.map_hq, h.hq, B, kFF, kFF, kMkRows, 128, true))) return rc;
  for (int l = 0; l < n_layers; ++l) {
    const void* const* p = layer_ptrs + l * 28;
    for (int k = 0; k < 28; ++k) OT_REQUIRE(p[k] != nullptr, "null layer pointer");
    MkLayer& L = h.layer[l];
    auto f = [&](int k) { return static_cast<const float*>(p[k]); };
    L.ln1_g = f(0); L.ln1_b = f(1); L.ln2_g = f(2); L.ln2_b = f(3); L.ln3_g = f(4); L.ln3_b = f(5);
    L.qkv_sw = f(7); L.qkv_b = f(8); L.o_sw = f(10); L.o_b = f(11); L.cq_sw = f(13); L.cq_b = f(14);
    L.co_sw = f(16); L.co_b = f(17); L.w1_sw = f(19); L.w1_b = f(20); L.w2_sw = f(22); L.w2_b = f(23);
    L.kc = static_cast<int8_t*>(const_cast<void*>(p[24])); L.vc = static_cast<int8_t*>(const_cast<void*>(p[25]));
    L.skc = static_cast<float*>(const_cast<void*>(p[26])); L.svc = static_cast<float*>(const_cast<void*>(p[27]));
    const int wn[6] = {3 * kD, kD, kD, kD, kFF, kD};
    const int wk[6] = {kD, kD, kD, kD, kD, kFF};
    for (int w = 0; w < 6; ++w)
      if ((rc = get_tensor_map(&plan.map_w[l][w], p[6 + 3 * w], wn[w], wk[w], wk[w], kMkBN, 128, true))) return rc;
  }
  OT_CHECK_CUDA(cudaMemcpy(plan_dev, &plan, sizeof(plan), cudaMemcpyHostToDevice));
  return OT_OK;
}

// Runs greedy steps t0 .. t0+n_steps-1 (ys[:, t0] must hold the current tokens; caches hold positions < t0).
extern "C" int ot_decoder_run(const void* plan_dev, unsigned int* bar_dev, int t0, int n_steps, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(plan_dev && bar_dev && t0 >= 0 && n_steps >= 0, "bad arguments");
  if (n_steps == 0) return OT_OK;
  static int grid = 0;
  if (grid == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    OT_CHECK_CUDA(cudaGetDevice(&dev));
    OT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    OT_CHECK_CUDA(cudaFuncSetAttribute(decoder_steps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal + 1024));
    OT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decoder_steps_kernel, kMkThreads, kSmemTotal + 1024));
    OT_REQUIRE(per_sm >= 1 && sms >= kFF / kMkBN, "the persistent decoder needs >= 128 co-resident CTAs");
    grid = sms;
  }
  cudaStream_t s = as_stream(stream);
  OT_CHECK_CUDA(cudaMemsetAsync(bar_dev, 0, sizeof(unsigned int), s));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kMkThreads);
  cfg.dynamicSmemBytes = kSmemTotal + 1024;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // co-residency of all CTAs is validated by the runtime
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const MkPlan* plan = static_cast<const MkPlan*>(plan_dev);
  OT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decoder_steps_kernel, plan, t0, n_steps));
  count_launch();
  return OT_OK;
}
