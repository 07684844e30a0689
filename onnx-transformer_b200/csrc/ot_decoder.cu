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
#include <stdlib.h>
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
constexpr int kGenKC = 32;       // generator K chunk
constexpr int kGenStages = 4;
constexpr int kGenStageFloats = kGenKC * (kMkRows + kGenVT);   // h chunk [32][64] + W chunk [32][32] = 12 KB
constexpr int kKsPitch = kD + 16;   // padded K row pitch in shared memory: conflict-free 128-bit reads across keys
constexpr int kMaxKeys = 32 * kDecKeysPerLane;   // 96

// shared-memory map (dynamic, base aligned to 1024)
constexpr int kSmemA = 0;                                    // [kMkKB][64 x 128 B] activation operand (TMA, after the barrier)
constexpr int kSmemW = kMkKB * kMkRows * 128;                // [kMkKB][16 x 128 B] weight operand (TMA, prefetched before the barrier;
                                                             //  also the never-read-back rows 64..127 of the last A block)
constexpr int kSmemConst = kSmemW + kMkKB * kMkBN * 128;     // 40960: staged per-column constants of the next row phase (16 KB)
constexpr int kSmemVh = kSmemConst + 16 * 1024;              // 57344: attention V slices [8][96][64] int8 (48 KB) | generator ring
constexpr int kSmemKs = kSmemVh + kHeads * kMaxKeys * kDk;   // 106496: attention K rows [96][528] int8
constexpr int kSmemRow = kSmemKs + kMaxKeys * kKsPitch;      // 157184: this sentence's quantized q | k | v row (1536 B)
constexpr int kSmemCtx = kSmemRow + 1536;                    // 158720: merged context row, 512 floats
constexpr int kSmemRed = kSmemCtx + 2048;                    // 160768: 64 floats of reduction scratch
constexpr int kSmemBars = kSmemRed + 256;                    // 161024: mbarriers + TMEM slot
constexpr int kSmemHot = kSmemBars + 128;                    // MkHot copy (<= 8 KB)
constexpr int kSmemTotal = kSmemHot + 8 * 1024;              // ~166 KB: one CTA per SM
static_assert(kGenStages * kGenStageFloats * 4 <= kHeads * kMaxKeys * kDk, "generator ring must fit the V-slice region");
static_assert(kSmemHot % 16 == 0 && kSmemConst % 16 == 0 && kSmemKs % 16 == 0, "alignment");

struct MkLayer {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
  const float *qkv_sw, *qkv_b, *o_sw, *o_b, *cq_sw, *cq_b, *co_sw, *co_b, *w1_sw, *w1_b, *w2_sw, *w2_b;
  int8_t *kc, *vc;     // self-attention KV cache [B, cap, 512]
  float *skc, *svc;    // [B, cap]
};

// Pointers and sizes: copied to shared memory at kernel start (the grid barrier's fences invalidate L1, and a phase must not
// begin with a dependent L2 round trip just to learn where its operands are).
struct alignas(16) MkHot {
  MkLayer layer[kMkMaxLayers];
  int n_layers, B, S, cap, vocab, n_gen_tiles;
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
  float* houtT;          // [512][64]: final-norm output, k-major (generator operand)
  const float *gen_wt, *gen_b;                 // generator weight, tile-major / k-major: [n_gen_tiles][512][32]
  float* gen_pv; int* gen_pi;                  // per generator tile, per row: best logit / its index
  const float *tgt_lut, *pe;
  int64_t* ys; int64_t ys_ld;
  unsigned int* bar;     // grid barrier counter (zeroed by the host before every launch)
  unsigned long long* trace;   // optional [256]: %globaltimer of CTA 0 at entry / exit of every barrier of the last step of a launch
  long long debug_extra_syncs; // profiling aid (OT_DECODER_EXTRA_SYNCS): empty grid barriers appended to every step
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
  uint64_t* bars;       // abar[kMkKB], wbar, tmem_full
  uint32_t tmem_base;
  unsigned int bar_target;
  uint32_t parity;      // of abar / tmem_full (flips per executed tile)
  uint32_t wparity;     // of wbar (flips per consumed weight prefetch)
  int trace_slot;
  bool trace_on;
};
constexpr int kBarW = kMkKB, kBarFull = kMkKB + 1;

// Grid-wide barrier: every thread's earlier global writes (generic proxy) are visible to every thread's later reads, through
// the generic proxy and through TMA.  Bounded spin: a protocol bug traps instead of hanging the box.
__device__ __forceinline__ void grid_sync(MkCtx& c) {
  asm volatile("fence.proxy.async;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  c.bar_target += gridDim.x;
  if (threadIdx.x == 0) {
    if (c.trace_on) c.P->trace[2 * c.trace_slot] = tl_now();
    // release (cumulative over the CTA's writes ordered by the bar.sync above) + acquire: measured 1.4 us per barrier at 148
    // CTAs vs 1.9 us with explicit __threadfence() pairs (tools/bench_barrier.cu)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(c.P->bar) : "memory");
    unsigned int v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c.P->bar) : "memory");
      if (++spins > (1u << 24)) __trap();
    } while (static_cast<int>(v - c.bar_target) < 0);
    asm volatile("fence.proxy.async;" ::: "memory");
    if (c.trace_on) c.P->trace[2 * c.trace_slot + 1] = tl_now();
  }
  ++c.trace_slot;
  __syncthreads();
  tc_fence_after();
}

// Copy `nfloats` (multiple of 4) constants into the staging region at float offset `dst_f` (all threads, cp.async; the caller
// commits the group).  Constants never change, so this runs BEFORE the barrier that precedes their use.
__device__ __forceinline__ void stage_floats(MkCtx& c, int dst_f, const float* src, int nfloats) {
  float* dst = reinterpret_cast<float*>(c.smem + kSmemConst) + dst_f;
  for (int i = threadIdx.x * 4; i < nfloats; i += blockDim.x * 4) cp_async16(smem_u32(dst + i), src + i);
}
// Staging layout (floats): [0,1536) sw | [1536,3072) bias of the preceding GEMM; [3072,3584) gamma | [3584,4096) beta.
// One call site: a LayerNorm phase needs sw/bias[512] + gamma/beta, an attention phase only sw/bias[n].
constexpr int kCstBias = 3 * kD, kCstGamma = 6 * kD, kCstBeta = 7 * kD;
__device__ __forceinline__ void stage_consts(MkCtx& c, const float* sw, const float* bias, int n, const float* gamma, const float* beta) {
  if (static_cast<int>(blockIdx.x) < c.P->B) {
    __syncthreads();                  // everyone is done with the previous contents of the staging region
    if (sw) { stage_floats(c, 0, sw, n); stage_floats(c, kCstBias, bias, n); }
    if (gamma) { stage_floats(c, kCstGamma, gamma, kD); stage_floats(c, kCstBeta, beta, kD); }
    cp_async_commit();
  }
}

// ------------------------------------------------------------------------------------------------ GEMM tiles
struct TileRef {           // one 64 x 16 x 512 tile of a GEMM: which weight map, which slice
  const CUtensorMap* wmap;
  int k0, n0;
  bool active;
};
__device__ __forceinline__ TileRef tile_of(const CUtensorMap* wmap, int N, int ksplit) {
  const int n_tiles = N / kMkBN;
  const int tile = blockIdx.x;
  TileRef t;
  t.wmap = wmap;
  t.active = tile < n_tiles * ksplit;
  t.n0 = (tile % n_tiles) * kMkBN;
  t.k0 = (tile / n_tiles) * kMkKB * 128;
  return t;
}
// one elected thread of warp 7: weights are constants, so their tile is fetched while the grid is still in earlier phases
__device__ __forceinline__ void issue_w_prefetch(MkCtx& c, const TileRef& t) {
  const uint32_t wb = smem_u32(&c.bars[kBarW]);
  mbar_arrive_expect_tx(wb, kMkKB * kMkBN * 128);
  for (int kb = 0; kb < kMkKB; ++kb) tma_load_2d(smem_u32(c.smem + kSmemW + kb * kMkBN * 128), t.wmap, wb, t.k0 + kb * 128, t.n0);
}

// Accumulate A[0:64, k0:k0+512] * W_tile^T into TMEM columns [0,16); the weight tile was prefetched.  Warp 7 (one elected
// lane) issues the activation loads and the MMAs and, once they have completed, the prefetch of this CTA's NEXT weight tile;
// on return warps 0 and 1 (thread = row) hold the 16 int32 accumulators of their row in r[].
__device__ __forceinline__ void gemm_tile(MkCtx& c, const CUtensorMap* amap, const TileRef& cur, const TileRef& next, uint32_t (&r)[16]) {
  const int warp = threadIdx.x >> 5;
  uint8_t* sA = c.smem + kSmemA;
  uint8_t* sW = c.smem + kSmemW;
  if (warp == 7) {
    if (elect_one()) {
      for (int kb = 0; kb < kMkKB; ++kb) {
        const uint32_t fb = smem_u32(&c.bars[kb]);
        mbar_arrive_expect_tx(fb, kMkRows * 128);
        tma_load_2d(smem_u32(sA + kb * kMkRows * 128), amap, fb, cur.k0 + kb * 128, 0);
      }
      constexpr uint32_t idesc = make_idesc_i8(128, kMkBN);
      mbar_wait(smem_u32(&c.bars[kBarW]), c.wparity);
      for (int kb = 0; kb < kMkKB; ++kb) {
        mbar_wait(smem_u32(&c.bars[kb]), c.parity);
        tc_fence_after();
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sA + kb * kMkRows * 128));
        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(sW + kb * kMkBN * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_i8_ss(c.tmem_base, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
      }
      mma_commit(smem_u32(&c.bars[kBarFull]));
      if (next.active) {
        mbar_wait(smem_u32(&c.bars[kBarFull]), c.parity);   // the MMAs have read the weight tile: its buffer is free again
        issue_w_prefetch(c, next);
      }
    }
    __syncwarp();
  } else if (warp < 2) {
    mbar_wait(smem_u32(&c.bars[kBarFull]), c.parity);
    tc_fence_after();
    tmem_ld_32x16(c.tmem_base + (static_cast<uint32_t>(warp * 32) << 16), r);
    tmem_wait_ld();
  }
  c.parity ^= 1u;
  c.wparity ^= 1u;
}

// Plain GEMM phase: raw int32 partials -> acc[ks][row][N].
__device__ __forceinline__ void phase_gemm(MkCtx& c, const CUtensorMap* amap, const TileRef& cur, int N, const TileRef& next) {
  const MkHot& P = *c.P;
  const int warp = threadIdx.x >> 5;
  if (cur.active) {
    uint32_t r[16];
    gemm_tile(c, amap, cur, next, r);
    const int row = threadIdx.x;
    if (warp < 2 && row < P.B) {
      int4* dst = reinterpret_cast<int4*>(P.acc + (static_cast<int64_t>(cur.k0 / (kMkKB * 128)) * kMkRows + row) * N + cur.n0);
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = make_int4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
  } else if (next.active) {            // idle in this GEMM, busy in the next one: its weight buffer is free
    if (warp == 7) {
      if (elect_one()) issue_w_prefetch(c, next);
      __syncwarp();
    }
  }
}

// FFN1: h = RowQuant_2048(ReLU(x_hat W1^T + b1)) written as the int8 operand of FFN2.  Contains one grid barrier.
__device__ __forceinline__ void phase_ffn1(MkCtx& c, int l, const TileRef& cur, const TileRef& next) {
  const MkHot& P = *c.P;
  const MkLayer& L = P.layer[l];
  const int warp = threadIdx.x >> 5;
  const int row = threadIdx.x;
  const bool owner = cur.active && warp < 2 && row < P.B;
  float y[16];
  if (cur.active) {
    float sw[16], bb[16];
    float sxr = 0.f;
    if (owner) {      // per-column constants and the row scale: in flight while the operands arrive
      sxr = __ldcg(P.sx + row);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(L.w1_sw + cur.n0) + j);
        const float4 b = __ldg(reinterpret_cast<const float4*>(L.w1_b + cur.n0) + j);
        sw[4 * j] = a.x; sw[4 * j + 1] = a.y; sw[4 * j + 2] = a.z; sw[4 * j + 3] = a.w;
        bb[4 * j] = b.x; bb[4 * j + 1] = b.y; bb[4 * j + 2] = b.z; bb[4 * j + 3] = b.w;
      }
    }
    uint32_t r[16];
    gemm_tile(c, &c.G->map_xq, cur, next, r);
    if (owner) {
      float amax = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(static_cast<int>(r[j])), sxr), sw[j]), bb[j]);
        y[j] = fmaxf(v, 0.0f);
        amax = fmaxf(amax, fabsf(y[j]));
      }
      atomicMax(P.rowmax + l * kMkRows + row, __float_as_uint(amax));
    }
  } else if (next.active) {
    if (warp == 7) {
      if (elect_one()) issue_w_prefetch(c, next);
      __syncwarp();
    }
  }
  grid_sync(c);
  if (owner) {
    const float s = quant_scale(__uint_as_float(__ldcg(P.rowmax + l * kMkRows + row)));
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = pack4(quant_one(y[4 * j], s), quant_one(y[4 * j + 1], s), quant_one(y[4 * j + 2], s), quant_one(y[4 * j + 3], s));
    *reinterpret_cast<uint4*>(P.hq + static_cast<int64_t>(row) * kFF + cur.n0) = make_uint4(w[0], w[1], w[2], w[3]);
    if (cur.n0 == 0) P.sh[row] = s;
  }
}

// ------------------------------------------------------------------------------------------------ generator reduce
// First arg-max of sentence b over the per-tile maxima of the previous generator phase (one warp).
__device__ __forceinline__ int generator_reduce_row(const MkHot& P, int b, int lane) {
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  float pv[5];
  int pi[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {            // up to 160 tiles: all loads in flight together
    const int tile = lane + 32 * i;
    const bool ok = tile < P.n_gen_tiles;
    pv[i] = ok ? __ldcg(P.gen_pv + tile * kMkRows + b) : -INFINITY;
    pi[i] = ok ? __ldcg(P.gen_pi + tile * kMkRows + b) : 0x7fffffff;
  }
#pragma unroll
  for (int i = 0; i < 5; ++i)
    if (pv[i] > best || (pv[i] == best && pi[i] < bidx)) { best = pv[i]; bidx = pi[i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
  }
  return (bidx >= 0 && bidx < P.vocab) ? bidx : 0;
}

// ------------------------------------------------------------------------------------------------ LN phase
// SRC 0: x = embedding(token) * sqrt(d) + pe[t]   (embeddings.py:13, positional_encodings.py:24); the token is ys[:, t], or --
//        when the previous step ran in this launch -- the arg-max of its generator tiles, appended to ys here
//        (greedy_decode: parallelized_inject_onnx_transformer.py:753-758)
// SRC 1: x = x + (fl(fl(float(sum_ks acc)*sa[row])*sw[n]) + b[n])     (epilogue + residual of the previous GEMM, split-K 4)
// then LayerNorm; quant: RowQuant -> xq, sx; else y -> houtT (k-major).   Constants come from the staging region.
__device__ __forceinline__ void phase_ln(MkCtx& c, int t, const int SRC, bool quant, const float* a_scale, bool from_tiles) {
  const MkHot& P = *c.P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (SRC == 0 && blockIdx.x == gridDim.x - 1)
    for (int i = threadIdx.x; i < P.n_layers * kMkRows; i += blockDim.x) P.rowmax[i] = 0u;
  const int row = blockIdx.x;
  if (row >= P.B) return;
  cp_async_wait<0>();
  __syncthreads();              // staged constants visible to warp 0
  if (warp != 0) return;
  const float* cst = reinterpret_cast<const float*>(c.smem + kSmemConst);
  float4 v[4];
  float* xr = P.x + static_cast<int64_t>(row) * kD;
  if (SRC == 0) {
    int64_t id;
    if (from_tiles) {
      id = generator_reduce_row(P, row, lane);
      if (lane == 0) P.ys[row * P.ys_ld + t] = id;
    } else {
      id = __ldcg(P.ys + row * P.ys_ld + t);
    }
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
    int4 a[4], p[3][4];
    float4 res[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {          // 20 independent 16-byte loads: one L2 round trip
      a[i] = ldcg_i4(P.acc + static_cast<int64_t>(row) * kD + (i * 32 + lane) * 4);
      res[i] = ldcg_f4(xr + (i * 32 + lane) * 4);
#pragma unroll
      for (int ks = 1; ks < 4; ++ks) p[ks - 1][i] = ldcg_i4(P.acc + (static_cast<int64_t>(ks) * kMkRows + row) * kD + (i * 32 + lane) * 4);
    }
#pragma unroll
    for (int ks = 0; ks < 3; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i].x += p[ks][i].x; a[i].y += p[ks][i].y; a[i].z += p[ks][i].z; a[i].w += p[ks][i].w; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 w4 = reinterpret_cast<const float4*>(cst)[i * 32 + lane];
      const float4 b4 = reinterpret_cast<const float4*>(cst + kCstBias)[i * 32 + lane];
      v[i].x = __fadd_rn(res[i].x, __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a[i].x), sa), w4.x), b4.x));
      v[i].y = __fadd_rn(res[i].y, __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a[i].y), sa), w4.y), b4.y));
      v[i].z = __fadd_rn(res[i].z, __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a[i].z), sa), w4.z), b4.z));
      v[i].w = __fadd_rn(res[i].w, __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a[i].w), sa), w4.w), b4.w));
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(xr)[i * 32 + lane] = v[i];
  const float amax = layernorm_row<4>(v, lane, kD, cst + kCstGamma, cst + kCstBeta, 1e-6f);
  if (quant) {
    const float s = quant_scale(warp_max(amax));
    uint32_t* qr = reinterpret_cast<uint32_t*>(P.xq + static_cast<int64_t>(row) * kD);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      qr[i * 32 + lane] = pack4(quant_one(v[i].x, s), quant_one(v[i].y, s), quant_one(v[i].z, s), quant_one(v[i].w, s));
    if (lane == 0) P.sx[row] = s;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = (i * 32 + lane) * 4;
      P.houtT[(k + 0) * kMkRows + row] = v[i].x;
      P.houtT[(k + 1) * kMkRows + row] = v[i].y;
      P.houtT[(k + 2) * kMkRows + row] = v[i].z;
      P.houtT[(k + 3) * kMkRows + row] = v[i].w;
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention phases
// What an attention phase can fetch before the barrier that precedes it: the old K/V rows of its sentence (into shared
// memory, cp.async) and, per lane, the scales / mask of its keys j = 32*kk + lane.
struct AttnPre {
  float skl[kDecKeysPerLane], svl[kDecKeysPerLane];
  uint8_t keepl[kDecKeysPerLane];
};
__device__ __forceinline__ void attn_prefetch(MkCtx& c, int b, int n_old, const int8_t* k, const int8_t* v, int64_t ldk, int64_t row0,
                                              const float* sk, const float* sv, int64_t sstride, const uint8_t* key_mask, int mask_stride,
                                              AttnPre& pre) {
  int8_t* Ks = reinterpret_cast<int8_t*>(c.smem + kSmemKs);
  int8_t* Vh = reinterpret_cast<int8_t*>(c.smem + kSmemVh);
  for (int idx = threadIdx.x; idx < n_old * 32; idx += blockDim.x) {
    const int j = idx >> 5, ch = idx & 31;
    const int64_t src = (row0 + j) * ldk + ch * 16;
    cp_async16(smem_u32(Ks + j * kKsPitch + ch * 16), k + src);
    cp_async16(smem_u32(Vh + ((ch >> 2) * kMaxKeys + j) * kDk + (ch & 3) * 16), v + src);
  }
  cp_async_commit();
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk) {
    const int j = kk * 32 + lane;
    const bool old = j < n_old;
    const int jc = old ? j : 0;
    pre.skl[kk] = (n_old > 0) ? __ldcg(sk + (row0 + jc) * sstride) : 0.f;
    pre.svl[kk] = (n_old > 0) ? __ldcg(sv + (row0 + jc) * sstride) : 0.f;
    pre.keepl[kk] = (key_mask != nullptr && old) ? key_mask[static_cast<int64_t>(b) * mask_stride + j] : 1;
  }
}

// Epilogue + RowQuant (groups of 512 features) of this sentence's projection row: acc -> int8 row in shared memory.
// NG = 3: fused Q|K|V (no split-K, N = 1536); NG = 1: cross-attention Q (split-K 4, N = 512).  Constants are staged.
__device__ __forceinline__ void project_row(MkCtx& c, const int NG, int b, const float* a_scale, float (&scale)[3]) {
  const MkHot& P = *c.P;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int8_t* rowbuf = reinterpret_cast<int8_t*>(c.smem + kSmemRow);
  float* red = reinterpret_cast<float*>(c.smem + kSmemRed);
  const float* cst = reinterpret_cast<const float*>(c.smem + kSmemConst);
  const int N = NG * kD;
  const float sa = __ldcg(a_scale + b);
  float4 y[2];
  float am[2] = {0.f, 0.f};
  int4 a[2];
  int4 p[3];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = tid + 256 * j;           // float4 index within the row
    a[j] = make_int4(0, 0, 0, 0);
    if (i < N / 4) a[j] = ldcg_i4(P.acc + static_cast<int64_t>(b) * N + i * 4);
  }
  if (NG == 1 && tid < N / 4) {
#pragma unroll
    for (int ks = 1; ks < 4; ++ks) p[ks - 1] = ldcg_i4(P.acc + (static_cast<int64_t>(ks) * kMkRows + b) * N + tid * 4);
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) { a[0].x += p[ks].x; a[0].y += p[ks].y; a[0].z += p[ks].z; a[0].w += p[ks].w; }
  }
  cp_async_wait<0>();
  __syncthreads();                         // staged constants and prefetched K/V rows have landed
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = tid + 256 * j;
    y[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < N / 4) {
      const float4 w4 = reinterpret_cast<const float4*>(cst)[i];
      const float4 b4 = reinterpret_cast<const float4*>(cst + kCstBias)[i];
      y[j].x = __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a[j].x), sa), w4.x), b4.x);
      y[j].y = __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a[j].y), sa), w4.y), b4.y);
      y[j].z = __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a[j].z), sa), w4.z), b4.z);
      y[j].w = __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(a[j].w), sa), w4.w), b4.w);
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

// Decode attention of sentence b with K rows in Ks and V slices in Vh (shared memory): the arithmetic of
// attention_decode_body (ot_attention_decode.cuh), instruction for instruction.  warp h = head h.
__device__ __forceinline__ void attention_smem(MkCtx& c, int b, int Tk, int q_pos0, int mask_kind, float sqi, const AttnPre& pre) {
  const MkHot& P = *c.P;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int8_t* rowbuf = reinterpret_cast<const int8_t*>(c.smem + kSmemRow);
  const int8_t* Ks = reinterpret_cast<const int8_t*>(c.smem + kSmemKs);
  AttnDecVh Vh = reinterpret_cast<AttnDecVh>(c.smem + kSmemVh);
  float* ctx = reinterpret_cast<float*>(c.smem + kSmemCtx);
  uint32_t qw[16];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(rowbuf + h * kDk);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint4 t = qp[w];
      qw[4 * w] = t.x; qw[4 * w + 1] = t.y; qw[4 * w + 2] = t.z; qw[4 * w + 3] = t.w;
    }
  }
  float sc[kDecKeysPerLane], svl[kDecKeysPerLane];
  float mx = -INFINITY;
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk) {
    const int j = kk * 32 + lane;
    const int jc = min(j, Tk - 1);
    const uint4* kp = reinterpret_cast<const uint4*>(Ks + jc * kKsPitch + h * kDk);
    int dot = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint4 t = kp[w];
      dot = __dp4a(static_cast<int>(qw[4 * w]), static_cast<int>(t.x), dot);
      dot = __dp4a(static_cast<int>(qw[4 * w + 1]), static_cast<int>(t.y), dot);
      dot = __dp4a(static_cast<int>(qw[4 * w + 2]), static_cast<int>(t.z), dot);
      dot = __dp4a(static_cast<int>(qw[4 * w + 3]), static_cast<int>(t.w), dot);
    }
    const float s = __fdiv_rn(__fmul_rn(__fmul_rn(__int2float_rn(dot), sqi), pre.skl[kk]), 8.0f);
    const bool visible = pre.keepl[kk] != 0 && (mask_kind != 2 || j <= q_pos0);
    const bool live = j < Tk;
    sc[kk] = live ? (visible ? s : -1e9f) : -INFINITY;
    svl[kk] = live ? pre.svl[kk] : 0.f;
    mx = live ? fmaxf(mx, sc[kk]) : mx;
  }
  mx = warp_max_f(mx);
  float sum = 0.f;
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk) {
    if (kk * 32 + lane < Tk) {
      sc[kk] = expf(__fsub_rn(sc[kk], mx));
      sum += sc[kk];
    }
  }
  sum = warp_sum_f(sum);
  float pq[kDecKeysPerLane];
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk)
    pq[kk] = (kk * 32 + lane < Tk) ? __fdiv_rn(rintf(__fmul_rn(__fdiv_rn(sc[kk], sum), 127.0f)), 127.0f) : 0.f;
  float acc0 = 0.f, acc1 = 0.f;
  const int d0 = 2 * lane;
  // keys in order j = 0..; a key with p = 0 (masked, or beyond Tk where sv = 0 and the V bytes are stale but finite)
  // contributes exactly +-0, as in attention_decode_body
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk) {
    if (kk * 32 >= Tk) break;
#pragma unroll 8
    for (int jj = 0; jj < 32; ++jj) {
      const float ph = __shfl_sync(0xffffffffu, pq[kk], jj);
      const float svj = __shfl_sync(0xffffffffu, svl[kk], jj);
      const char2 vv = *reinterpret_cast<const char2*>(&Vh[h][kk * 32 + jj][d0]);
      acc0 = fmaf(ph, __fmul_rn(__int2float_rn(vv.x), svj), acc0);
      acc1 = fmaf(ph, __fmul_rn(__int2float_rn(vv.y), svj), acc1);
    }
  }
  *reinterpret_cast<float2*>(ctx + h * kDk + d0) = make_float2(acc0, acc1);
  __syncthreads();
  if (h == 0) {
    float4 v[4];
    float amax = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      v[t] = *reinterpret_cast<const float4*>(ctx + (t * 32 + lane) * 4);
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[t].x), fabsf(v[t].y)), fmaxf(fabsf(v[t].z), fabsf(v[t].w))));
    }
    const float s = __fdiv_rn(fmaxf(warp_max_f(amax), 1e-5f), 127.0f);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int q0i = __float2int_rn(rintf(__fdiv_rn(v[t].x, s))), q1i = __float2int_rn(rintf(__fdiv_rn(v[t].y, s)));
      const int q2i = __float2int_rn(rintf(__fdiv_rn(v[t].z, s))), q3i = __float2int_rn(rintf(__fdiv_rn(v[t].w, s)));
      const uint32_t w = (static_cast<uint32_t>(q0i) & 0xFFu) | ((static_cast<uint32_t>(q1i) & 0xFFu) << 8) |
                         ((static_cast<uint32_t>(q2i) & 0xFFu) << 16) | ((static_cast<uint32_t>(q3i) & 0xFFu) << 24);
      *reinterpret_cast<uint32_t*>(P.cq + static_cast<int64_t>(b) * kDm + (t * 32 + lane) * 4) = w;
    }
    if (lane == 0) P.cs[b] = s;
  }
}

// self: Q|K|V epilogue of this sentence's row, KV-cache append, causal attention over t+1 keys;
// cross: Q epilogue, attention over the S cached memory keys with the key-padding mask.
__device__ __forceinline__ void phase_attention(MkCtx& c, bool self, int t, int l, AttnPre& pre) {
  const MkHot& P = *c.P;
  const MkLayer& L = P.layer[l];
  const int b = blockIdx.x;
  if (b >= P.B) return;
  float sc[3];
  project_row(c, self ? 3 : 1, b, P.sx, sc);
  if (self) {
    // this step's K/V row: into the cache (global) and next to the prefetched rows (shared memory)
    const int tid = threadIdx.x;
    const int8_t* rowbuf = reinterpret_cast<const int8_t*>(c.smem + kSmemRow);
    int8_t* Ks = reinterpret_cast<int8_t*>(c.smem + kSmemKs);
    int8_t* Vh = reinterpret_cast<int8_t*>(c.smem + kSmemVh);
    const int64_t dst = (static_cast<int64_t>(b) * P.cap + t) * kD;
    if (tid < 32) {
      const uint4 kk = *reinterpret_cast<const uint4*>(rowbuf + kD + tid * 16);
      *reinterpret_cast<uint4*>(Ks + t * kKsPitch + tid * 16) = kk;
      *reinterpret_cast<uint4*>(L.kc + dst + tid * 16) = kk;
    } else if (tid < 64) {
      const int ch = tid - 32;
      const uint4 vv = *reinterpret_cast<const uint4*>(rowbuf + 2 * kD + ch * 16);
      *reinterpret_cast<uint4*>(Vh + ((ch >> 2) * kMaxKeys + t) * kDk + (ch & 3) * 16) = vv;
      *reinterpret_cast<uint4*>(L.vc + dst + ch * 16) = vv;
    } else if (tid == 64) {
      L.skc[static_cast<int64_t>(b) * P.cap + t] = sc[1];
      L.svc[static_cast<int64_t>(b) * P.cap + t] = sc[2];
    }
    const int lane = tid & 31;
#pragma unroll
    for (int kk = 0; kk < kDecKeysPerLane; ++kk)
      if (kk * 32 + lane == t) { pre.skl[kk] = sc[1]; pre.svl[kk] = sc[2]; }
    __syncthreads();
  }
  attention_smem(c, b, self ? t + 1 : P.S, self ? t : 0, self ? 2 : 1, sc[0], pre);
}

// ------------------------------------------------------------------------------------------------ generator
// logits[r, v] = bias[v] + sum_k h[r,k] * W[v,k]  (k ascending, fmaf: the order of generator_logits_kernel), 64 rows x 32
// vocab entries per tile, thread = 4 rows x 2 vocab entries.  Both operands are k-major (houtT written so by the final norm,
// gen_wt re-laid out once on the host), so a k step is one 128-bit and one 64-bit conflict-free shared load for 8 FMAs; chunks of
// 32 k arrive through a 4-stage cp.async ring whose first weight chunks are fetched before the barrier.
__device__ __forceinline__ void generator_issue(MkCtx& c, int chunk, bool w_part, bool h_part) {
  const MkHot& P = *c.P;
  float* st = reinterpret_cast<float*>(c.smem + kSmemVh) + (chunk % kGenStages) * kGenStageFloats;
  const int tid = threadIdx.x;
  if (h_part) {
    const float* src = P.houtT + static_cast<int64_t>(chunk) * kGenKC * kMkRows;   // 8 KB contiguous
#pragma unroll
    for (int i = 0; i < 2; ++i) cp_async16(smem_u32(st + (tid + 256 * i) * 4), src + (tid + 256 * i) * 4);
  }
  if (w_part) {
    const float* src = P.gen_wt + (static_cast<int64_t>(blockIdx.x) * kD + chunk * kGenKC) * kGenVT;   // 4 KB contiguous
    cp_async16(smem_u32(st + kGenKC * kMkRows + tid * 4), src + tid * 4);
  }
}
__device__ __forceinline__ void generator_prefetch(MkCtx& c) {     // before the barrier: weights only
  if (static_cast<int>(blockIdx.x) < c.P->n_gen_tiles) {
    for (int ch = 0; ch < kGenStages - 1; ++ch) generator_issue(c, ch, true, false);
  }
  cp_async_commit();
}
__device__ __forceinline__ void phase_generator_logits(MkCtx& c) {
  const MkHot& P = *c.P;
  const int tid = threadIdx.x, lane = tid & 31;
  const int tile = blockIdx.x;
  if (tile >= P.n_gen_tiles) { cp_async_wait<0>(); return; }
  constexpr int kChunks = kD / kGenKC;
  for (int ch = 0; ch < kGenStages - 1; ++ch) { generator_issue(c, ch, false, true); cp_async_commit(); }
  const int ty = tid >> 4, tx = tid & 15;      // rows 4ty..4ty+3, vocab entries 2tx, 2tx+1 of the tile
  float acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
  for (int chunk = 0; chunk < kChunks; ++chunk) {
    cp_async_wait<kGenStages - 2>();            // this chunk's operands have landed (groups complete in order)
    __syncthreads();                            // ... for everyone; the stage consumed in the previous iteration is free
    if (chunk + kGenStages - 1 < kChunks) generator_issue(c, chunk + kGenStages - 1, true, true);
    cp_async_commit();                          // (possibly empty: keeps the group count uniform)
    const float* hs = reinterpret_cast<const float*>(c.smem + kSmemVh) + (chunk % kGenStages) * kGenStageFloats;
    const float* ws = hs + kGenKC * kMkRows;
#pragma unroll 8
    for (int k = 0; k < kGenKC; ++k) {
      const float4 h4 = *reinterpret_cast<const float4*>(hs + k * kMkRows + ty * 4);
      const float2 w2 = *reinterpret_cast<const float2*>(ws + k * kGenVT + tx * 2);
      acc[0][0] = fmaf(h4.x, w2.x, acc[0][0]); acc[0][1] = fmaf(h4.x, w2.y, acc[0][1]);
      acc[1][0] = fmaf(h4.y, w2.x, acc[1][0]); acc[1][1] = fmaf(h4.y, w2.y, acc[1][1]);
      acc[2][0] = fmaf(h4.z, w2.x, acc[2][0]); acc[2][1] = fmaf(h4.z, w2.y, acc[2][1]);
      acc[3][0] = fmaf(h4.w, w2.x, acc[3][0]); acc[3][1] = fmaf(h4.w, w2.y, acc[3][1]);
    }
  }
  cp_async_wait<0>();
  const int v0 = tile * kGenVT + tx * 2;
  float bv[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) bv[j] = (v0 + j < P.vocab && P.gen_b) ? __ldg(P.gen_b + v0 + j) : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float best = -INFINITY;
    int bidx = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float lg = __fadd_rn(acc[i][j], bv[j]);
      if (lg != lg) lg = INFINITY;               // torch.max / np.argmax: a NaN logit ranks above every number
      if (v0 + j < P.vocab && (lg > best || (lg == best && v0 + j < bidx))) { best = lg; bidx = v0 + j; }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {            // the 16 lanes that share these rows
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if ((lane & 15) == 0) {
      P.gen_pv[tile * kMkRows + ty * 4 + i] = best;
      P.gen_pi[tile * kMkRows + ty * 4 + i] = bidx;
    }
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
  c.wparity = 0;
  c.trace_slot = 0;
  c.trace_on = false;
  const MkHot& P = *c.P;
  const MkPlan& G = *plan;
  const int warp = threadIdx.x >> 5;

  if (warp == 7) {
    if (elect_one()) {
      for (int i = 0; i <= kBarFull; ++i) mbar_init(smem_u32(&c.bars[i]), 1);
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

  const int nl = P.n_layers;
  const int t_last = t0 + n_steps - 1;
  // tile of this CTA in GEMM w (0 qkv, 1 o, 2 cq, 3 co, 4 w1, 5 w2) of layer l
  auto T = [&](int l, int w) {
    const int N = (w == 0) ? 3 * kD : (w == 4 ? kFF : kD);
    const int ks = (w == 0 || w == 4) ? 1 : 4;
    return tile_of(&G.map_w[l][w], N, ks);
  };

  // before the first barrier: the first weight tile and the constants of the first LayerNorm
  {
    const TileRef first = T(0, 0);
    if (warp == 7 && first.active) {
      if (elect_one()) issue_w_prefetch(c, first);
      __syncwarp();
    }
    stage_consts(c, nullptr, nullptr, 0, P.layer[0].ln1_g, P.layer[0].ln1_b);
  }

  // Phase schedule of a step: 11 phases per layer, then the final norm and the generator.  The loop body holds ONE copy of
  // every phase's code (the kernel must stay within the instruction cache: a fully inlined schedule was 214 KB of SASS and
  // every phase started with instruction-fetch misses).
  //   q: 0 LN1  1 QKV gemm  2 self-attn  3 O gemm  4 LN2  5 CQ gemm  6 cross-attn  7 CO gemm  8 LN3  9 FFN1  10 FFN2 gemm
  const int phases_per_step = nl * 11 + 2;
  AttnPre pre;
#pragma unroll 1
  for (int t = t0; t <= t_last; ++t) {
    c.trace_slot = 0;
    c.trace_on = P.trace != nullptr && blockIdx.x == 0 && t == t_last;
    if (c.trace_on && threadIdx.x == 0) P.trace[255] = tl_now();
#pragma unroll 1
    for (int ph = 0; ph < phases_per_step; ++ph) {
      const int l = min(ph / 11, nl - 1);
      const int q = (ph < nl * 11) ? ph % 11 : 11 + (ph - nl * 11);      // 11: final norm, 12: generator
      const MkLayer& L = P.layer[l];
      // constants to stage after this phase's work (for the next row phase)
      const float *st_sw = nullptr, *st_b = nullptr, *st_g = nullptr, *st_be = nullptr;
      int st_n = 0;
      bool st = false;
      if (q == 0 || q == 4 || q == 8 || q == 11) {
        // ---- LayerNorm phases
        const bool embed = (q == 0 && l == 0);
        const float* a_scale = (q == 0 || q == 11) ? P.sh : P.cs;
        phase_ln(c, t, embed ? 0 : 1, q != 11, a_scale, t > t0);
        st = true;
        if (q == 0) { st_sw = L.qkv_sw; st_b = L.qkv_b; st_n = 3 * kD; }
        else if (q == 4) { st_sw = L.cq_sw; st_b = L.cq_b; st_n = kD; }
        else if (q == 8) {
          st_sw = L.w2_sw; st_b = L.w2_b; st_n = kD;
          st_g = (l + 1 < nl) ? P.layer[l + 1].ln1_g : P.fin_g;
          st_be = (l + 1 < nl) ? P.layer[l + 1].ln1_b : P.fin_b;
        } else {
          generator_prefetch(c);
          st = t < t_last;
          st_g = P.layer[0].ln1_g; st_be = P.layer[0].ln1_b;
        }
      } else if (q == 1 || q == 3 || q == 5 || q == 7 || q == 10) {
        // ---- plain GEMM phases (raw int32 partials); attention operands are prefetched alongside
        if ((q == 1 || q == 5) && static_cast<int>(blockIdx.x) < P.B) {
          const bool self = q == 1;
          const int b = blockIdx.x;
          attn_prefetch(c, b, self ? t : P.S, self ? L.kc : P.ckv + 2 * kD * l, self ? L.vc : P.ckv + 2 * kD * l + kD,
                        self ? kD : 2 * kD * nl, static_cast<int64_t>(b) * (self ? P.cap : P.S), self ? L.skc : P.sckv + 2 * l,
                        self ? L.svc : P.sckv + 2 * l + 1, self ? 1 : 2 * nl, self ? nullptr : P.mask, P.S, pre);
        }
        const int w = (q == 1) ? 0 : (q == 3) ? 1 : (q == 5) ? 2 : (q == 7) ? 3 : 5;
        const CUtensorMap* amap = (q == 1 || q == 5) ? &G.map_xq : (q == 10 ? &G.map_hq : &G.map_cq);
        const bool more = (q != 10) || (l + 1 < nl) || (t < t_last);
        TileRef next = (q == 10) ? T((l + 1) % nl, 0) : T(l, w + 1);
        next.active = next.active && more;
        phase_gemm(c, amap, T(l, w), (w == 0) ? 3 * kD : kD, next);
      } else if (q == 2 || q == 6) {
        // ---- attention phases
        phase_attention(c, q == 2, t, l, pre);
        st = true;
        st_n = kD;
        if (q == 2) { st_sw = L.o_sw; st_b = L.o_b; st_g = L.ln2_g; st_be = L.ln2_b; }
        else { st_sw = L.co_sw; st_b = L.co_b; st_g = L.ln3_g; st_be = L.ln3_b; }
      } else if (q == 9) {
        phase_ffn1(c, l, T(l, 4), T(l, 5));          // one barrier inside
      } else {
        phase_generator_logits(c);
      }
      if (st) stage_consts(c, st_sw, st_b, st_n, st_g, st_be);
      grid_sync(c);
    }
    for (int i = 0; i < static_cast<int>(P.debug_extra_syncs); ++i) grid_sync(c);
  }
  // the last step's token (inside the loop the next step's embedding phase does this)
  if (static_cast<int>(blockIdx.x) < P.B && warp == 0) {
    const int lane = threadIdx.x & 31;
    const int id = generator_reduce_row(P, blockIdx.x, lane);
    if (lane == 0) P.ys[static_cast<int64_t>(blockIdx.x) * P.ys_ld + t_last + 1] = id;
  }

  cp_async_wait<0>();
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
// ws_ptrs: x xq sx acc cq cs hq sh rowmax ckv sckv mask fin_g fin_b houtT gen_wt gen_b gen_pv gen_pi tgt_lut pe ys bar trace
#include <mutex>
#include <unordered_map>
namespace {
std::mutex g_mk_mu;
std::unordered_map<const void*, int> g_mk_caps;      // plan -> KV-cache capacity it was built for (checked by ot_decoder_run)
}  // namespace

extern "C" int ot_decoder_plan_build(void* plan_dev, int n_layers, int B, int S, int cap, int vocab, int64_t ys_ld,
                                     const void* const* layer_ptrs, const void* const* ws_ptrs) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(plan_dev && layer_ptrs && ws_ptrs, "null argument");
  OT_REQUIRE(n_layers >= 1 && n_layers <= kMkMaxLayers, "1..8 decoder layers");
  OT_REQUIRE(B >= 1 && B <= kMkRows, "the persistent decoder handles 1..64 sentences per launch");
  OT_REQUIRE(S >= 1 && S <= 32 * kDecKeysPerLane && cap >= 2 && cap <= 32 * kDecKeysPerLane, "source length and cache capacity must be <= 96");
  OT_REQUIRE(vocab > 1 && (vocab + kGenVT - 1) / kGenVT <= 148, "at most 148 generator tiles of 32 vocabulary entries (one per CTA)");
  MkPlan plan;
  memset(&plan, 0, sizeof(plan));
  MkHot& h = plan.hot;
  h.n_layers = n_layers; h.B = B; h.S = S; h.cap = cap; h.vocab = vocab;
  h.n_gen_tiles = (vocab + kGenVT - 1) / kGenVT;
  h.emb_scale = sqrtf(static_cast<float>(kD));
  int i = 0;
  auto nxt = [&]() { return const_cast<void*>(ws_ptrs[i++]); };
  h.x = static_cast<float*>(nxt()); h.xq = static_cast<int8_t*>(nxt()); h.sx = static_cast<float*>(nxt());
  h.acc = static_cast<int32_t*>(nxt()); h.cq = static_cast<int8_t*>(nxt()); h.cs = static_cast<float*>(nxt());
  h.hq = static_cast<int8_t*>(nxt()); h.sh = static_cast<float*>(nxt()); h.rowmax = static_cast<unsigned int*>(nxt());
  h.ckv = static_cast<const int8_t*>(nxt()); h.sckv = static_cast<const float*>(nxt()); h.mask = static_cast<const uint8_t*>(nxt());
  h.fin_g = static_cast<const float*>(nxt()); h.fin_b = static_cast<const float*>(nxt()); h.houtT = static_cast<float*>(nxt());
  h.gen_wt = static_cast<const float*>(nxt()); h.gen_b = static_cast<const float*>(nxt());
  h.gen_pv = static_cast<float*>(nxt()); h.gen_pi = static_cast<int*>(nxt());
  h.tgt_lut = static_cast<const float*>(nxt()); h.pe = static_cast<const float*>(nxt());
  h.ys = static_cast<int64_t*>(nxt()); h.bar = static_cast<unsigned int*>(nxt()); h.trace = static_cast<unsigned long long*>(nxt());
  h.ys_ld = ys_ld;
  if (const char* e = getenv("OT_DECODER_EXTRA_SYNCS")) h.debug_extra_syncs = atoi(e);
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
  {
    std::lock_guard<std::mutex> lock(g_mk_mu);
    g_mk_caps[plan_dev] = cap;
  }
  return OT_OK;
}

// Runs greedy steps t0 .. t0+n_steps-1 (ys[:, t0] must hold the current tokens; caches hold positions < t0).
extern "C" int ot_decoder_run(const void* plan_dev, unsigned int* bar_dev, int t0, int n_steps, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(plan_dev && bar_dev && t0 >= 0 && n_steps >= 0, "bad arguments");
  {
    std::lock_guard<std::mutex> lock(g_mk_mu);
    auto it = g_mk_caps.find(plan_dev);
    OT_REQUIRE(it != g_mk_caps.end(), "plan_dev was not built by ot_decoder_plan_build in this process");
    OT_REQUIRE(t0 + n_steps <= it->second - 1, "t0 + n_steps exceeds the KV-cache capacity of the plan (cap - 1 greedy steps)");
  }
  if (n_steps == 0) return OT_OK;
  static int grids[64] = {};              // per device: cudaFuncSetAttribute and the SM count are per device
  int dev = 0;
  OT_CHECK_CUDA(cudaGetDevice(&dev));
  OT_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  int& grid = grids[dev];
  if (grid == 0) {
    int sms = 0, per_sm = 0;
    OT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    OT_CHECK_CUDA(cudaFuncSetAttribute(decoder_steps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal + 1024));
    OT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decoder_steps_kernel, kMkThreads, kSmemTotal + 1024));
    OT_REQUIRE(per_sm >= 1 && sms >= 148, "the persistent decoder needs 148 co-resident CTAs (B200)");
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
