// ot_cdecoder_run: the KV-cached greedy decoder as a CLUSTER-RESIDENT persistent kernel (fault-free fast path of
// greedy_decode, parallelized_inject_onnx_transformer.py:616-758 / batch_output.py:659-672).
//
// A greedy step at batch 64 is a chain of ~68 dependent operations of a few hundred KB each.  It is bound by the latency of
// the all-to-all exchange between "column owners" (GEMM tiles need every sentence row) and "row owners" (LayerNorm, RowQuant,
// softmax need every feature of a sentence), not by bandwidth or by the tensor cores.  ot_decoder.cu does that exchange through
// L2 with a grid-wide barrier (~1.3 us + L2 round trips per link).  Here the exchange never leaves the SM-to-SM network:
//
//   * sentences are independent, so the batch is cut into groups of <= 8 sentences and every group is decoded by ONE thread-block
//     cluster of 8 CTAs; clusters never talk to each other (no grid barrier, no co-residency requirement);
//   * CTA r of a cluster OWNS sentence r of the group: its residual row x, its attention (warp = head), its LayerNorm / RowQuant;
//   * every CTA also owns a fixed 1/8 slice of the output features of each of the 6 GEMMs of a layer.  The weight slice streams
//     through a 4 x 32 KB TMA ring (128-byte swizzle) that is refilled right after every GEMM, so each GEMM finds all of its
//     weights resident (weights are constants: the ring runs ahead of the dependency chain).  tcgen05.mma kind::i8 with the 8
//     sentences on the M side (M = 64) and the weight rows on the N side (64..256 per instruction), int32 accumulators in TMEM;
//   * GEMM epilogue -> SCATTER: y[s][f] = fl(fl(float(acc)*sx[s])*sw[f]) + b[f] goes straight into the recv buffer of the CTA
//     that owns sentence s; row phase -> ALL-GATHER: the owner writes its quantized int8 row + scale into the swizzled operand
//     buffer of every CTA of the cluster;
//   * both exchanges are DATA-FLOW synchronised: every remote store is a st.async that also signals complete_tx on an mbarrier
//     of the destination CTA, and a consumer waits only for the bytes it is about to read.  There is no barrier.cluster in the
//     steady state (it cost 0.7 us per exchange, 68 exchanges per greedy step).
//
// Arithmetic is instruction-for-instruction that of the stand-alone kernels (ot_rowmath.cuh, ot_attention_decode.cuh,
// ot_generator.cu), so tokens and KV caches are bit-identical to the per-op engine path (tests/test_decoder_gpu.py).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ot_attention_decode.cuh"
#include "ot_common.h"
#include "ot_ptx.cuh"
#include "ot_rowmath.cuh"

namespace ot {

int get_tensor_map_kblocks(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t K, uint64_t ld, uint32_t box_rows, uint32_t box_kb);

namespace cd {

constexpr int kD = 512;
constexpr int kFF = 2048;
constexpr int kCS = 8;              // CTAs per cluster = max sentences per cluster
constexpr int kThreads = 256;
constexpr int kIssuer = 224;        // warp 7, lane 0: MMA issuer
constexpr int kLoader = 192;        // warp 6, lane 0: TMA producer of the weight ring.  NOT in the issuer's warp: a blocked
                                    // mbarrier.try_wait of one lane stalls the divergent lanes of its warp (measured: 0.5 us per chunk)
constexpr int kSlots = 4;
constexpr int kSlotBytes = 32768;   // up to 256 weight rows x 128 B (one k-block), or 64 rows x 4 k-blocks
constexpr int kChunks = 15;         // weight chunks per layer per CTA
constexpr int kMaxLayers = 8;
constexpr int kMaxKeys = 32 * kDecKeysPerLane;   // 96
constexpr int kGenVT = 32;          // vocabulary entries per generator tile
constexpr int kTmemCols = 512;      // accumulator: lanes 0..7 = sentences, columns = the CTA's output features of the GEMM (<= 256);
                                    // the screening generator double-buffers two 192-column accumulators at columns 0 and 256
constexpr int kGenRows = 576;       // vocabulary rows per CTA in the screening generator: 3 row blocks of 192 (8 x 576 = 4608 >= vocab)
constexpr int kGenRB = 192;
constexpr int kGenChunks = 3 * 8;   // (row block, k-block of 64 fp16) chunks of 192 rows x 128 B per CTA and step

// shared-memory map (dynamic, base aligned to 1024)
constexpr int kSmRing = 0;                                  // weight ring
constexpr int kSmBx = kSmRing + kSlots * kSlotBytes;        // 131072: activation operand, K = 512: [4 k-blocks][8 rows][128 B]
constexpr int kSmBh = kSmBx + 4 * 1024;                     // 135168: activation operand, K = 2048: [16 k-blocks][8 rows][128 B]
constexpr int kSmVs = kSmBh + 16 * 1024;                    // 151552: attention V rows [96][512]; aliased by the generator input [8][512] fp32
constexpr int kSmHb = kSmVs;
constexpr int kSmRecv = kSmVs + kMaxKeys * kD;              // 200704: fp32 row scattered by the GEMM epilogues (<= 2048 floats)
constexpr int kSmX = kSmRecv + 8192;                        // 208896: residual row x of my sentence (512 floats)
constexpr int kSmRow = kSmX + 2048;                         // 210944: int8 staging: q|k|v (1536 B) + ctx (512 B), or my 2 k-blocks of the FFN hidden operand
constexpr int kSmCtx = kSmRow + 2048;                       // 212992: epilogue staging (2 x 1 KB) / final-norm row (512 floats)
constexpr int kSmMisc = kSmCtx + 2048;                      // 215040: scales, reduction scratch, generator partials, hidden-row maxima (2 KB)
constexpr int kSmHot = kSmMisc + 2048;                      // 217088: CdHot copy (<= 1920 B) + 128 B of barriers
constexpr int kSmW2 = kSmHot + 2048;                        // 219136: FFN2 column scales + bias (2 x 512 floats), prefetched for the LayerNorm after it
constexpr int kSmGen = kSmW2 + 4096;                        // 223232: screening generator: logits of row block 2 (6 KB), gathered maxima, 13 mbarriers
constexpr int kSmGenMax = kSmGen + 6144;                    // float [8 ranks][8 sentences] (written by peers)
constexpr int kSmGenBars = kSmGenMax + 256;                 // gfull[4], gempty[4], gaccf[2], gacce[2], gmax
constexpr int kSmTotal = kSmGen + 7168;                     // 230400
static_assert(kSmTotal + 1024 <= 232448, "shared memory budget");
// The M = 64 MMA reads 8 row groups (8 KB) from each operand k-block although only the first group (8 sentences) is real:
// the over-read past Bh must stay inside the CTA's allocation.
static_assert(kSmBh + 16 * 1024 + 7 * 1024 <= kSmTotal, "operand over-read");

// misc region (floats unless noted)
constexpr int kMiSB = 0;        // [8] scale of each sentence's current operand row
constexpr int kMiTok = 12;      // int: the token my sentence feeds this step
constexpr int kMiRed = 16;      // [64] reduction scratch
constexpr int kMiGenV = 80;     // [8 ranks][8 sentences]: best logit of rank's vocabulary slice  (written by peers)
constexpr int kMiGenI = 144;    // [8][8] int: its index
constexpr int kMiPartV = 208;   // [8 warps][8 sentences] per-warp partials (local)
constexpr int kMiPartI = 272;   // [8][8] int -> 336 floats = 1344 B
constexpr int kMiMax = 384;     // [8 sentences][16]: abs-max of the FFN hidden row over each (CTA, epilogue warp)'s 128 columns (written by peers)
static_assert((kMiPartI + 64) <= kMiMax && (kMiMax + 128) * 4 <= 2048, "misc region");
constexpr int kSmBars = kSmHot + 1920;   // 13 mbarriers (104 B) + TMEM slot at +120

struct CdLayer {
  const float *ln_g[3], *ln_b[3];     // ln1, ln2, ln3
  const float *sw[6], *bias[6];       // qkv, o, cq, co, w1, w2
  int8_t *kc, *vc;                    // self-attention KV cache [B, cap, 512]
  float *skc, *svc;                   // [B, cap]
};

struct alignas(16) CdHot {
  CdLayer layer[kMaxLayers];
  int n_layers, B, S, cap, vocab, n_gen_tiles, spc, pad0;
  float emb_scale;
  int pad1;
  const int8_t* ckv; const float* sckv;       // cross K/V projections [B*S, 2*512*n_layers], scales [B*S, 2*n_layers]
  const uint8_t* mask;                        // [B, S]
  const float *fin_g, *fin_b;
  const float *gen_w4, *gen_b;                // generator weight [n_gen_tiles][128][32][4] (tile / k4 / vocab entry / k), bias
  const float *tgt_lut, *pe;
  int64_t* ys; int64_t ys_ld;
  unsigned long long* trace;                  // optional [256]
  float gen_eps;                              // screening generator: max_v ||w_v||_2 * 2^-9 (0 = exact generator only)
  float gen_abs;                              //   + absolute term max_v ||w_v||_2 * 1e-6
  int gen_tc, pad2;
};
static_assert(sizeof(CdHot) % 16 == 0 && sizeof(CdHot) <= 1920, "CdHot is copied to shared memory in 16-byte pieces");

struct CdPlan {
  CUtensorMap map_w[kMaxLayers][6];           // qkv, o, cq, co, w1, w2: (128 B, row, k-block) views, box = one ring chunk
  CUtensorMap map_g;                          // fp16 generator weight [4608][512]: (128 B = 64 k, row, k-block), box = 192 rows x 1 k-block
  CdHot hot;
};

// mbarriers: full[kSlots], empty[kSlots] (weight ring), accfull (MMAs of a GEMM done), kvfull (V rows prefetched),
// gather (operand rows + scales from every owner), scatter (my sentence's row from every CTA), maxima (the 16 partial abs-maxima
// of every sentence's FFN hidden row)
constexpr int kBarEmpty = kSlots, kBarAcc = 2 * kSlots, kBarKv = 2 * kSlots + 1, kBarG = 2 * kSlots + 2, kBarS = 2 * kSlots + 3,
              kBarM = 2 * kSlots + 4;

struct Ctx {
  const CdHot* P;
  const CdPlan* G;
  uint8_t* smem;
  uint64_t* bars;
  uint32_t tmem;
  int rank, n_own, b;    // cluster rank, sentences of this cluster, my sentence (or -1)
  uint32_t pn, cn, total;   // weight chunks issued (loader thread) / consumed (issuer thread) / to do
  uint32_t acc_parity, kv_parity, g_parity, s_parity, m_parity;
  uint32_t gen_n, gen_acc[2], gen_max_parity;   // screening generator: chunks issued / consumed so far (same count in both roles), accumulator uses, exchange parity
  int trace_slot;
  bool trace_on;
  bool fine;                       // intra-phase marks of one layer (profiling aid, trace slots 150..249)
  int mark_slot;
  unsigned long long t_step;
};

// profiling aid, compiled in with -DOT_CD_MARKS (tools/decoder_trace.py --marks): (id << 32 | ns since the step began) into
// trace[150 + n] at points inside the phases of one layer
__device__ __forceinline__ void mark(Ctx& c, int id) {
#ifndef OT_CD_MARKS
  (void)c; (void)id;
  return;
#endif
  if (c.fine && threadIdx.x == 0 && c.mark_slot < 100) {
    c.P->trace[150 + c.mark_slot] = (static_cast<unsigned long long>(id) << 32) | ((tl_now() - c.t_step) & 0xffffffffull);
    ++c.mark_slot;
  }
}
__device__ __forceinline__ float* misc(Ctx& c) { return reinterpret_cast<float*>(c.smem + kSmMisc); }

// remote (or local) shared-memory stores that also signal complete_tx(bytes) on an mbarrier of the destination CTA
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t addr, uint32_t v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(v), "r"(mbar) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit (async proxy), completion signalled on an mbarrier
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Start of a phase: wait until the bytes this phase consumes have landed in this CTA (bar = kBarG: operand rows + scales of an
// all-gather, kBarS: my sentence's row of a scatter).  A phase of the mbarrier = one arrival (thread 0) + the expected byte count.
// Thread 0 arms the NEXT use of the barrier (`next_bytes`, 0 = none) the moment this one completes -- bytes that arrive before the
// arming are accounted for by the signed tx-count, and the wait then ends when the last byte lands, not when thread 0 gets there
// (it reaches the next wait 0.3-0.6 us after the row phase it also works in).  The first use is armed at kernel start.
// A CTA cannot run two phases ahead of its own threads: between a wait and this CTA's contribution to the next exchange on the
// same barrier lies a block barrier of the phase in between.  Buffers are reused every second exchange; that is safe because a CTA
// sends its contribution to exchange k+1 only after it has consumed exchange k AND every exchange is an all-to-all -- except the
// FFN1 -> FFN2 hand-off (pair-wise): a pair may be one GEMM phase ahead of the rest of the cluster, so nothing a CTA still uses in
// its FFN1 tail may live in a buffer that FFN2's epilogue writes remotely (recv).  Also the profiling hook: CTA 0
// stamps the begin / end of every wait.
__device__ __forceinline__ void xwait(Ctx& c, int bar, uint32_t next_bytes, uint32_t& parity, bool waits) {
  mark(c, 1);
  if (c.trace_on && threadIdx.x == 0) c.P->trace[2 * c.trace_slot] = tl_now();
  if (waits) {
    const uint32_t b = smem_u32(&c.bars[bar]);
    mbar_wait(b, parity);
    parity ^= 1u;
    if (threadIdx.x == 0 && next_bytes != 0) mbar_arrive_expect_tx(b, next_bytes);
  }
  if (c.trace_on && threadIdx.x == 0) c.P->trace[2 * c.trace_slot + 1] = tl_now();
  __syncwarp();
  ++c.trace_slot;
  mark(c, 2);
}

// ------------------------------------------------------------------------------------------------ weight ring
// A measured property of tcgen05.mma shapes this layout is built on: an instruction costs about as long as its M-side operand
// takes to read (~1 cycle per row) however small N is, so the many weight rows go on the N side (up to 256 per instruction)
// and the 8 sentences on the M side (M = 64, the smallest).  Chunk = what one ring slot holds:
//   qkv  (192 rows / CTA): 4 chunks = k-blocks, 3 boxes of 64 rows        o, cq, co (64 rows): 1 chunk = 4 k-blocks (4 boxes)
//   ffn1 (256 rows / CTA): 4 chunks = k-blocks                            ffn2: SPLIT-K -- CTA r takes output rows 256*(r&1)..+256
//   and k-blocks 4*(r>>1)..+4 (4 chunks): 16 MMAs of N = 256 instead of 64 of N = 64 (an M = 64 MMA costs ~86 cycles whatever N);
//   the 4 int32 partial planes are summed by the row owner (exact), which also applies the fp32 epilogue
// Every chunk is ONE TMA instruction (a 3-D box over the (128 B, row, k-block) view of the weight): a warp needs ~0.1 us per TMA
// instruction, and the loader's warp is also a worker of the row phase that follows.
struct Chunk { int g, kb0, bytes; };
__device__ __forceinline__ Chunk chunk_of(int j) {
  Chunk k;
  if (j < 4) { k.g = 0; k.kb0 = j; k.bytes = 192 * 128; }
  else if (j < 7) { k.g = j - 3; k.kb0 = 0; k.bytes = 4 * 64 * 128; }
  else if (j < 11) { k.g = 4; k.kb0 = j - 7; k.bytes = 256 * 128; }
  else { k.g = 5; k.kb0 = j - 11; k.bytes = 256 * 128; }
  return k;
}
__device__ __forceinline__ int slice_rows(int g) { return g == 0 ? 192 : (g >= 4 ? 256 : 64); }

// loader WARP (all lanes keep c.pn; one elected lane issues): load weight chunk c.pn into its ring slot (the slot must be free)
__device__ __forceinline__ void issue_chunk(Ctx& c) {
  const uint32_t n = c.pn;
  const int slot = n % kSlots;
  const int l = (n / kChunks) % c.P->n_layers;
  const Chunk k = chunk_of(n % kChunks);
  const CUtensorMap* map = &c.G->map_w[l][k.g];
  const uint32_t fb = smem_u32(&c.bars[slot]);
  const uint32_t dst = smem_u32(c.smem + kSmRing + slot * kSlotBytes);
  if (elect_one()) {
    mbar_arrive_expect_tx(fb, k.bytes);
    if (k.g == 5) tma_load_3d(dst, map, fb, 0, 256 * (c.rank & 1), 4 * (c.rank >> 1) + k.kb0);
    else tma_load_3d(dst, map, fb, 0, slice_rows(k.g) * c.rank, k.kb0);
  }
  __syncwarp();
  ++c.pn;
}
// loader warp: issue every chunk below `upto` (chunk n reuses the slot of chunk n - kSlots, free once that chunk's MMAs have
// completed: the wait below).  Called at the start of every GEMM phase with upto = kSlots chunks past the GEMM's last one: no GEMM
// has more than kSlots chunks, so every GEMM starts with all of its weights in flight or resident.
__device__ __forceinline__ void fill_until(Ctx& c, uint32_t upto) {
  upto = min(upto, c.total);
  while (c.pn < upto) {
    const uint32_t round = c.pn / kSlots;
    if (round > 0) mbar_wait(smem_u32(&c.bars[kBarEmpty + c.pn % kSlots]), (round - 1) & 1);
    issue_chunk(c);
  }
}

// loader warp: generator chunk j of this step (row block j / 8, k-block j % 8) into ring slot (gen_n + j) % kSlots
__device__ __forceinline__ void gen_issue_chunk(Ctx& c, int j) {
  uint64_t* gb = reinterpret_cast<uint64_t*>(c.smem + kSmGenBars);
  const uint32_t G = c.gen_n + j;
  const int slot = G % kSlots;
  if (j < kSlots) {
    // the main ring's last use of this slot (the ring is drained: every chunk of this step has been issued)
    const uint32_t N = c.pn;
    const uint32_t m = N - 1 - ((N - 1 - slot) % kSlots);             // largest main chunk index < N with index % kSlots == slot
    mbar_wait(smem_u32(&c.bars[kBarEmpty + slot]), (m / kSlots) & 1);
  }
  if (G >= kSlots) mbar_wait(smem_u32(&gb[4 + slot]), ((G / kSlots) - 1) & 1);
  if (elect_one()) {
    const uint32_t fb = smem_u32(&gb[slot]);
    mbar_arrive_expect_tx(fb, kGenRB * 128);
    tma_load_3d(smem_u32(c.smem + kSmRing + slot * kSlotBytes), &c.G->map_g, fb, 0, c.rank * kGenRows + (j >> 3) * kGenRB, j & 7);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ GEMM phase
// g: 0 qkv, 1 o, 2 cq, 3 co, 4 w1 (ReLU), 5 w2.  D[s][f] = sum_k a[s][k] * W[f][k] (rows s >= 8 of the M = 64 tile are whatever
// follows the 8 operand rows in shared memory: never read back); then scatter y[s][f] to the owner of sentence s.
__device__ __forceinline__ void phase_gemm(Ctx& c, int l, int g, uint32_t gend, uint32_t next_gather_bytes) {
  const CdHot& P = *c.P;
  const CdLayer& L = P.layer[l];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows = slice_rows(g);
  // epilogue threads: warps 0 and 4 (the two that may read TMEM lanes 0..31), each half of the columns in groups of 32;
  // 16x256b fragment: thread t holds sentence t/4, features 8*i + 2*(t%4) + {0,1} of every 8-column block i
  const bool epi = (warp & 3) == 0;
  const int half = rows >> 1;                       // 96 / 32 / 128 columns per warp
  const int col0 = (warp >> 2) * half;
  const int ngroups = half >> 5;                    // 3 / 1 / 4
  const int srow = lane >> 2;
  float2 sw[16], bb[16];
  if (epi && g != 5) {        // per-feature constants: in flight while the operand rows arrive
    const float* swp = L.sw[g] + rows * c.rank + col0 + 2 * (lane & 3);
    const float* bp = L.bias[g] + rows * c.rank + col0 + 2 * (lane & 3);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < 4 * ngroups) {
        sw[i] = __ldg(reinterpret_cast<const float2*>(swp + 8 * i));
        bb[i] = __ldg(reinterpret_cast<const float2*>(bp + 8 * i));
      }
    }
  }
  // consumes 512 operand bytes + the scale per sentence; FFN2: 8 rows x 256 hidden bytes from me and from my split-K partner
  xwait(c, kBarG, next_gather_bytes, c.g_parity, true);
  // Issuer and loader are WHOLE warps that elect one lane around the tcgen05 / TMA instructions.  With `tid == kIssuer` (a branch
  // ptxas cannot prove warp-uniform) every UTCIMMA was wrapped in an ELECT / 4 x R2UR.BROADCAST / BRA.U.ANY loop: ~100 cycles per
  // instruction whatever its shape (the "floor" of profiles/r1_mma_shape_microbench.txt; tools/bench_mma2.cu, profiles/r2_mma_issue_microbench.txt);
  // in a uniform region the MMAs issue back to back and cost what the tensor pipe needs.  All lanes keep the ring counters.
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  if (warp_u == kIssuer / 32) {
    const uint32_t abase = smem_u32(c.smem + (g == 5 ? kSmBh : kSmBx));
    const uint32_t idesc = make_idesc_i8(64, rows);
    const bool by_rows = (g == 0 || g >= 4);
    const int nchunks = (g >= 1 && g <= 3) ? 1 : 4;
    const int kb_base = (g == 5) ? 4 * (c.rank >> 1) : 0;
    if (elect_one()) fence_proxy_async_smem();      // operand rows were written through the generic proxy (shared memory only: the narrow fence)
    for (int ch = 0; ch < nchunks; ++ch) {
      const int slot = c.cn % kSlots;
      mbar_wait(smem_u32(&c.bars[slot]), (c.cn / kSlots) & 1);
      tc_fence_after();
      const uint32_t sbase = smem_u32(c.smem + kSmRing + slot * kSlotBytes);
      if (elect_one()) {
        if (by_rows) {
          const uint64_t a_desc = make_smem_desc_sw128(abase + (kb_base + ch) * 1024);
          const uint64_t b_desc = make_smem_desc_sw128(sbase);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_i8_ss(c.tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc, (ch | k) != 0 ? 1u : 0u);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {           // one chunk = 4 k-blocks of 64 weight rows
            const uint64_t a_desc = make_smem_desc_sw128(abase + i * 1024);
            const uint64_t b_desc = make_smem_desc_sw128(sbase + i * 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_i8_ss(c.tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc, (i | k) != 0 ? 1u : 0u);
          }
        }
        mma_commit(smem_u32(&c.bars[kBarEmpty + slot]));
      }
      __syncwarp();
      ++c.cn;
    }
    if (elect_one()) mma_commit(smem_u32(&c.bars[kBarAcc]));
    __syncwarp();
  } else if (warp_u == kLoader / 32) {
    // Refill the ring WHILE this GEMM's MMAs run: a slot is reloaded as soon as the MMAs that read it have completed (its `empty`
    // barrier), so all but the last chunk of the refill streams in during the MMA window.  Measured: a 128 KB TMA burst issued
    // after the last MMA stalls every shared-memory load of the SM for ~0.9 us -- exactly when the epilogue and the next row
    // phase need them -- while loads that land during the MMAs cost nothing (the MMA time is unchanged).
    if (c.P->gen_tc && g == 5 && l == c.P->n_layers - 1) {
      // the screening generator borrows the ring's slots: no run-ahead into the next step across it; its own first chunks (constants
      // too) are requested right here, two phases before they are needed
      fill_until(c, gend);
      for (int j = 0; j < kSlots; ++j) gen_issue_chunk(c, j);
    } else {
      fill_until(c, gend + kSlots);
    }
  }
  __syncwarp();
  if (epi) {
    mbar_wait(smem_u32(&c.bars[kBarAcc]), c.acc_parity);
    tc_fence_after();
    mark(c, 20 + g);
    const float sx = (misc(c) + kMiSB)[srow];
    // y of one 32-column group is staged per warp as [sentence][32 features] (1 KB) and leaves as 16-byte chunks, 128 contiguous
    // bytes per destination CTA and instruction
    float* stage = reinterpret_cast<float*>(c.smem + kSmCtx) + (warp >> 2) * 256;
    const uint32_t recv0 = smem_u32(c.smem + kSmRecv) + 4u * static_cast<uint32_t>(rows * c.rank + col0);
    const uint32_t sbar = smem_u32(&c.bars[kBarS]);
    if (g == 4) {
      // FFN1: bias + ReLU, then the RowQuant of the hidden rows right here.  Every (CTA, epilogue warp) sends the abs-max of its 128
      // columns of each sentence to all 8 CTAs (512 bytes per CTA in total); with the 16 partials of a sentence every CTA derives
      // the same scale, quantizes its own columns and writes them -- already in the swizzled operand layout -- into the FFN2
      // operand buffers of the two CTAs whose split-K slice covers them: itself and rank ^ 1.  (Before: fp32 scatter to the
      // sentence's owner, a RowQuant phase there, and an all-gather: one more exchange per layer.)
      // Here: y -> shared memory [sentence][256 columns] (8-float groups XOR-swizzled by the sentence against bank conflicts).
      // (in the V region, idle between the attention phases: NOT in recv -- FFN2's split-K planes are scattered into recv by CTAs
      // that depend only on their own pair's hidden bytes and may get there while this CTA still reads y)
      float* ybuf = reinterpret_cast<float*>(c.smem + kSmVs) + srow * 256;
      float am = 0.f;
#pragma unroll
      for (int gq = 0; gq < 4; ++gq) {
        uint32_t r[16];
        tmem_ld_16x256b_x4(c.tmem + col0 + 32 * gq, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 w2 = sw[4 * gq + i], b2 = bb[4 * gq + i];
          const float y0 = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(static_cast<int>(r[4 * i])), sx), w2.x), b2.x), 0.0f);
          const float y1 = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(static_cast<int>(r[4 * i + 1])), sx), w2.y), b2.y), 0.0f);
          const int col = col0 + 32 * gq + 8 * i;
          *reinterpret_cast<float2*>(ybuf + (col ^ (srow << 3)) + 2 * (lane & 3)) = make_float2(y0, y1);
          am = fmaxf(am, fmaxf(y0, y1));          // y >= 0 (fmaxf(NaN, 0) = 0): the abs-max is the max
        }
      }
      tc_fence_before();
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 1));
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 2));
      const uint32_t mbar = smem_u32(&c.bars[kBarM]);
      const uint32_t slot = smem_u32(misc(c) + kMiMax + srow * 16 + 2 * c.rank + (warp >> 2));
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int peer = 2 * (lane & 3) + j;
        st_async_b32(mapa_shared(slot, peer), __float_as_uint(am), mapa_shared(mbar, peer));
      }
      if (tid == 0) mbar_arrive_expect_tx(mbar, kCS * 16 * 4);
      mark(c, 70);
    } else {
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      if (gq < ngroups) {
        uint32_t r[16];
        tmem_ld_16x256b_x4(c.tmem + col0 + 32 * gq, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (g == 5) {      // split-K partial: raw int32, the owner sums the 4 planes and applies the epilogue
            *reinterpret_cast<uint2*>(stage + srow * 32 + 8 * i + 2 * (lane & 3)) = make_uint2(r[4 * i], r[4 * i + 1]);
          } else {
            const float2 w2 = sw[4 * gq + i], b2 = bb[4 * gq + i];
            float y0 = __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(static_cast<int>(r[4 * i])), sx), w2.x), b2.x);
            float y1 = __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn(static_cast<int>(r[4 * i + 1])), sx), w2.y), b2.y);
            *reinterpret_cast<float2*>(stage + srow * 32 + 8 * i + 2 * (lane & 3)) = make_float2(y0, y1);
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int id = lane + 32 * j, s = id >> 3, ch = id & 7;     // sentence, 16-byte chunk of its 32 features
          if (s < c.n_own) {
            const uint4 v = *reinterpret_cast<const uint4*>(stage + s * 32 + ch * 4);
            st_async_v4(mapa_shared(recv0 + 4u * static_cast<uint32_t>(32 * gq + 4 * ch), s), v, mapa_shared(sbar, s));
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
    }
  }
  if (g == 4) {
    // all 256 threads: 8 columns of one sentence each
    mbar_wait(smem_u32(&c.bars[kBarM]), c.m_parity);
    c.m_parity ^= 1u;
    __syncthreads();                 // y of the two epilogue warps is in shared memory
    mark(c, 71);
    const float4* m4 = reinterpret_cast<const float4*>(misc(c) + kMiMax + warp * 16);
    const float4 ma = m4[0], mb = m4[1], mc = m4[2], md = m4[3];
    const float amax = fmaxf(fmaxf(fmaxf(fmaxf(ma.x, ma.y), fmaxf(ma.z, ma.w)), fmaxf(fmaxf(mb.x, mb.y), fmaxf(mb.z, mb.w))),
                             fmaxf(fmaxf(fmaxf(mc.x, mc.y), fmaxf(mc.z, mc.w)), fmaxf(fmaxf(md.x, md.y), fmaxf(md.z, md.w))));
    const float s = quant_scale_x(amax);
    const float rinv = __frcp_rn(s);
    const float4* y4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(c.smem + kSmVs) + warp * 256 + ((8 * lane) ^ (warp << 3)));
    uint2 w;
    w.x = quant4_pack(y4[0], s, rinv);
    w.y = quant4_pack(y4[1], s, rinv);
    // columns 8*lane .. +7 of sentence `warp`: k-block lane >> 4, 16-byte chunk (lane >> 1) & 7 (swizzled by the row), half lane & 1
    *reinterpret_cast<uint2*>(c.smem + kSmRow + (lane >> 4) * 1024 + warp * 128 + ((((lane >> 1) & 7) ^ warp) << 4) + 8 * (lane & 1)) = w;
    __syncthreads();
    mark(c, 72);
    {
      const int ch = tid & 127, dest = (tid >> 7) ? (c.rank ^ 1) : c.rank;       // 2 KB = 128 chunks, to me and to my split-K partner
      const uint4 v = *reinterpret_cast<const uint4*>(c.smem + kSmRow + ch * 16);
      const uint32_t dst = smem_u32(c.smem + kSmBh + 2 * c.rank * 1024 + ch * 16);
      st_async_v4(mapa_shared(dst, dest), v, mapa_shared(smem_u32(&c.bars[kBarG]), dest));
    }
  }
  mark(c, 30 + g);
  c.acc_parity ^= 1u;
}

// ------------------------------------------------------------------------------------------------ all-gather helpers
// Push the 512-byte int8 row staged at `src` into row `rank` of every CTA's operand buffer (128-byte swizzle: 16-byte chunk c of a
// row lands at chunk c ^ (row & 7) of its 128-byte line; k-blocks 1 KB apart), plus its scale: a peer receives 512 + 4 bytes per
// sentence.
__device__ __forceinline__ void push_row_q8(Ctx& c, const uint8_t* src, float scale) {
  const int tid = threadIdx.x;
  const uint32_t gbar = smem_u32(&c.bars[kBarG]);
  {
    const int peer = tid >> 5, ch = tid & 31;                     // 256 threads = 8 peers x 32 chunks of 16 bytes
    const uint4 v = *reinterpret_cast<const uint4*>(src + ch * 16);
    const int kb = ch >> 3, cc = ch & 7;
    const uint32_t local = smem_u32(c.smem + kSmBx + kb * 1024 + c.rank * 128 + ((cc ^ (c.rank & 7)) << 4));
    st_async_v4(mapa_shared(local, peer), v, mapa_shared(gbar, peer));
  }
  if (tid < kCS) st_async_b32(mapa_shared(smem_u32(misc(c) + kMiSB + c.rank), tid), __float_as_uint(scale), mapa_shared(gbar, tid));
}

// ------------------------------------------------------------------------------------------------ row phases (owner CTA)
__device__ __forceinline__ void bar_sync_128() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// LayerNorm of my sentence.  SRC 0: x = embedding(token)*sqrt(d) + pe[t]; SRC 1: x = x + recv (the scattered O / CO row);
// SRC 2: x = x + (fl(fl(float(p0+p1+p2+p3)*s_h)*s_w[f]) + b[f]) with the 4 split-K int32 planes of FFN2 in recv (s_w / b staged in kSmW2).
// quant: RowQuant -> all-gather into Bx; else (final norm) the fp32 row goes to every CTA's generator input.
// The row is spread over 4 warps -- warp i holds float4 i*32+lane, exactly the element layernorm_row<4> gives lane `lane` in its
// i-th register -- and every reduction is evaluated in layernorm_row's order: per lane ((p0 + p1) + p2) + p3 over the four
// registers, then the xor-shuffle tree over lanes.  Same instructions on the same operands => bit-identical results, at a quarter
// of the dependent-division chain (a single warp spends 2.2 us per row in IEEE divisions).
__device__ __forceinline__ void phase_ln(Ctx& c, int SRC, int64_t token, int t, const float* gamma, const float* beta, bool quant) {
  const CdHot& P = *c.P;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* xr = reinterpret_cast<float*>(c.smem + kSmX);
  uint8_t* rowq = c.smem + kSmRow;
  float* yrow = reinterpret_cast<float*>(c.smem + kSmCtx);
  float* red = misc(c) + kMiRed;
  float* part = reinterpret_cast<float*>(c.smem + kSmRow + 1024);    // [2][4][32] partial sums (the staging row only uses its first 512 B here)
  if (c.b >= 0 && warp < 4) {
    const int i4 = warp * 32 + lane;
    float4 v;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i4);
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + i4);
    if (SRC == 0) {
      const float4 e = __ldg(reinterpret_cast<const float4*>(P.tgt_lut + token * kD) + i4);
      const float4 q = __ldg(reinterpret_cast<const float4*>(P.pe + static_cast<int64_t>(t) * kD) + i4);
      v = make_float4(__fadd_rn(__fmul_rn(e.x, P.emb_scale), q.x), __fadd_rn(__fmul_rn(e.y, P.emb_scale), q.y),
                      __fadd_rn(__fmul_rn(e.z, P.emb_scale), q.z), __fadd_rn(__fmul_rn(e.w, P.emb_scale), q.w));
    } else if (SRC == 1) {
      const float4 res = reinterpret_cast<const float4*>(xr)[i4], y = reinterpret_cast<const float4*>(c.smem + kSmRecv)[i4];
      v = make_float4(__fadd_rn(res.x, y.x), __fadd_rn(res.y, y.y), __fadd_rn(res.z, y.z), __fadd_rn(res.w, y.w));
    } else {
      // FFN2's column scales / bias: copied to shared memory by this very thread during the FFN2 GEMM phase (an L2 round trip
      // -- 0.35 us -- would otherwise sit in front of the first reduction)
      asm volatile("cp.async.wait_all;" ::: "memory");
      const float4 w4 = reinterpret_cast<const float4*>(c.smem + kSmW2)[i4], b4 = reinterpret_cast<const float4*>(c.smem + kSmW2)[128 + i4];
      const int4* pl = reinterpret_cast<const int4*>(c.smem + kSmRecv);
      const int4 p0 = pl[i4], p1 = pl[128 + i4], p2 = pl[256 + i4], p3 = pl[384 + i4];
      // scale of my sentence's quantized hidden row: from its 16 partial maxima (exchanged during FFN1; every CTA derives the same value)
      mbar_wait(smem_u32(&c.bars[kBarM]), c.m_parity ^ 1u);
      const float4* m4 = reinterpret_cast<const float4*>(misc(c) + kMiMax + c.rank * 16);
      const float4 ma = m4[0], mb = m4[1], mc = m4[2], md = m4[3];
      const float sh = quant_scale_x(fmaxf(fmaxf(fmaxf(fmaxf(ma.x, ma.y), fmaxf(ma.z, ma.w)), fmaxf(fmaxf(mb.x, mb.y), fmaxf(mb.z, mb.w))),
                                            fmaxf(fmaxf(fmaxf(mc.x, mc.y), fmaxf(mc.z, mc.w)), fmaxf(fmaxf(md.x, md.y), fmaxf(md.z, md.w)))));
      const float4 res = reinterpret_cast<const float4*>(xr)[i4];
      v.x = __fadd_rn(res.x, __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn((p0.x + p1.x) + (p2.x + p3.x)), sh), w4.x), b4.x));
      v.y = __fadd_rn(res.y, __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn((p0.y + p1.y) + (p2.y + p3.y)), sh), w4.y), b4.y));
      v.z = __fadd_rn(res.z, __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn((p0.z + p1.z) + (p2.z + p3.z)), sh), w4.z), b4.z));
      v.w = __fadd_rn(res.w, __fadd_rn(__fmul_rn(__fmul_rn(__int2float_rn((p0.w + p1.w) + (p2.w + p3.w)), sh), w4.w), b4.w));
    }
    reinterpret_cast<float4*>(xr)[i4] = v;
    mark(c, 41);
    const float nf = static_cast<float>(kD);
    part[warp * 32 + lane] = (v.x + v.y) + (v.z + v.w);
    bar_sync_128();
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) sum += part[i * 32 + lane];
    const float mu = __fmul_rn(warp_sum(sum), 1.0f / kD);       // division by 512 = multiplication by 2^-9, bit for bit (also for subnormals)
    v.x = __fsub_rn(v.x, mu); v.y = __fsub_rn(v.y, mu); v.z = __fsub_rn(v.z, mu); v.w = __fsub_rn(v.w, mu);
    part[128 + warp * 32 + lane] = (__fmul_rn(v.x, v.x) + __fmul_rn(v.y, v.y)) + (__fmul_rn(v.z, v.z) + __fmul_rn(v.w, v.w));
    bar_sync_128();
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) sq += part[128 + i * 32 + lane];
    float var = __fmul_rn(warp_sum(sq), 1.0f / kD);
    var = div511_exact(__fmul_rn(var, nf));                     // * N / (N-1)
    const float denom = __fadd_rn(__fsqrt_rn(var), 1e-6f);
    v.x = __fadd_rn(__fdiv_rn(__fmul_rn(g.x, v.x), denom), be.x);
    v.y = __fadd_rn(__fdiv_rn(__fmul_rn(g.y, v.y), denom), be.y);
    v.z = __fadd_rn(__fdiv_rn(__fmul_rn(g.z, v.z), denom), be.z);
    v.w = __fadd_rn(__fdiv_rn(__fmul_rn(g.w, v.w), denom), be.w);
    mark(c, 42);
    if (quant) {
      const float am = warp_max_nonneg(fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
      if (lane == 0) red[32 + warp] = am;
      bar_sync_128();
      const float s = quant_scale_x(fmaxf(fmaxf(red[32], red[33]), fmaxf(red[34], red[35])));
      reinterpret_cast<uint32_t*>(rowq)[i4] = quant4_pack(v, s, __frcp_rn(s));
      if (tid == 0) red[0] = s;
    } else {
      reinterpret_cast<float4*>(yrow)[i4] = v;
    }
  }
  mark(c, 43);
  __syncthreads();
  mark(c, 44);
  if (c.b >= 0) {
    if (quant) {
      push_row_q8(c, rowq, red[0]);
    } else {
      const uint32_t gbar = smem_u32(&c.bars[kBarG]);
      for (int idx = tid; idx < 128 * kCS; idx += kThreads) {
        const int peer = idx >> 7, ch = idx & 127;
        const uint4 v = *reinterpret_cast<const uint4*>(yrow + ch * 4);
        st_async_v4(mapa_shared(smem_u32(c.smem + kSmHb + (c.rank * kD + ch * 4) * 4), peer), v, mapa_shared(gbar, peer));
      }
    }
  }
  mark(c, 45);
}

// ------------------------------------------------------------------------------------------------ attention (owner CTA)
struct AttnPre {
  float skl[kDecKeysPerLane], svl[kDecKeysPerLane];
  uint8_t keepl[kDecKeysPerLane];
};
// Issued during the preceding GEMM phase: the old V rows of my sentence into shared memory (one 512-byte bulk copy per row; the
// copies run in the TMA unit) and, per lane, the scales / mask of its keys j = 32*kk + lane.  Issued by warps 1, 2, 3 and 5: idle
// during a GEMM phase (0 / 4: epilogue, 6: weight ring, 7: MMA issuer); a warp needs ~20 ns per bulk-copy instruction.
__device__ __forceinline__ void attn_prefetch(Ctx& c, int n_old, const int8_t* k, const int8_t* v, int64_t ldk, int64_t row0, const float* sk, const float* sv,
                                              int64_t sstride, const uint8_t* key_mask, int mask_stride, AttnPre& pre) {
  const uint32_t bar = smem_u32(&c.bars[kBarKv]);
  const int w = threadIdx.x >> 5;
  if (threadIdx.x == 32) {
    fence_proxy_async_smem();     // the V region was last touched through the generic proxy
    mbar_arrive_expect_tx(bar, static_cast<uint32_t>(n_old) * kD);
  }
  __syncwarp();
  if (w == 1 || w == 2 || w == 3 || w == 5) {
    for (int j = (w == 5 ? 96 : (w - 1) * 32) + (threadIdx.x & 31); j < n_old; j += 128) {
      bulk_load(smem_u32(c.smem + kSmVs + j * kD), v + (row0 + j) * ldk, kD, bar);
      // the K row of the same key: towards L1 (the attention phase reads the head slices with ld.global.ca)
      const int8_t* kr = k + (row0 + j) * ldk;
#pragma unroll
      for (int q = 0; q < 4; ++q) asm volatile("prefetch.global.L1 [%0];" ::"l"(kr + q * 128));
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk) {
    const int j = kk * 32 + lane;
    const bool old = j < n_old;
    const int jc = old ? j : 0;
    pre.skl[kk] = (n_old > 0) ? __ldcg(sk + (row0 + jc) * sstride) : 0.f;
    pre.svl[kk] = (n_old > 0) ? __ldcg(sv + (row0 + jc) * sstride) : 0.f;
    pre.keepl[kk] = (key_mask != nullptr && old) ? key_mask[static_cast<int64_t>(c.b) * mask_stride + j] : 1;
  }
}

// RowQuant (groups of 512 features) of my projection row in recv: NG = 3: q | k | v, NG = 1: cross-attention q.  Warp h handles the
// slices of head h -- lanes 0..15 its 64 q features (and, second register, its 64 v features), lanes 16..31 its 64 k features --
// which are exactly the bytes warp h consumes afterwards: one block barrier (the row maxima), then only the warp's own lanes.
__device__ __forceinline__ void quant_groups(Ctx& c, const int NG, float (&scale)[3]) {
  const int tid = threadIdx.x, h = tid >> 5, lane = tid & 31, l16 = lane & 15, hi = lane >> 4;
  uint32_t* rowbuf = reinterpret_cast<uint32_t*>(c.smem + kSmRow);
  float* red = misc(c) + kMiRed;
  const float4* y4 = reinterpret_cast<const float4*>(c.smem + kSmRecv);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool has0 = hi == 0 || NG == 3, has1 = hi == 0 && NG == 3;
  const float4 y0 = has0 ? y4[hi * 128 + h * 16 + l16] : zero;           // q (lanes 0..15) or k (lanes 16..31) slice of head h
  const float4 y1 = has1 ? y4[256 + h * 16 + l16] : zero;                // v slice of head h
  const float a0 = fmaxf(fmaxf(fabsf(y0.x), fabsf(y0.y)), fmaxf(fabsf(y0.z), fabsf(y0.w)));
  const float a1 = fmaxf(fmaxf(fabsf(y1.x), fabsf(y1.y)), fmaxf(fabsf(y1.z), fabsf(y1.w)));
  // maxima over each 16-lane half (non-negative floats order like their bit patterns; NaN counts as 0, as in warp_max_nonneg)
  uint32_t u0 = (a0 == a0) ? __float_as_uint(a0) : 0u, u1 = (a1 == a1) ? __float_as_uint(a1) : 0u;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    u0 = max(u0, __shfl_xor_sync(0xffffffffu, u0, o));
    u1 = max(u1, __shfl_xor_sync(0xffffffffu, u1, o));
  }
  if (l16 == 0) {
    red[hi * 8 + h] = __uint_as_float(u0);
    if (hi == 0) red[16 + h] = __uint_as_float(u1);
  }
  __syncthreads();
  const float4* r4 = reinterpret_cast<const float4*>(red);
  {
    const float4 a = r4[0], b = r4[1];
    scale[0] = quant_scale_x(fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))));
  }
  scale[1] = scale[2] = 0.f;
  if (NG == 3) {
    const float4 a = r4[2], b = r4[3], cc = r4[4], d = r4[5];
    scale[1] = quant_scale_x(fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))));
    scale[2] = quant_scale_x(fmaxf(fmaxf(fmaxf(cc.x, cc.y), fmaxf(cc.z, cc.w)), fmaxf(fmaxf(d.x, d.y), fmaxf(d.z, d.w))));
  }
  if (has0) {
    const float s = hi ? scale[1] : scale[0];
    rowbuf[hi * 128 + h * 16 + l16] = quant4_pack(y0, s, __frcp_rn(s));
  }
  if (has1) rowbuf[256 + h * 16 + l16] = quant4_pack(y1, scale[2], __frcp_rn(scale[2]));
  __syncwarp();
}

// self: RowQuant of q | k | v, KV-cache append, causal attention over t+1 keys; cross: RowQuant of q, attention over the S cached
// memory keys with the key-padding mask.  warp h = head h; the arithmetic of attention_decode_body (ot_attention_decode.cuh),
// instruction for instruction: K rows straight from the (L2-resident) cache into registers -- issued before the RowQuant, which
// hides their latency --, V rows from the prefetched shared-memory copy.  Ends with the all-gather of the quantized context row.
__device__ __forceinline__ void phase_attention(Ctx& c, bool self, int t, int l, AttnPre& pre) {
  const CdHot& P = *c.P;
  const CdLayer& L = P.layer[l];
  if (c.b < 0) return;
  const int tid = threadIdx.x, h = tid >> 5, lane = tid & 31;
  const int8_t* rowbuf = reinterpret_cast<const int8_t*>(c.smem + kSmRow);
  int8_t* Vs = reinterpret_cast<int8_t*>(c.smem + kSmVs);      // [key][512]
  float* red = misc(c) + kMiRed;
  const int Tk = self ? t + 1 : P.S;
  const int n_old = self ? t : P.S;
  // K: instruction i reads the head slices of keys 8i .. 8i+7, four lanes per key and 16 bytes per lane (64 contiguous bytes per
  // row: coalesced, unlike one key per lane), straight from the (L2-resident) cache; issued before the RowQuant, which hides the
  // latency.  Indices are clamped, never predicated on the loaded value.
  const int m8 = lane >> 2, ch = lane & 3;
  uint4 kq[4 * kDecKeysPerLane];
  {
    const int8_t* kbase = self ? L.kc + static_cast<int64_t>(c.b) * P.cap * kD : P.ckv + static_cast<int64_t>(c.b) * P.S * (2 * kD * P.n_layers) + 2 * kD * l;
    const int64_t ldk = self ? kD : 2 * kD * P.n_layers;
#pragma unroll
    for (int i = 0; i < 4 * kDecKeysPerLane; ++i) {
      kq[i] = make_uint4(0, 0, 0, 0);
      if (8 * i < n_old) {
        const int j = min(8 * i + m8, n_old - 1);
        kq[i] = __ldca(reinterpret_cast<const uint4*>(kbase + j * ldk + h * kDk) + ch);
      }
    }
  }
  float sc3[3];
  mark(c, 50);
  quant_groups(c, self ? 3 : 1, sc3);
  mark(c, 51);
  mbar_wait(smem_u32(&c.bars[kBarKv]), c.kv_parity);
  c.kv_parity ^= 1u;
  mark(c, 52);
  if (self) {
    // this step's K / V row: into the cache (global), V also next to the prefetched rows, K into the registers of its 4 lanes.
    // Every warp moves the 64 bytes of its own head (lanes 0..3: k, lanes 4..7: v): no block barrier.
    const int64_t dst = (static_cast<int64_t>(c.b) * P.cap + t) * kD + h * kDk;
    if (lane < 4) {
      *reinterpret_cast<uint4*>(L.kc + dst + lane * 16) = *reinterpret_cast<const uint4*>(rowbuf + kD + h * kDk + lane * 16);
    } else if (lane < 8) {
      const int c16 = lane - 4;
      const uint4 vv = *reinterpret_cast<const uint4*>(rowbuf + 2 * kD + h * kDk + c16 * 16);
      *reinterpret_cast<uint4*>(Vs + t * kD + h * kDk + c16 * 16) = vv;
      *reinterpret_cast<uint4*>(L.vc + dst + c16 * 16) = vv;
    } else if (tid == 8) {
      L.skc[static_cast<int64_t>(c.b) * P.cap + t] = sc3[1];
      L.svc[static_cast<int64_t>(c.b) * P.cap + t] = sc3[2];
    }
    if (lane < 8) asm volatile("fence.proxy.async.global;" ::: "memory");   // later steps read these rows through bulk copies (async proxy)
#pragma unroll
    for (int kk = 0; kk < kDecKeysPerLane; ++kk)
      if (kk * 32 + lane == t) { pre.skl[kk] = sc3[1]; pre.svl[kk] = sc3[2]; }
#pragma unroll
    for (int i = 0; i < 4 * kDecKeysPerLane; ++i)
      if (8 * i + m8 == t) kq[i] = *reinterpret_cast<const uint4*>(rowbuf + kD + h * kDk + ch * 16);
    __syncwarp();
  }
  const int q_pos0 = self ? t : 0, mask_kind = self ? 2 : 1;
  const float sqi = sc3[0];
  mark(c, 53);
  // int8 dot products: 4 dp4a per lane on its 16 bytes, summed over the key's 4 lanes (integer: exact in any order), then handed
  // to the lane that owns the key in the softmax (key 32*kk + lane)
  int dotl[kDecKeysPerLane];
  {
    const uint4 qv = *reinterpret_cast<const uint4*>(rowbuf + h * kDk + ch * 16);
#pragma unroll
    for (int kk = 0; kk < kDecKeysPerLane; ++kk) {
      dotl[kk] = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 tk = kq[4 * kk + q];
        int d = __dp4a(static_cast<int>(qv.x), static_cast<int>(tk.x), 0);
        d = __dp4a(static_cast<int>(qv.y), static_cast<int>(tk.y), d);
        d = __dp4a(static_cast<int>(qv.z), static_cast<int>(tk.z), d);
        d = __dp4a(static_cast<int>(qv.w), static_cast<int>(tk.w), d);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        const int got = __shfl_sync(0xffffffffu, d, 4 * (lane & 7));     // key 8*(4kk+q) + (lane & 7)
        if ((lane >> 3) == q) dotl[kk] = got;
      }
    }
  }
  float sc[kDecKeysPerLane], svl[kDecKeysPerLane];
  float mx = -INFINITY;
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk) {
    const int j = kk * 32 + lane;
    const float s = __fmul_rn(__fmul_rn(__fmul_rn(__int2float_rn(dotl[kk]), sqi), pre.skl[kk]), 0.125f);      // / sqrt(d_k) = / 8: exact scaling
    const bool visible = pre.keepl[kk] != 0 && (mask_kind != 2 || j <= q_pos0);
    const bool live = j < Tk;
    sc[kk] = live ? (visible ? s : -1e9f) : -INFINITY;
    svl[kk] = live ? pre.svl[kk] : 0.f;
    mx = live ? fmaxf(mx, sc[kk]) : mx;
  }
  mark(c, 54);
  mx = warp_max_any(mx);
  float sum = 0.f;
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk) {
    if (kk * 32 + lane < Tk) {
      sc[kk] = expf(__fsub_rn(sc[kk], mx));
      sum += sc[kk];
    }
  }
  sum = warp_sum_f(sum);
  // quantized probability (already divided by 127, attention.py:35) and V scale of every key, staged for broadcast reads: the
  // P.V loop is bound by the shared-memory / shuffle pipe, and one 16-byte broadcast load per TWO keys replaces four shuffles
  float2* pv = reinterpret_cast<float2*>(c.smem + kSmRecv) + h * kMaxKeys;     // (the scattered q|k|v row has been consumed)
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk) {
    const float pqk = (kk * 32 + lane < Tk) ? div127_exact(rintf(__fmul_rn(__fdiv_rn(sc[kk], sum), 127.0f))) : 0.f;
    pv[kk * 32 + lane] = make_float2(pqk, svl[kk]);
  }
  __syncwarp();
  mark(c, 55);
  float acc0 = 0.f, acc1 = 0.f;
  const int d0 = 2 * lane;
  // keys in order j = 0..; a key with p = 0 (masked, or the odd one at Tk where sv = 0 and the V bytes are stale but finite)
  // contributes exactly +-0, as in attention_decode_body: adding +0 to a sum that is never -0 changes nothing.  The int8 -> fp32
  // conversions are exact bit constructions (s8_as_float) on the integer / FMA pipes: I2F shares the MIO queue with the
  // shared-memory loads of this loop and was its bound (profiles/r1_ncu_cdecoder_v4_final.txt: stall_mio on the I2F lines).
  {
    const float4* pv4 = reinterpret_cast<const float4*>(pv);
    const uint8_t* vcol = reinterpret_cast<const uint8_t*>(Vs) + h * kDk + d0;
#pragma unroll 8
    for (int j = 0; j < Tk; j += 2) {
      const float4 ps = pv4[j >> 1];                                             // p, sv of keys j and j+1
      const uint32_t wa = static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(vcol + j * kD)) ^ 0x8080u;
      const uint32_t wb = static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(vcol + (j + 1) * kD)) ^ 0x8080u;
      acc0 = fmaf(ps.x, __fmul_rn(s8_as_float<0>(wa), ps.y), acc0);
      acc1 = fmaf(ps.x, __fmul_rn(s8_as_float<1>(wa), ps.y), acc1);
      acc0 = fmaf(ps.z, __fmul_rn(s8_as_float<0>(wb), ps.w), acc0);
      acc1 = fmaf(ps.z, __fmul_rn(s8_as_float<1>(wb), ps.w), acc1);
    }
  }
  // RowQuant of the merged context row (all 8 heads): the row abs-max is a max (exact in any order), every lane quantizes its own
  // two features with the instructions of attention_decode_body
  mark(c, 56);
  const float am = warp_max_nonneg(fmaxf(fabsf(acc0), fabsf(acc1)));
  if (lane == 0) red[24 + h] = am;
  __syncthreads();
  const float amax = fmaxf(fmaxf(fmaxf(red[24], red[25]), fmaxf(red[26], red[27])), fmaxf(fmaxf(red[28], red[29]), fmaxf(red[30], red[31])));
  const float s = quant_scale_x(amax);
  const uint32_t q01 = quant4_pack(make_float4(acc0, acc1, 0.f, 0.f), s, __frcp_rn(s));
  *reinterpret_cast<uint16_t*>(c.smem + kSmRow + 1536 + h * kDk + d0) = static_cast<uint16_t>(q01 & 0xFFFFu);
  __syncwarp();
  mark(c, 58);
  // all-gather: warp h sends the 64 bytes of head h (4 chunks of 16) to the 8 CTAs -- lane = (peer, chunk) -- in push_row_q8's layout
  {
    const uint32_t gbar = smem_u32(&c.bars[kBarG]);
    const int peer = lane >> 2, ch = h * 4 + (lane & 3);
    const uint4 v = *reinterpret_cast<const uint4*>(c.smem + kSmRow + 1536 + ch * 16);
    const int kb = ch >> 3, cc = ch & 7;
    const uint32_t local = smem_u32(c.smem + kSmBx + kb * 1024 + c.rank * 128 + ((cc ^ (c.rank & 7)) << 4));
    st_async_v4(mapa_shared(local, peer), v, mapa_shared(gbar, peer));
    if (tid < kCS) st_async_b32(mapa_shared(smem_u32(misc(c) + kMiSB + c.rank), tid), __float_as_uint(s), mapa_shared(gbar, tid));
  }
  mark(c, 59);
}

// ------------------------------------------------------------------------------------------------ generator
// logits[s, v] = bias[v] + sum_k h[s,k] * W[v,k]  (k ascending, fmaf: the order of generator_logits_kernel).  The vocabulary is cut
// into tiles of 32 entries; tile j belongs to CTA j % 8 of every cluster, and inside the CTA to warp (j / 8) % 8.  A lane owns one
// entry of each of its warp's (up to NT) tiles for all 8 sentences; weights arrive as one coalesced 16-byte load per lane per 4 k.
// Packed fp32 pairs stay in 64-bit registers for the whole loop: one FFMA2 (fma.rn.f32x2, sm_100) = two IEEE fp32 FMAs for one
// issue slot.  (Packing / unpacking around every FMA cost four extra moves each: the first version of this loop was 50 % MOV / IMAD.)
__device__ __forceinline__ void fma2(unsigned long long& acc, const unsigned long long a, const unsigned long long ww) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(ww));
}
__device__ __forceinline__ unsigned long long dup2(const float w) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(w));
  return r;
}
__device__ __forceinline__ float2 unpack2(const unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

template <int NT>
__device__ __forceinline__ void generator_warp(const CdHot& P, const float* hT, int rank, int warp, int lane, float (&best)[kCS], int (&bidx)[kCS]) {
  unsigned long long acc[NT][kCS / 2];            // [tile][sentence pair], packed (lo = even sentence)
  const float4* wp[NT];
  int v[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int tile = rank + kCS * (warp + 8 * i);
    wp[i] = reinterpret_cast<const float4*>(P.gen_w4) + static_cast<int64_t>(tile) * 128 * 32 + lane;
    v[i] = tile * kGenVT + lane;
#pragma unroll
    for (int s = 0; s < kCS / 2; ++s) acc[i][s] = 0ull;
  }
  const ulonglong2* h2 = reinterpret_cast<const ulonglong2*>(hT);      // k-major: h2[2k] = sentence pairs (0,1),(2,3); h2[2k+1] = (4,5),(6,7)
  constexpr int G = 4;                     // k-quads per register group: the next group's weights are in flight during this group's FMAs
  float4 wn[G][NT];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < NT; ++i) wn[g][i] = __ldg(wp[i] + g * 32);
#pragma unroll 1
  for (int k0 = 0; k0 < 128; k0 += G) {
    float4 w[G][NT];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        w[g][i] = wn[g][i];
        if (k0 + G < 128) wn[g][i] = __ldg(wp[i] + (k0 + G + g) * 32);
      }
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {      // k = 4*(k0+g) + kk, ascending: the fmaf chain of generator_logits_kernel per (sentence, entry)
        const ulonglong2 ha = h2[2 * (4 * (k0 + g) + kk)], hb = h2[2 * (4 * (k0 + g) + kk) + 1];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          const unsigned long long ww = dup2(kk == 0 ? w[g][i].x : (kk == 1 ? w[g][i].y : (kk == 2 ? w[g][i].z : w[g][i].w)));
          fma2(acc[i][0], ha.x, ww);
          fma2(acc[i][1], ha.y, ww);
          fma2(acc[i][2], hb.x, ww);
          fma2(acc[i][3], hb.y, ww);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const bool ok = v[i] < P.vocab;
    const float bv = (ok && P.gen_b) ? __ldg(P.gen_b + v[i]) : 0.f;
#pragma unroll
    for (int s2 = 0; s2 < kCS / 2; ++s2) {
      const float2 pr = unpack2(acc[i][s2]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int s = 2 * s2 + e;
        float lg = __fadd_rn(e ? pr.y : pr.x, bv);
        if (lg != lg) lg = INFINITY;               // torch.max / np.argmax: a NaN logit ranks above every number
        if (ok && (lg > best[s] || (lg == best[s] && v[i] < bidx[s]))) { best[s] = lg; bidx[s] = v[i]; }
      }
    }
  }
}

__device__ __forceinline__ void phase_generator(Ctx& c) {
  const CdHot& P = *c.P;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // the gathered rows [sentence][512] -> k-major [512][8 sentences] (second 16 KB of the V region): a sentence PAIR of one
  // feature is then one 64-bit operand of an FFMA2
  const float* hb = reinterpret_cast<const float*>(c.smem + kSmHb);
  float* hT = reinterpret_cast<float*>(c.smem + kSmHb + kCS * kD * 4);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int k = tid + 256 * j;
    float hv[kCS];
#pragma unroll
    for (int s = 0; s < kCS; ++s) hv[s] = hb[s * kD + k];
    reinterpret_cast<float4*>(hT)[2 * k] = make_float4(hv[0], hv[1], hv[2], hv[3]);
    reinterpret_cast<float4*>(hT)[2 * k + 1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
  }
  __syncthreads();
  float best[kCS];
  int bidx[kCS];
#pragma unroll
  for (int s = 0; s < kCS; ++s) { best[s] = -INFINITY; bidx[s] = 0x7fffffff; }
  // tiles of this warp: rank + 8*(warp + 8*i) < n_gen_tiles
  int nt = 0;
  while (nt < 3 && c.rank + kCS * (warp + 8 * nt) < P.n_gen_tiles) ++nt;
  if (nt == 3) generator_warp<3>(P, hT, c.rank, warp, lane, best, bidx);
  else if (nt == 2) generator_warp<2>(P, hT, c.rank, warp, lane, best, bidx);
  else if (nt == 1) generator_warp<1>(P, hT, c.rank, warp, lane, best, bidx);
  float* pv = misc(c) + kMiPartV;
  int* pi = reinterpret_cast<int*>(misc(c) + kMiPartI);
#pragma unroll
  for (int s = 0; s < kCS; ++s) {
    float bv = best[s];
    int bi = bidx[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > bv || (ob == bv && oi < bi)) { bv = ob; bi = oi; }
    }
    if (lane == 0) { pv[warp * kCS + s] = bv; pi[warp * kCS + s] = bi; }
  }
  __syncthreads();
  if (tid < c.n_own) {                 // sentence tid: combine the 8 warps, send to its owner
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float ob = pv[w * kCS + tid];
      const int oi = pi[w * kCS + tid];
      if (ob > bv || (ob == bv && oi < bi)) { bv = ob; bi = oi; }
    }
    const uint32_t sbar = mapa_shared(smem_u32(&c.bars[kBarS]), tid);
    st_async_b32(mapa_shared(smem_u32(misc(c) + kMiGenV + c.rank), tid), __float_as_uint(bv), sbar);
    st_async_b32(mapa_shared(smem_u32(misc(c) + kMiGenI + c.rank), tid), static_cast<uint32_t>(bi), sbar);
  }
}
// ------------------------------------------------------------------------------------------------ generator, tensor-core screening
// The exact generator above is FP32-bound (35.8 k FFMA2 per CTA >= 12 us) and streams 1.1 MB of fp32 weights per CTA and step.  Here
// the tensor cores SCREEN the vocabulary and the fp32 chain only re-evaluates the entries that can still win:
//   1. approximate logits of the CTA's 576 vocabulary rows for the 8 sentences: tcgen05.mma.kind::f16 (M = 64: the 8 sentences,
//      N = 192, K = 16 per instruction) on fp16 copies of h and W, fp32 accumulation in TMEM; the fp16 weight slice (576 KB) streams
//      through the weight ring's slots (the ring does not run ahead into the next step across this phase);
//      (+ the fp32 bias, so that the screened quantity is the logit itself);
//   2. |approx - exact| <= eps_s = ||h_s||_2 max_v||w_v||_2 2^-9 + ||h_s||_1 2^-24 + max_v||w_v||_2 1e-6 (fp16 rounding of both operands
//      is 2^-10 relative to sum |h_k w_k| <= ||h|| ||w||; the other half of the 2^-9 covers the accumulation and the fp32 chain's own
//      rounding; the absolute terms cover fp16 underflow of either operand).  The CTAs all-gather their per-sentence maxima; an entry
//      can be the arg-max only if its approximate logit is within 2 eps_s of the largest approximate logit M_s;
//   3. those entries (typically one or two per sentence and cluster) are evaluated EXACTLY -- k-ascending fmaf chain + bias, the
//      arithmetic of generator_logits_kernel -- and compared by (logit, lowest index), so the token is the exact generator's token.
// Anything not finite (NaN / Inf in h, fp16 overflow) makes every CTA of the cluster fall back to the exact generator for that step.
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 instruction descriptor: D = F32 (1 at [4,6)), A = B = F16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ float* gen_logits(Ctx& c, int rb) {      // [8 sentences][192] approximate logits of row block rb
  return reinterpret_cast<float*>(c.smem + (rb == 0 ? kSmRecv : rb == 1 ? kSmX : kSmGen));
}

__device__ __forceinline__ void phase_generator_tc(Ctx& c, bool more_steps) {
  const CdHot& P = *c.P;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  uint64_t* gb = reinterpret_cast<uint64_t*>(c.smem + kSmGenBars);     // gfull[0..3] gempty[4..7] gaccf[8..9] gacce[10..11] gmax[12]
  const float* hb = reinterpret_cast<const float*>(c.smem + kSmHb);
  // ---- a. fp16 A operand (8 sentence rows per k-block of 64, 128-byte swizzle) + the norms of my warp's sentence
  float eps_s;
  {
    const int srow = warp;                                              // warp = sentence
    const float4* h4 = reinterpret_cast<const float4*>(hb + srow * kD + lane * 16);
    float v[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float4 x = h4[q]; v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w; }
    float ss = 0.f, sa = 0.f;
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ss = fmaf(v[2 * j], v[2 * j], fmaf(v[2 * j + 1], v[2 * j + 1], ss));
      sa += fabsf(v[2 * j]) + fabsf(v[2 * j + 1]);
      const __half2 hh = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
      pk[j] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    ss = warp_sum(ss);
    sa = warp_sum(sa);
    eps_s = fmaf(sqrtf(ss) * 1.0001f, P.gen_eps, fmaf(sa, 5.9604645e-08f, P.gen_abs));
    const int kb = lane >> 2, c16 = (2 * lane) & 7;
    uint8_t* arow = c.smem + kSmBh + kb * 1024 + srow * 128;
    *reinterpret_cast<uint4*>(arow + ((c16 ^ (srow & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(arow + (((c16 + 1) ^ (srow & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
  if (tid == 0) mbar_arrive_expect_tx(smem_u32(&gb[12]), kCS * kCS * 4);      // the gathered maxima of this step
  fence_proxy_async_smem();
  __syncthreads();
  if (c.trace_on && tid == 0) P.trace[240] = tl_now();
  // ---- b. approximate logits: loader warp / issuer warp / two epilogue warps (TMEM lanes 0..15 belong to warps 0 and 4)
  const uint32_t G0 = c.gen_n;
  if (warp_u == kLoader / 32) {
    for (int j = kSlots; j < kGenChunks; ++j) gen_issue_chunk(c, j);       // chunks 0..3 were issued behind the last FFN2 (phase_gemm)
    // hand the slots back to the layer weights: the last generator chunk of every slot has been consumed, then the ring runs ahead again
    for (int j = kGenChunks - kSlots; j < kGenChunks; ++j) {
      const uint32_t G = G0 + j;
      mbar_wait(smem_u32(&gb[4 + G % kSlots]), (G / kSlots) & 1);
    }
    if (more_steps) fill_until(c, c.pn + kSlots);
  } else if (warp_u == kIssuer / 32) {
    const uint32_t idesc = make_idesc_f16(64, kGenRB);
    const uint32_t abase = smem_u32(c.smem + kSmBh);
    for (int rb = 0; rb < 3; ++rb) {
      const int buf = rb & 1;
      const uint32_t use = c.gen_acc[buf] + (rb >> 1);                   // buffer 0 is used twice per step
      if (use > 0) { mbar_wait(smem_u32(&gb[10 + buf]), (use - 1) & 1); tc_fence_after(); }
      for (int kb = 0; kb < 8; ++kb) {
        const uint32_t G = G0 + rb * 8 + kb;
        const int slot = G % kSlots;
        mbar_wait(smem_u32(&gb[slot]), (G / kSlots) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = make_smem_desc_sw128(abase + kb * 1024);
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(c.smem + kSmRing + slot * kSlotBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_f16_ss(c.tmem + buf * 256, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          mma_commit(smem_u32(&gb[4 + slot]));
        }
        __syncwarp();
      }
      if (elect_one()) mma_commit(smem_u32(&gb[8 + buf]));
      __syncwarp();
    }
  } else if ((warp & 3) == 0) {
    const int col0 = (warp >> 2) * 96, srow = lane >> 2;
    for (int rb = 0; rb < 3; ++rb) {
      const int buf = rb & 1;
      const uint32_t use = c.gen_acc[buf] + (rb >> 1);
      mbar_wait(smem_u32(&gb[8 + buf]), use & 1);
      tc_fence_after();
      float* L = gen_logits(c, rb) + srow * kGenRB + col0;
#pragma unroll
      for (int gq = 0; gq < 3; ++gq) {
        uint32_t r[16];
        tmem_ld_16x256b_x4(c.tmem + buf * 256 + col0 + 32 * gq, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float2*>(L + 32 * gq + 8 * i + 2 * (lane & 3)) = make_float2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&gb[10 + buf]));
    }
  }
  c.gen_n += kGenChunks;
  c.gen_acc[0] += 2;
  c.gen_acc[1] += 1;
  __syncthreads();
  if (c.trace_on && tid == 0) P.trace[241] = tl_now();
  // ---- c. per-sentence maximum of my 576 approximate logits -> all-gather (warp = sentence, lane p sends to CTA p)
  const int srow = warp;
  float lmax = -INFINITY;
  bool lbad = false;
  {
    // + bias (the exact logit has it), coalesced and all in flight at once; rows past the vocabulary become -Inf: never candidates
    float bias[kGenRows / 32];
#pragma unroll
    for (int q = 0; q < kGenRows / 32; ++q) {
      const int v = c.rank * kGenRows + q * 32 + lane;
      bias[q] = (P.gen_b && v < P.vocab) ? __ldg(P.gen_b + v) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < kGenRows / 32; ++q) {
      const int col = q * 32 + lane;
      float* L = gen_logits(c, col / kGenRB) + srow * kGenRB + col % kGenRB;
      const float x = (c.rank * kGenRows + col < P.vocab) ? __fadd_rn(*L, bias[q]) : -INFINITY;
      *L = x;
      lbad = lbad || (x != x) || (x > 3.0e38f);          // NaN / +Inf
      lmax = fmaxf(lmax, x);
    }
  }
  lmax = warp_max_any(lmax);
  lbad = __any_sync(0xffffffffu, lbad);
  if (lbad) lmax = __uint_as_float(0x7fc00000u);
  float* gm = reinterpret_cast<float*>(c.smem + kSmGenMax);
  if (lane < kCS) st_async_b32(mapa_shared(smem_u32(gm + c.rank * kCS + srow), lane), __float_as_uint(lmax), mapa_shared(smem_u32(&gb[12]), lane));
  mbar_wait(smem_u32(&gb[12]), c.gen_max_parity);
  c.gen_max_parity ^= 1u;
  if (c.trace_on && tid == 0) P.trace[242] = tl_now();
  float M = (lane < kCS) ? gm[lane * kCS + srow] : -INFINITY;
  bool bad = (lane < kCS) && !(fabsf(M) <= 3.0e38f);
  bad = __any_sync(0xffffffffu, bad) || !(eps_s <= 3.0e38f);
  M = warp_max_any(M);
  // cluster-uniform decision: every CTA sees the same 64 maxima and the same eps_s
  int* flag = reinterpret_cast<int*>(misc(c) + kMiRed);
  if (tid == 0) *flag = 0;
  __syncthreads();
  if (bad && lane == 0 && srow < c.n_own) atomicOr(flag, 1);
  __syncthreads();
  if (*flag != 0) {
    __syncthreads();
    phase_generator(c);                  // exact generator for this step (h is still in place)
    return;
  }
  if (c.trace_on && tid == 0) P.trace[243] = tl_now();
  // ---- d. exact re-evaluation of the entries within 2 eps of the maximum; first arg-max by (logit, lowest index)
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  if (srow < c.n_own) {
    const float thr = M - 2.0f * eps_s;
    const float4* hrow = reinterpret_cast<const float4*>(hb + srow * kD);
#pragma unroll 1
    for (int q = 0; q < kGenRows / 32; ++q) {
      const int col = q * 32 + lane;                                      // 0..575 within my slice
      const float x = gen_logits(c, col / kGenRB)[srow * kGenRB + col % kGenRB];
      const int v = c.rank * kGenRows + col;
      unsigned todo = __ballot_sync(0xffffffffu, x >= thr && v < P.vocab);
      while (todo != 0u) {
        // one candidate at a time, the whole warp on it: the 2 KB weight row arrives with ONE L2 round trip (lane l fetches k-quads
        // l, l + 32, l + 64, l + 96) and is staged in this warp's 2 KB of shared memory; then every lane runs the k-ascending fmaf
        // chain of generator_logits_kernel on broadcast loads (512 dependent FMAs: ~1 us; shuffling the weights instead cost 3 us)
        const int src = __ffs(todo) - 1;
        todo &= todo - 1u;
        const int vv = __shfl_sync(0xffffffffu, v, src);
        const float4* wp = reinterpret_cast<const float4*>(P.gen_w4) + static_cast<int64_t>(vv >> 5) * 128 * 32 + (vv & 31);
        float4* wrow = reinterpret_cast<float4*>(c.smem + kSmHb + kCS * kD * 4) + warp * 128;      // the exact generator's hT region
        {
          float4 wreg[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) wreg[j] = __ldg(wp + (lane + 32 * j) * 32);
          __syncwarp();                                                    // the previous candidate's chain has read the row
#pragma unroll
          for (int j = 0; j < 4; ++j) wrow[lane + 32 * j] = wreg[j];
        }
        const float bias = P.gen_b ? __ldg(P.gen_b + vv) : 0.f;
        __syncwarp();
        float acc = 0.f;
#pragma unroll 8
        for (int k4 = 0; k4 < 128; ++k4) {
          const float4 w = wrow[k4];
          const float4 hv = hrow[k4];
          acc = __fmaf_rn(hv.x, w.x, acc);
          acc = __fmaf_rn(hv.y, w.y, acc);
          acc = __fmaf_rn(hv.z, w.z, acc);
          acc = __fmaf_rn(hv.w, w.w, acc);
        }
        float lg = __fadd_rn(acc, bias);
        if (lg != lg) lg = INFINITY;
        if (lg > best || (lg == best && vv < bidx)) { best = lg; bidx = vv; }       // uniform across the warp
      }
    }
    if (lane == 0) {
      const uint32_t sbar = mapa_shared(smem_u32(&c.bars[kBarS]), srow);
      st_async_b32(mapa_shared(smem_u32(misc(c) + kMiGenV + c.rank), srow), __float_as_uint(best), sbar);
      st_async_b32(mapa_shared(smem_u32(misc(c) + kMiGenI + c.rank), srow), static_cast<uint32_t>(bidx), sbar);
    }
  }
}
// owner, warp 0: first arg-max over the 8 vocabulary slices
__device__ __forceinline__ int generator_pick(Ctx& c, int lane) {
  const float* gv = misc(c) + kMiGenV;
  const int* gi = reinterpret_cast<const int*>(misc(c) + kMiGenI);
  float bv = (lane < kCS) ? gv[lane] : -INFINITY;
  int bi = (lane < kCS) ? gi[lane] : 0x7fffffff;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > bv || (ob == bv && oi < bi)) { bv = ob; bi = oi; }
  }
  bi = __shfl_sync(0xffffffffu, bi, 0);
  return (bi >= 0 && bi < c.P->vocab) ? bi : 0;
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kThreads, 1) cdecoder_kernel(const CdPlan* __restrict__ plan, int t0, int n_steps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  Ctx c;
  c.G = plan;
  c.smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);    // same offset in every CTA of the cluster
  c.P = reinterpret_cast<const CdHot*>(c.smem + kSmHot);
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(CdHot) / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(c.smem + kSmHot)[i] = reinterpret_cast<const uint4*>(&plan->hot)[i];
  c.bars = reinterpret_cast<uint64_t*>(c.smem + kSmBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c.smem + kSmBars + 120);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 7) {
    if (lane == 0) {
      for (int i = 0; i <= kBarM; ++i) mbar_init(smem_u32(&c.bars[i]), 1);
      {
        uint64_t* gb = reinterpret_cast<uint64_t*>(c.smem + kSmGenBars);
        for (int i = 0; i < 13; ++i) mbar_init(smem_u32(&gb[i]), (i == 10 || i == 11) ? 2 : 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *tmem_slot;
  const CdHot& P = *c.P;
  const int nl = P.n_layers;
  c.rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x / kCS;
  c.n_own = min(P.spc, P.B - cluster_id * P.spc);
  c.b = (c.rank < c.n_own) ? cluster_id * P.spc + c.rank : -1;
  c.pn = c.cn = 0;
  c.total = static_cast<uint32_t>(n_steps) * nl * kChunks;
  c.acc_parity = c.kv_parity = c.g_parity = c.s_parity = c.m_parity = 0;
  c.gen_n = c.gen_acc[0] = c.gen_acc[1] = c.gen_max_parity = 0;
  c.trace_slot = 0;
  c.trace_on = false;
  c.fine = false;
  c.mark_slot = 0;
  c.t_step = 0;
  if (__shfl_sync(0xffffffffu, warp, 0) == kLoader / 32) fill_until(c, kSlots);
  if (tid == 0 && n_steps > 0) {       // first use of the gather / scatter barriers (the first step reads its token from ys: no token exchange)
    mbar_arrive_expect_tx(smem_u32(&c.bars[kBarG]), static_cast<uint32_t>(c.n_own) * (kD + 4));
    if (c.b >= 0) mbar_arrive_expect_tx(smem_u32(&c.bars[kBarS]), 3 * kD * 4);
  }
  __syncwarp();
  // every CTA of the cluster is running (its shared memory may be written) and has its barriers initialised
  cluster_arrive_release();
  cluster_wait_acquire();

  const int t_last = t0 + n_steps - 1;
  const bool own = c.b >= 0;
  AttnPre pre;
#pragma unroll 1
  for (int t = t0; t <= t_last; ++t) {
    c.trace_slot = 0;
    c.trace_on = P.trace != nullptr && blockIdx.x == 0 && t == t_last;
    if (c.trace_on && tid == 0) { c.t_step = tl_now(); P.trace[255] = c.t_step; }
    // ---- token (arg-max over the 8 vocabulary slices of the previous step) -> embedding + positional encoding -> LayerNorm 1
    //      of layer 0 -> all-gather
    {
      xwait(c, kBarS, 3 * kD * 4, c.s_parity, own && t > t0);            // 64 bytes: the 8 slices' best (logit, index); next: q|k|v
      int* tok = reinterpret_cast<int*>(misc(c) + kMiTok);
      if (own && warp == 0) {
        int64_t token;
        if (t > t0) {
          token = generator_pick(c, lane);
          if (lane == 0) P.ys[static_cast<int64_t>(c.b) * P.ys_ld + t] = token;
        } else {
          token = __ldcg(P.ys + static_cast<int64_t>(c.b) * P.ys_ld + t);
        }
        if (lane == 0) *tok = static_cast<int>(token);
      }
      __syncthreads();
      phase_ln(c, 0, *tok, t, P.layer[0].ln_g[0], P.layer[0].ln_b[0], true);
    }
#pragma unroll 1
    for (int l = 0; l < nl; ++l) {
      const CdLayer& L = P.layer[l];
      c.fine = c.trace_on && l == min(2, nl - 1);
#pragma unroll 1
      for (int q = 0; q < 11; ++q) {
        if (q == 9) continue;      // the FFN hidden rows are quantized and exchanged by the FFN1 epilogue itself
        if ((q & 1) == 0) {
          // ---- GEMM phases; the attention operands of my sentence are prefetched alongside
          if ((q == 0 || q == 4) && own) {
            const bool self = q == 0;
            attn_prefetch(c, self ? t : P.S, self ? L.kc : P.ckv + 2 * kD * l, self ? L.vc : P.ckv + 2 * kD * l + kD, self ? kD : 2 * kD * nl,
                          static_cast<int64_t>(c.b) * (self ? P.cap : P.S), self ? L.skc : P.sckv + 2 * l, self ? L.svc : P.sckv + 2 * l + 1,
                          self ? 1 : 2 * nl, self ? nullptr : P.mask, P.S, pre);
          }
          // index one past this GEMM's last weight chunk in the launch-wide chunk sequence
          const int g = q >> 1;
          const uint32_t gend = (static_cast<uint32_t>(t - t0) * nl + l) * kChunks + (g == 0 ? 4 : g == 1 ? 5 : g == 2 ? 6 : g == 3 ? 7 : g == 4 ? 11 : 15);
          // bytes of the gather that follows this one: operand rows + scales; FFN2's two 2 KB halves; the generator's fp32 rows
          const uint32_t row_bytes = static_cast<uint32_t>(c.n_own) * (kD + 4);
          const uint32_t next_gather = g < 4 ? row_bytes : g == 4 ? 2u * kCS * 256u : (l + 1 < nl ? row_bytes : static_cast<uint32_t>(c.n_own) * kD * 4);
          if (g == 5 && own && warp < 4) {      // thread i of the LayerNorm that follows: its own 16 bytes of FFN2's scales and bias
            const int i4 = warp * 32 + lane;
            cp_async16(smem_u32(c.smem + kSmW2) + 16u * i4, reinterpret_cast<const float4*>(L.sw[5]) + i4);
            cp_async16(smem_u32(c.smem + kSmW2) + 2048u + 16u * i4, reinterpret_cast<const float4*>(L.bias[5]) + i4);
            asm volatile("cp.async.commit_group;" ::: "memory");
          }
          phase_gemm(c, l, g, gend, next_gather);
        } else {
          // ---- row phases: my sentence's row has arrived from the 8 column owners
          xwait(c, kBarS, (q == 7 ? 4 * kD : kD) * 4, c.s_parity, own);     // next scatter: a 512-float row, or FFN2's 4 split-K planes
          if (q == 1 || q == 5) phase_attention(c, q == 1, t, l, pre);
          else phase_ln(c, 1, 0, t, L.ln_g[q == 3 ? 1 : 2], L.ln_b[q == 3 ? 1 : 2], true);
        }
      }
      // ---- residual + LayerNorm 1 of the next layer, or the final norm (fp32 row to every CTA's generator input)
      xwait(c, kBarS, l + 1 < nl ? 3 * kD * 4 : 64, c.s_parity, own);      // 4 split-K planes of int32 partials; next: q|k|v or the token
      if (l + 1 < nl) phase_ln(c, 2, 0, t, P.layer[l + 1].ln_g[0], P.layer[l + 1].ln_b[0], true);
      else phase_ln(c, 2, 0, t, P.fin_g, P.fin_b, false);
    }
    c.fine = false;
    xwait(c, kBarG, t < t_last ? static_cast<uint32_t>(c.n_own) * (kD + 4) : 0u, c.g_parity, true);
    if (P.gen_tc) phase_generator_tc(c, t < t_last);
    else phase_generator(c);
    if (c.trace_on && tid == 0) P.trace[254] = tl_now();
  }
  // the last step's token
  xwait(c, kBarS, 0, c.s_parity, own);
  if (own && warp == 0) {
    const int id = generator_pick(c, lane);
    if (lane == 0) P.ys[static_cast<int64_t>(c.b) * P.ys_ld + t_last + 1] = id;
  }
  tc_fence_before();
  __syncthreads();
  // no CTA exits while a peer may still address its shared memory
  cluster_arrive_release();
  cluster_wait_acquire();
  if (warp == 7) tmem_dealloc(c.tmem, kTmemCols);
}

}  // namespace cd
}  // namespace ot

using namespace ot;
using namespace ot::cd;

#include <mutex>
#include <unordered_map>

// Host-side record of what a device plan was built for: ot_cdecoder_run checks its arguments against it (a mismatching B / spc would
// change the cluster count and the indexing; t0 + n_steps past the cache capacity would write the KV caches and ys out of bounds).
namespace {
struct CdPlanInfo { int B, spc, cap; };
std::mutex g_plan_mu;
std::unordered_map<const void*, CdPlanInfo> g_plans;
}  // namespace

extern "C" int ot_cdecoder_plan_size(void) { return static_cast<int>(sizeof(CdPlan)); }

// How many of the decoder's 8-CTA clusters the current device keeps resident at once (cudaOccupancyMaxActiveClusters: 15 on a
// 148-SM B200 -- clusters live inside GPCs).  A greedy step takes the same time whether 1 or all of them are in use, so callers that
// choose their own batch (fault-trial campaigns: independent trials) get the most out of a launch at 8 x this many sentences; one
// cluster more and the launch runs in two waves.
extern "C" int ot_cdecoder_max_clusters(int* n_out) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(n_out != nullptr, "null argument");
  OT_CHECK_CUDA(cudaFuncSetAttribute(cdecoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmTotal + 1024));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCS * 64);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmTotal + 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  OT_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, cdecoder_kernel, &cfg));
  *n_out = n;
  return OT_OK;
}

// layer_ptrs: n_layers x 28 device pointers in the order
//   ln1_g ln1_b ln2_g ln2_b ln3_g ln3_b | qkv_w qkv_sw qkv_b | o_w o_sw o_b | cq_w cq_sw cq_b | co_w co_sw co_b |
//   w1_w w1_sw w1_b | w2_w w2_sw w2_b | kc vc skc svc
// ws_ptrs (12): ckv sckv mask fin_g fin_b gen_w4 gen_b tgt_lut pe ys trace(optional) reserved
extern "C" int ot_cdecoder_plan_build(void* plan_dev, int n_layers, int B, int S, int cap, int vocab, int spc, int64_t ys_ld,
                                      const void* const* layer_ptrs, const void* const* ws_ptrs) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(plan_dev && layer_ptrs && ws_ptrs, "null argument");
  OT_REQUIRE(n_layers >= 1 && n_layers <= kMaxLayers, "1..8 decoder layers");
  OT_REQUIRE(B >= 1 && spc >= 1 && spc <= kCS, "1..8 sentences per cluster");
  OT_REQUIRE(S >= 1 && S <= kMaxKeys && cap >= 2 && cap <= kMaxKeys, "source length and cache capacity must be <= 96");
  OT_REQUIRE(vocab > 1 && (vocab + kGenVT - 1) / kGenVT <= kCS * 8 * 3, "at most 192 generator tiles of 32 vocabulary entries");
  CdPlan plan;
  memset(&plan, 0, sizeof(plan));
  CdHot& h = plan.hot;
  h.n_layers = n_layers; h.B = B; h.S = S; h.cap = cap; h.vocab = vocab; h.spc = spc;
  h.n_gen_tiles = (vocab + kGenVT - 1) / kGenVT;
  h.emb_scale = sqrtf(static_cast<float>(kD));
  for (int k = 0; k < 10; ++k) OT_REQUIRE(ws_ptrs[k] != nullptr || k == 6, "null workspace pointer");
  h.ckv = static_cast<const int8_t*>(ws_ptrs[0]); h.sckv = static_cast<const float*>(ws_ptrs[1]);
  h.mask = static_cast<const uint8_t*>(ws_ptrs[2]);
  h.fin_g = static_cast<const float*>(ws_ptrs[3]); h.fin_b = static_cast<const float*>(ws_ptrs[4]);
  h.gen_w4 = static_cast<const float*>(ws_ptrs[5]); h.gen_b = static_cast<const float*>(ws_ptrs[6]);
  h.tgt_lut = static_cast<const float*>(ws_ptrs[7]); h.pe = static_cast<const float*>(ws_ptrs[8]);
  h.ys = static_cast<int64_t*>(const_cast<void*>(ws_ptrs[9]));
  h.trace = static_cast<unsigned long long*>(const_cast<void*>(ws_ptrs[10]));
  h.ys_ld = ys_ld;
  int rc;
  for (int l = 0; l < n_layers; ++l) {
    const void* const* p = layer_ptrs + l * 28;
    for (int k = 0; k < 28; ++k) OT_REQUIRE(p[k] != nullptr, "null layer pointer");
    CdLayer& L = h.layer[l];
    auto f = [&](int k) { return static_cast<const float*>(p[k]); };
    for (int i = 0; i < 3; ++i) { L.ln_g[i] = f(2 * i); L.ln_b[i] = f(2 * i + 1); }
    for (int w = 0; w < 6; ++w) { L.sw[w] = f(7 + 3 * w); L.bias[w] = f(8 + 3 * w); }
    L.kc = static_cast<int8_t*>(const_cast<void*>(p[24])); L.vc = static_cast<int8_t*>(const_cast<void*>(p[25]));
    L.skc = static_cast<float*>(const_cast<void*>(p[26])); L.svc = static_cast<float*>(const_cast<void*>(p[27]));
    const int wn[6] = {3 * kD, kD, kD, kD, kFF, kD};
    const int wk[6] = {kD, kD, kD, kD, kD, kFF};
    const int box_rows[6] = {192, 64, 64, 64, 256, 256};     // weight rows of a ring chunk
    const int box_kb[6] = {1, 4, 4, 4, 1, 1};                // k-blocks of a ring chunk
    for (int w = 0; w < 6; ++w)
      if ((rc = get_tensor_map_kblocks(&plan.map_w[l][w], p[6 + 3 * w], wn[w], wk[w], wk[w], box_rows[w], box_kb[w]))) return rc;
  }
  // screening generator (ws_ptrs[11], optional): a device blob = 1024-byte header (float max_v ||w_v||_2) + fp16 weight [8 * 576][512]
  if (ws_ptrs[11] != nullptr && !(getenv("OT_CD_GEN_TC") && atoi(getenv("OT_CD_GEN_TC")) == 0)) {
    OT_REQUIRE(vocab <= kCS * kGenRows, "screening generator: vocabulary > 4608");
    float wnorm = 0.f;
    OT_CHECK_CUDA(cudaMemcpy(&wnorm, ws_ptrs[11], sizeof(float), cudaMemcpyDeviceToHost));
    OT_REQUIRE(wnorm > 0.f && wnorm < 1e30f, "screening generator: bad weight norm");
    if ((rc = get_tensor_map_kblocks(&plan.map_g, static_cast<const uint8_t*>(ws_ptrs[11]) + 1024, kCS * kGenRows, 2 * kD, 2 * kD, kGenRB, 1))) return rc;
    h.gen_tc = 1;
    h.gen_eps = wnorm * 1.001f * 0.001953125f;      // 2^-9
    h.gen_abs = wnorm * 1e-6f;
  }
  OT_CHECK_CUDA(cudaMemcpy(plan_dev, &plan, sizeof(plan), cudaMemcpyHostToDevice));
  {
    std::lock_guard<std::mutex> lock(g_plan_mu);
    g_plans[plan_dev] = CdPlanInfo{B, spc, cap};
  }
  return OT_OK;
}

// Runs greedy steps t0 .. t0+n_steps-1 for B sentences (ys[:, t0] must hold the current tokens; caches hold positions < t0).
extern "C" int ot_cdecoder_run(const void* plan_dev, int B, int spc, int t0, int n_steps, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(plan_dev && B >= 1 && spc >= 1 && spc <= kCS && t0 >= 0 && n_steps >= 0, "bad arguments");
  {
    std::lock_guard<std::mutex> lock(g_plan_mu);
    auto it = g_plans.find(plan_dev);
    OT_REQUIRE(it != g_plans.end(), "plan_dev was not built by ot_cdecoder_plan_build in this process");
    OT_REQUIRE(it->second.B == B && it->second.spc == spc, "B / spc differ from what the plan was built for");
    OT_REQUIRE(t0 + n_steps <= it->second.cap - 1, "t0 + n_steps exceeds the KV-cache capacity of the plan (cap - 1 greedy steps)");
  }
  if (n_steps == 0) return OT_OK;
  static bool attr_set[64] = {};          // cudaFuncSetAttribute applies per device
  int dev = 0;
  OT_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    OT_CHECK_CUDA(cudaFuncSetAttribute(cdecoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmTotal + 1024));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int n_clusters = (B + spc - 1) / spc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_clusters * kCS);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmTotal + 1024;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const CdPlan* plan = static_cast<const CdPlan*>(plan_dev);
  OT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cdecoder_kernel, plan, t0, n_steps));
  count_launch();
  return OT_OK;
}
