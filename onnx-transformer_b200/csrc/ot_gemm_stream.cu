// Encoder-size int8 GEMM (M >= 2048 rows, fault-free): the PERSISTENT, warp-specialised variant of ot_gemm_i8.cu.
//
// Why a second kernel.  At M = 65,536 (cfg3) the one-tile-per-CTA kernel spends almost all of its time in the epilogue: K = 512 is
// 16 MMAs (~1.4 us) per 128 x 256 tile while the requant epilogue is ~16 fp32 instructions per element (~2.2 us at full issue
// rate), the tile's TMA loads, MMAs and epilogue run back to back, and the accumulator is read from TMEM twice around a
// barrier.cluster (profiles/r1_ncu_gemm_enc_ffn1_q8_v3.txt: tensor pipe 5.5 % active).  Here
//   * one CTA per SM loops over its share of the tiles; warp 0 = TMA producer (3-stage ring of A 128x128 B + W 256x128 B k-blocks),
//     warp 1 = tcgen05.mma issuer, warps 2..17 = epilogue (4 per scheduler: one row x 64 columns per thread);
//   * TWO accumulators (2 x 256 TMEM columns = the whole TMEM): the MMAs of tile i+1 run under the epilogue of tile i;
//   * the requant epilogue computes y = fl(fl(float(acc)*sx)*sw)+b (ReLU) ONCE, writes it back over its accumulator columns with
//     tcgen05.st (TMEM as the per-thread stash: 128 values per thread do not fit in registers without unrolling 4k instructions),
//     and after the row maxima are known quantizes from that stash: no recomputation;
//   * the CTAs that share a quant group (N/256 tiles of the same 128 rows: 2 for q|k|v groups of 512, 8 for the FFN hidden layer)
//     form a cluster and exchange row maxima by st.async + mbarrier complete_tx (0.15 us vs 0.65 us for barrier.cluster, measured
//     in ot_cdecoder.cu); nobody waits at a cluster barrier in the steady state.
// Results are bit-identical to ot_gemm_i8.cu: same fp32 finish order; the quantization is rint(RN(y/s)) in both (a one-multiply fast
// pass + the true division for the elements within 2^-15 of a rounding boundary here, see quant_fast2_bits).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "ot_common.h"
#include "ot_ptx.cuh"
#include "ot_rowmath.cuh"

namespace ot {

int get_tensor_map(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols, bool swizzle128);
int get_tensor_map_sw(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols, int swizzle);
int launch_gemm_wres(const int8_t* A, int64_t lda, const int8_t* W, int64_t ldw, int M, int N, int K, const float* row_scale, const float* col_scale,
                     const float* bias, int relu, void* out, int64_t ldo, float* out_scale, int quant_group, cudaStream_t stream, int w4);

constexpr int kSBM = 128, kSBN = 256, kSBK = 128, kSStages = 3;
constexpr int kSEpiWarps = 16, kSEpiThreads = kSEpiWarps * 32, kSThreads = 64 + kSEpiThreads;
constexpr int kSColsPerWarp = kSBN / (kSEpiWarps / 4);           // 64 columns per epilogue thread
constexpr int kSA = kSBM * kSBK, kSB = kSBN * kSBK;             // bytes per stage
constexpr int kSMaxCl = 8;
constexpr int kSBarOff = kSStages * (kSA + kSB);                 // full[4] empty[4] tfull[2] tempty[2] xbar[2] + tmem slot
constexpr int kSColOff = kSBarOff + 256;                         // float [16 warps][2 buffers][cs 64 | bias 64]
constexpr int kSRowmaxOff = kSColOff + kSEpiWarps * 2 * 2 * kSColsPerWarp * 4;   // float rowmax[2][4*kSMaxCl][128]
constexpr int kSStageOff = (kSRowmaxOff + 1023) & ~1023;           // fp32 output: per-warp staging block, 4 KB x 16 warps (aliases rowmax; 1024-byte aligned: TMA swizzle)
constexpr int kSResBarOff = kSStageOff + 64 * 1024;             // fp32 output through TMA: uint64 resbar[16 warps][2 buffers]
constexpr int kSSmem = kSResBarOff + 256 + 1024;
static_assert(2 * 4 * kSMaxCl * 128 * 4 <= 64 * 1024 && kSEpiWarps * 4096 <= 64 * 1024 && kSSmem <= 232448, "shared memory budget");

typedef unsigned long long sf2;     // two fp32 values in a 64-bit register (lo = first)
__device__ __forceinline__ sf2 s_pack2(float lo, float hi) {
  sf2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 s_unpack2(sf2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ sf2 s_fma2(sf2 a, sf2 b, sf2 c) {
  sf2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

__device__ __forceinline__ void s_tma_store_2d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src_smem), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void s_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void s_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void s_bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void s_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- CTA pair (cta_group::2): one 256-row MMA over two SMs, each CTA holding 128 rows of A and HALF of the weight block
__device__ __forceinline__ void s_tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void s_tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void s_tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void s_mma_i8_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all MMAs issued so far, signalled on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void s_mma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are credited to a barrier of the pair's leader (a shared::cluster address)
__device__ __forceinline__ void s_tma_load_2d_2cta(uint32_t dst_smem, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void s_mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

struct StreamArgs {
  int M, N, K;
  const float* row_scale;
  const float* col_scale;
  const float* bias;
  const float* residual;
  int64_t ldr;
  int relu;
  void* out;
  int64_t ldo;
  float* out_scale;
  int cluster_n;   // CTAs per quant group (Q8), 1 for fp32 output
  unsigned long long* trace;   // profiling aid (OT_GEMM_STREAM_TRACE): CTA 0 stamps %globaltimer per tile and role
  // {-0,-0}, {1,1} as ARGUMENTS: ptxas rewrites fma(x, y, -0) -> mul and fma(x, 1, b) -> add when it can see the constants and then
  // contracts the mul / add pair into ONE FFMA2 (a different rounding), -fmad=false or not (see ot_gemm_wres.cu)
  unsigned long long neg0, one;
  int pair;        // fp32 output, K > 512: CTA pairs (cta_group::2); tmap_b then has a 128-row box (half a weight block per CTA)
  int tma_epi;     // fp32 output: residual in / result out through per-warp TMA boxes (tmap_r / tmap_o are valid)
};

// trace layout: [tile li < 32][slot]: 0 epi tile start, 1 accumulator ready, 2 pass 1 done, 3 exchange done, 4 pass 2 done (epilogue warp 0);
// 8 MMA: accumulator free, 9 MMA: all MMAs issued; 12 TMA: first k-block requested, 13 TMA: last k-block requested
__device__ __forceinline__ void wtrace(const StreamArgs& g, uint32_t li, int e, int slot) {   // every epilogue warp of CTA 0, first 8 tiles
  if (g.trace != nullptr && blockIdx.x == 0 && li < 8 && (threadIdx.x & 31) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g.trace[512 + (li * 16 + e) * 8 + slot] = t;
  }
}
__device__ __forceinline__ void strace(const StreamArgs& g, uint32_t li, int slot) {
  if (g.trace != nullptr && blockIdx.x == 0 && li < 32) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g.trace[li * 16 + slot] = t;
  }
}

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_async_f32(uint32_t addr, float v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(__float_as_uint(v)), "r"(mbar) : "memory");
}

// PAIR: the cta_group::2 form (a separate instantiation: a kernel that CONTAINS cta_group::2 instructions cannot be launched without clusters)
template <bool Q8, bool PAIR = false>
__global__ void __launch_bounds__(kSThreads, 1)
gemm_stream_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_r,
                   const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ StreamArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kSStages * kSA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSBarOff);
  // PAIR: a stage holds 128 rows of A + HALF a weight block (32 KB instead of 48): the same shared memory makes a ring of 4
  constexpr int NST = PAIR ? 4 : kSStages;
  constexpr int kStA = PAIR ? 32768 : kSA, kStB = PAIR ? 32768 : kSB;      // stage strides of the A / W rings
  if (PAIR) sB = smem + kSA;
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + NST;
  uint64_t* tfull_bar = bars + 2 * NST;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* xbar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 2);
  float* s_col = reinterpret_cast<float*>(smem + kSColOff);       // per-warp column parameters
  float* rowmax = reinterpret_cast<float*>(smem + kSRowmaxOff);   // [xb][4*kSMaxCl][128]
  uint64_t* resbar = reinterpret_cast<uint64_t*>(smem + kSResBarOff);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) / 32, 0);
  const int lane = threadIdx.x & 31;
  const int cn = g.cluster_n;
  const uint32_t rank = cn > 1 ? cluster_ctarank() : 0u;
  constexpr bool pair = PAIR;                                        // cluster of 2 = one MMA pair; cn == 1, fp32 output
  const uint32_t prank = pair ? cluster_ctarank() : 0u;
  const int unit = pair ? 2 : cn;
  const int cluster_id = blockIdx.x / unit, n_clusters = gridDim.x / unit;
  const int m_tiles = (g.M + kSBM - 1) / kSBM;
  const int groups = (g.N / kSBN) / cn;
  const int total = pair ? ((m_tiles + 1) / 2) * groups : m_tiles * groups;   // pair: a unit of work is TWO 128-row tiles of one column tile
  const int nkb = g.K / kSBK;

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp_idx == 1) {
    if (elect_one()) {
      for (int s = 0; s < NST; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(smem_u32(&tfull_bar[b]), 1);
        mbar_init(smem_u32(&tempty_bar[b]), pair ? 2 * kSEpiWarps : kSEpiWarps);   // pair: the leader's MMAs wait for both epilogues
        mbar_init(smem_u32(&xbar[b]), 1);
      }
      for (int i = 0; i < 2 * kSEpiWarps; ++i) mbar_init(smem_u32(&resbar[i]), 1);
      fence_mbar_init();
    }
    __syncwarp();
    if (pair) {
      s_tmem_alloc2(smem_u32(tmem_slot), 512);
      s_tmem_relinquish2();
    } else {
      tmem_alloc(smem_u32(tmem_slot), 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (cn > 1 || pair) cluster_sync_all();   // every CTA of the cluster has initialised its barriers before anyone stores into it
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t cnt = 0, li = 0;
      for (int it = cluster_id; it < total; it += n_clusters, ++li) {
        const int m_tile = pair ? 2 * (it / groups) + static_cast<int>(prank) : it / groups;
        const int n_tile = (it % groups) * cn + static_cast<int>(rank);
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const uint32_t s = cnt % NST, ph = (cnt / NST) & 1u;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[s]);
          if (pair) {
            // each CTA brings its 128 rows of A and its HALF of the 256-row weight block; all four boxes are credited to the leader's barrier
            const uint32_t lfb = mapa_shared(fb, 0);
            if (prank == 0) mbar_arrive_expect_tx(fb, 2 * (kSA + kSB / 2));
            s_tma_load_2d_2cta(smem_u32(sA + s * kStA), &tmap_a, lfb, kb * kSBK, m_tile * kSBM);
            s_tma_load_2d_2cta(smem_u32(sB + s * kStB), &tmap_b, lfb, kb * kSBK, n_tile * kSBN + static_cast<int>(prank) * (kSBN / 2));
            continue;
          }
          mbar_arrive_expect_tx(fb, kSA + kSB);
          tma_load_2d(smem_u32(sA + s * kSA), &tmap_a, fb, kb * kSBK, m_tile * kSBM);
          tma_load_2d(smem_u32(sB + s * kSB), &tmap_b, fb, kb * kSBK, n_tile * kSBN);
          if (kb == 0) strace(g, li, 12);
        }
        strace(g, li, 13);
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (pair && prank != 0) {
      // the leader issues for both CTAs
    } else if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_i8(kSBM, kSBN);
      constexpr uint32_t idesc2 = make_idesc_i8(2 * kSBM, kSBN);
      uint32_t cnt = 0, li = 0;
      for (int it = cluster_id; it < total; it += n_clusters, ++li) {
        const uint32_t buf = li & 1u;
        mbar_wait(smem_u32(&tempty_bar[buf]), ((li >> 1) & 1u) ^ 1u);     // the epilogue has drained this accumulator
        tc_fence_after();
        strace(g, li, 8);
        const uint32_t d_tmem = tmem_base + buf * kSBN;
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const uint32_t s = cnt % NST, ph = (cnt / NST) & 1u;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sA + s * kStA));
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(sB + s * kStB));
          if (pair) {
#pragma unroll
            for (int k = 0; k < kSBK / 32; ++k)
              s_mma_i8_ss_2cta(d_tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc2, (kb | k) != 0 ? 1u : 0u);
            s_mma_commit_2cta(smem_u32(&empty_bar[s]));
            continue;
          }
#pragma unroll
          for (int k = 0; k < kSBK / 32; ++k)
            mma_i8_ss(d_tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          mma_commit(smem_u32(&empty_bar[s]));
        }
        if (pair) s_mma_commit_2cta(smem_u32(&tfull_bar[buf]));
        else mma_commit(smem_u32(&tfull_bar[buf]));
        strace(g, li, 9);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    // 16 warps: warp e owns TMEM lanes [32*(warp_idx%4), +32) (hardware rule) = 32 rows, and column quarter cq = e/4 = 64 columns,
    // as 4 chunks of 16.  Four warps per scheduler hide the TMEM / shared-memory latencies of each other; nothing in a tile needs a
    // block barrier: the column parameters are staged per warp, the accumulator is handed back per warp.
    const int e = warp_idx - 2;
    const int quarter = warp_idx & 3;
    const int cq = e >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const bool has_bias = g.bias != nullptr, relu = g.relu != 0;
    float* wcol = s_col + e * (2 * 2 * kSColsPerWarp);               // [buf][cs 64 | bias 64], private to this warp
    uint32_t li = 0;
    for (int it = cluster_id; it < total; it += n_clusters, ++li) {
      const uint32_t buf = li & 1u, par = (li >> 1) & 1u;
      const int grp = it % groups, n_tile = grp * cn + static_cast<int>(rank);
      const int m_tile = pair ? 2 * (it / groups) + static_cast<int>(prank) : it / groups;
      const int row = m_tile * kSBM + row_in_tile;
      const bool row_ok = row < g.M;
      const int col0 = n_tile * kSBN + cq * kSColsPerWarp;           // first of this warp's 64 columns
      float* cs = wcol + buf * (2 * kSColsPerWarp);
      float* bs = cs + kSColsPerWarp;
      bool cols_bad;
      {
        const float c0 = g.col_scale ? __ldg(g.col_scale + col0 + lane) : 1.0f, c1 = g.col_scale ? __ldg(g.col_scale + col0 + 32 + lane) : 1.0f;
        const float b0 = g.bias ? __ldg(g.bias + col0 + lane) : 0.0f, b1 = g.bias ? __ldg(g.bias + col0 + 32 + lane) : 0.0f;
        cs[lane] = c0; cs[32 + lane] = c1; bs[lane] = b0; bs[32 + lane] = b1;
        cols_bad = __any_sync(0xffffffffu, !(fabsf(c0) < 3.0e38f) || !(fabsf(c1) < 3.0e38f) || !(fabsf(b0) < 3.0e38f) || !(fabsf(b1) < 3.0e38f));
        if (Q8 && cn > 1 && e == 0 && lane == 0) mbar_arrive_expect_tx(smem_u32(&xbar[buf]), static_cast<uint32_t>(cn) * 4u * 128u * 4u);
        __syncwarp();
      }
      const float sx = (g.row_scale && row_ok) ? __ldg(g.row_scale + row) : 1.0f;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * kSBN + cq * kSColsPerWarp;

      if (!Q8 && g.tma_epi) {
        // ---- fp32 output (+ residual) through TMA.  The LSU form below spends the tile in three phases that do not overlap (ablation,
        // profiles/r2_cdecoder_ab_experiments.txt: O-projection 37 us of compute path + 10 us of residual loads + 21 us of stores = 68 us):
        // 16 warps wait for their loads and push their stores at the same time.  Here every warp owns two 2 KB buffers [32 rows][64 B]
        // (64-byte swizzle = the layout a thread-per-row 16-byte access pattern reads and writes without bank conflicts): the residual
        // of a 16-column sub-block ARRIVES there by TMA two sub-blocks ahead, the row's thread adds its y in place, and the buffer LEAVES
        // by a TMA store -- no residual registers, no transposed read-back, no LSU global traffic, nothing waits for a store.
        const bool has_res = g.residual != nullptr;
        const bool plain = has_bias && !relu;
        const sf2 kNeg0 = g.neg0, kOne = g.one, sx2 = s_pack2(sx, sx);
        uint8_t* wbuf = smem + kSStageOff + e * 4096;
        uint64_t* rbar = resbar + 2 * e;
        const int row0 = m_tile * kSBM + quarter * 32;                // first row of this warp
        const bool has_next = it + n_clusters < total;
        int nrow0 = 0, ncol0 = 0;
        if (has_next) {
          const int it2 = it + n_clusters;
          nrow0 = (pair ? 2 * (it2 / groups) + static_cast<int>(prank) : it2 / groups) * kSBM + quarter * 32;
          ncol0 = ((it2 % groups) * cn + static_cast<int>(rank)) * kSBN + cq * kSColsPerWarp;
        }
        if (li == 0 && has_res && lane == 0) {
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            mbar_arrive_expect_tx(smem_u32(&rbar[b]), 2048);
            tma_load_2d(smem_u32(wbuf + b * 2048), &tmap_r, smem_u32(&rbar[b]), (col0 + 16 * b) * 4, row0);
          }
        }
        // (prefetch.global.L2 of the next tile's residual block, which helps the LSU form below, made this one 5 % slower: 66.7 vs 62.4 us)
        mbar_wait(smem_u32(&tfull_bar[buf]), par);
        tc_fence_after();
#pragma unroll
        for (int sb = 0; sb < 4; ++sb) {
          const uint32_t n = li * 4u + sb, b = sb & 1u;
          uint8_t* mybuf = wbuf + b * 2048 + lane * 64;
          uint32_t r[16];
          tmem_ld_32x16(taddr + 16 * sb, r);
          tmem_wait_ld();
          if (sb == 3) {
            // the whole accumulator slice is in registers: the MMAs of the tile after next may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (pair && prank != 0) s_mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[buf]), 0));
              else mbar_arrive(smem_u32(&tempty_bar[buf]));
            }
          }
          float y[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 c4 = *reinterpret_cast<const float4*>(cs + 16 * sb + 4 * j);
            const float4 b4 = *reinterpret_cast<const float4*>(bs + 16 * sb + 4 * j);
            const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
            if (plain) {
#pragma unroll
              for (int q = 0; q < 4; q += 2) {
                const sf2 a2 = s_pack2(__int2float_rn(static_cast<int>(r[4 * j + q])), __int2float_rn(static_cast<int>(r[4 * j + q + 1])));
                const float2 v2 = s_unpack2(s_fma2(s_fma2(s_fma2(a2, sx2, kNeg0), s_pack2(cv[q], cv[q + 1]), kNeg0), kOne, s_pack2(bv[q], bv[q + 1])));
                y[4 * j + q] = v2.x;
                y[4 * j + q + 1] = v2.y;
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float v = __fmul_rn(__fmul_rn(__int2float_rn(static_cast<int>(r[4 * j + q])), sx), cv[q]);
                const float vb = __fadd_rn(v, bv[q]);
                v = has_bias ? vb : v;
                const float vr = fmaxf(v, 0.0f);
                y[4 * j + q] = relu ? vr : v;
              }
            }
          }
          if (has_res) {
            mbar_wait(smem_u32(&rbar[b]), (n >> 1) & 1u);             // the residual of this sub-block has landed in the buffer
          } else if (n >= 2) {
            if (lane == 0) s_bulk_wait_read1();                       // the store that last used this buffer has read it
            __syncwarp();
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4* slot = reinterpret_cast<float4*>(mybuf + ((c ^ ((lane >> 1) & 3)) << 4));
            float4 v = make_float4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]);
            if (has_res) {
              const float4 rs = *slot;
              const float2 lo = s_unpack2(s_fma2(s_pack2(rs.x, rs.y), kOne, s_pack2(v.x, v.y)));      // res + y, the same fadd
              const float2 hi = s_unpack2(s_fma2(s_pack2(rs.z, rs.w), kOne, s_pack2(v.z, v.w)));
              v = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
            *slot = v;
          }
          fence_proxy_async_smem();                                   // generic-proxy writes of the buffer -> visible to the TMA store
          __syncwarp();
          if (lane == 0) {
            s_tma_store_2d(&tmap_o, smem_u32(wbuf + b * 2048), (col0 + 16 * sb) * 4, row0);
            s_bulk_commit();
            // the buffer's next residual: sub-block sb + 2 of this tile, or sub-block sb - 2 of the next one
            if (has_res && (sb < 2 || has_next)) {
              s_bulk_wait_read0();                                    // the store has read the buffer
              mbar_arrive_expect_tx(smem_u32(&rbar[b]), 2048);
              if (sb < 2) tma_load_2d(smem_u32(wbuf + b * 2048), &tmap_r, smem_u32(&rbar[b]), (col0 + 16 * (sb + 2)) * 4, row0);
              else tma_load_2d(smem_u32(wbuf + b * 2048), &tmap_r, smem_u32(&rbar[b]), (ncol0 + 16 * (sb - 2)) * 4, nrow0);
            }
          }
          __syncwarp();
        }
      } else if (!Q8) {
        // ---- fp32 output (+ residual).  These GEMMs are bound by the memory path of the epilogue, not by the MMAs: with one thread
        // per row, every 16-byte load / store instruction of a warp touched 32 different lines (profiles/r2_ncu_gemm_cfg3.txt: issue
        // 13 %, DRAM 32 %).  Here a warp stages 32 rows x 32 columns of y in shared memory (thread = row, 16-byte chunks XOR-swizzled by
        // the row) and reads them back transposed -- 8 lanes per row -- so that the residual loads and the output stores are full
        // 128-byte lines, 4 lines per instruction.  The residual is added in the transposed domain (same fadd, same operands).
        const bool has_res = g.residual != nullptr;
        const bool plain = has_bias && !relu;
        const sf2 kNeg0 = g.neg0, kOne = g.one, sx2 = s_pack2(sx, sx);
        uint8_t* wst = smem + kSStageOff + e * 4096;
        const int trow = lane >> 3, tch = lane & 7;                   // transposed role: row 4i + trow of the warp's 32, 16-byte chunk tch
        const int row0 = m_tile * kSBM + quarter * 32;                // first row of this warp
        float4 res[8];
        auto load_res = [&](int blk) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int grow = row0 + 4 * i + trow;
#if defined(OT_GS_ABL) && (OT_GS_ABL & 1)
            if (has_res && grow < 0) res[i] = __ldg(
#else
            if (has_res && grow < g.M) res[i] = __ldg(
#endif
                reinterpret_cast<const float4*>(g.residual + static_cast<int64_t>(grow) * g.ldr + col0 + 32 * blk) + tch);
          }
        };
        load_res(0);                                                  // in flight before the accumulator is waited for
#ifndef OT_GS_NOPF
        // the residual block this warp adds in its NEXT tile (32 rows x 256 bytes = 64 lines): on its way from HBM to L2 a whole tile
        // (~10 us) ahead, so that the next tile's load_res sees L2 latency instead of HBM latency
        if (has_res && g.K <= 512 && it + n_clusters < total) {   // (K = 2048: the prefetch competes with the 134 MB activation stream: +9 %)
          const int it2 = it + n_clusters;
          const int m2 = it2 / groups, n2 = (it2 % groups) * cn + static_cast<int>(rank);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int idx = lane + 32 * k;
            const int prow = m2 * kSBM + quarter * 32 + (idx >> 1);
            if (prow < g.M)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(g.residual + static_cast<int64_t>(prow) * g.ldr + n2 * kSBN + cq * kSColsPerWarp + 32 * (idx & 1)));
          }
        }
#endif
        mbar_wait(smem_u32(&tfull_bar[buf]), par);
        tc_fence_after();
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          uint32_t r[2][16];
          tmem_ld_32x16(taddr + 32 * blk, r[0]);
          tmem_ld_32x16(taddr + 32 * blk + 16, r[1]);
          tmem_wait_ld();
          if (blk == 1) {
            // the whole accumulator slice is in registers: the MMAs of the tile after next may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (pair && prank != 0) s_mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[buf]), 0));   // the leader's MMA warp owns both accumulators
              else mbar_arrive(smem_u32(&tempty_bar[buf]));
            }
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 c4 = *reinterpret_cast<const float4*>(cs + 32 * blk + 16 * c + 4 * j);
              const float4 b4 = *reinterpret_cast<const float4*>(bs + 32 * blk + 16 * c + 4 * j);
              const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
              float y[4];
              if (plain) {
                // bias, no ReLU (every fp32-output linear of the model): fl(fl(float(acc) * sx) * sw) + b, two columns per issue slot
#pragma unroll
                for (int b = 0; b < 4; b += 2) {
                  const sf2 a2 = s_pack2(__int2float_rn(static_cast<int>(r[c][4 * j + b])), __int2float_rn(static_cast<int>(r[c][4 * j + b + 1])));
                  const float2 v2 = s_unpack2(s_fma2(s_fma2(s_fma2(a2, sx2, kNeg0), s_pack2(cv[b], cv[b + 1]), kNeg0), kOne, s_pack2(bv[b], bv[b + 1])));
                  y[b] = v2.x;
                  y[b + 1] = v2.y;
                }
              } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                  float v = __fmul_rn(__fmul_rn(__int2float_rn(static_cast<int>(r[c][4 * j + b])), sx), cv[b]);
                  const float vb = __fadd_rn(v, bv[b]);
                  v = has_bias ? vb : v;
                  const float vr = fmaxf(v, 0.0f);
                  y[b] = relu ? vr : v;
                }
              }
              *reinterpret_cast<float4*>(wst + lane * 128 + (((4 * c + j) ^ (lane & 7)) << 4)) = make_float4(y[0], y[1], y[2], y[3]);
            }
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + trow, grow = row0 + rr;
            float4 v = *reinterpret_cast<const float4*>(wst + rr * 128 + ((tch ^ (rr & 7)) << 4));
#if defined(OT_GS_ABL) && (OT_GS_ABL & 2)
            if (grow < g.M && g.ldo < 0) {
#else
            if (grow < g.M) {
#endif
              if (has_res) {
                const float2 lo = s_unpack2(s_fma2(s_pack2(res[i].x, res[i].y), kOne, s_pack2(v.x, v.y)));      // res + y, the same fadd
                const float2 hi = s_unpack2(s_fma2(s_pack2(res[i].z, res[i].w), kOne, s_pack2(v.z, v.w)));
                v = make_float4(lo.x, lo.y, hi.x, hi.y);
              }
              reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + static_cast<int64_t>(grow) * g.ldo + col0 + 32 * blk)[tch] = v;
            }
          }
          if (blk == 0) load_res(1);
          __syncwarp();                                               // the staging block is rewritten by the next column block / tile
        }
      } else {
        if (e == 0 && lane == 0) strace(g, li, 0);
        wtrace(g, li, e, 0);
        mbar_wait(smem_u32(&tfull_bar[buf]), par);
        tc_fence_after();
        if (e == 0 && lane == 0) strace(g, li, 1);
        wtrace(g, li, e, 1);
        // ---- pass 1: y once, stashed over its own accumulator columns; row abs-max of this thread's 64 columns
        float amax = 0.f;
        {
          uint32_t r[4][16];
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld_32x16(taddr + 16 * c, r[c]);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 c4 = *reinterpret_cast<const float4*>(cs + 16 * c + 4 * j);
              const float4 b4 = *reinterpret_cast<const float4*>(bs + 16 * c + 4 * j);
              const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                float v = __fmul_rn(__fmul_rn(__int2float_rn(static_cast<int>(r[c][4 * j + b])), sx), cv[b]);
                const float vb = __fadd_rn(v, bv[b]);
                v = has_bias ? vb : v;
                const float vr = fmaxf(v, 0.0f);
                v = relu ? vr : v;
                amax = fmaxf(amax, fabsf(v));
                r[c][4 * j + b] = __float_as_uint(v);
              }
            }
            tmem_st_32x16(taddr + 16 * c, r[c]);
          }
        }
        if (!row_ok) amax = 0.f;
        if (e == 0 && lane == 0) strace(g, li, 2);
        wtrace(g, li, e, 2);
        // ---- exchange: this (CTA, column quarter)'s row maximum to every CTA of the quant group, data-flow synchronised
        float* xrow = rowmax + (buf * 4 * kSMaxCl) * 128 + row_in_tile;
        {
          const uint32_t slot = smem_u32(xrow + (rank * 4 + cq) * 128);
          const uint32_t xb = smem_u32(&xbar[buf]);
          if (cn > 1) {
            for (int peer = 0; peer < cn; ++peer) st_async_f32(mapa_shared(slot, peer), amax, mapa_shared(xb, peer));
            mbar_wait(xb, par);
          } else {
            xrow[cq * 128] = amax;                   // a quant group of one tile: the four column quarters meet at a block barrier
            asm volatile("bar.sync 1, %0;" ::"n"(kSEpiThreads) : "memory");
          }
        }
        if (e == 0 && lane == 0) strace(g, li, 3);
        wtrace(g, li, e, 3);
        float rmax = 0.f;
        for (int p = 0; p < 4 * cn; ++p) rmax = fmaxf(rmax, xrow[p * 128]);
        const float s = __fdiv_rn(fmaxf(rmax, 1e-5f), 127.0f);
        const float s_rcp = __frcp_rn(s);
        if (row_ok && cq == 0 && rank == 0) g.out_scale[static_cast<int64_t>(row) * groups + grp] = s;
        // rows whose inputs are not plain finite numbers take the exact division for every element (same results as ot_gemm_i8.cu)
        const bool row_slow = !(rmax < 3.0e38f) || !(fabsf(sx) < 3.0e38f) || cols_bad;
        // ---- pass 2: quantize from the stash
        int8_t* orow = reinterpret_cast<int8_t*>(g.out) + static_cast<int64_t>(row) * g.ldo + col0;
        {
          tmem_wait_st();
          uint32_t r[4][16];
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld_32x16(taddr + 16 * c, r[c]);
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));     // the stash is in registers: hand the accumulator back
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t tb[16];
            float yv[16];
            bool slow = row_slow;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              yv[j] = __uint_as_float(r[c][j]);
              tb[j] = quant_fast2_bits(yv[j], s_rcp, slow);
            }
            if (slow) quant_redo_chunk2<16>(yv, s, s_rcp, row_slow, tb);    // rare: only the flagged elements take the division
            uint32_t packed[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              packed[j] = __byte_perm(__byte_perm(tb[4 * j], tb[4 * j + 1], 0x0040), __byte_perm(tb[4 * j + 2], tb[4 * j + 3], 0x0040), 0x5410);
            if (row_ok) *reinterpret_cast<uint4*>(orow + 16 * c) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          }
        }
        if (e == 0 && lane == 0) strace(g, li, 4);
        wtrace(g, li, e, 4);
        if (cn == 1) asm volatile("bar.sync 1, %0;" ::"n"(kSEpiThreads) : "memory");   // the rowmax slots are rewritten two tiles later, but the
                                                                                          // block barrier above must not be overtaken by a fast warp
      }
    }
  }

  if (!Q8 && g.tma_epi && warp_idx >= 2 && lane == 0) s_bulk_wait0();     // this warp's TMA stores have left shared memory and are complete
  tc_fence_before();
  if (cn > 1 || pair) cluster_sync_all();
  else __syncthreads();
  if (warp_idx == 1) {
    if (pair) s_tmem_dealloc2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// Launch when the problem qualifies; returns 1 if it does not (caller falls back to gemm_i8_kernel), 0 on success, < 0 on error.
int launch_gemm_stream(const int8_t* A, int64_t lda, const int8_t* W, int64_t ldw, int M, int N, int K, const float* row_scale,
                       const float* col_scale, const float* bias, const float* residual, int64_t ldr, int relu, int out_kind, void* out,
                       int64_t ldo, float* out_scale, int quant_group, cudaStream_t stream) {
  static const int min_m = getenv("OT_GEMM_STREAM_MIN_M") ? atoi(getenv("OT_GEMM_STREAM_MIN_M")) : 2048;
  if (M < min_m || N % kSBN != 0 || K % kSBK != 0 || K < kSBK) return 1;
  if (out_kind != OT_OUT_F32 && out_kind != OT_OUT_Q8) return 1;
  if (out_kind == OT_OUT_Q8 && residual == nullptr) {     // K <= 512: the weight-stationary kernel (ot_gemm_wres.cu)
    const int rc = launch_gemm_wres(A, lda, W, ldw, M, N, K, row_scale, col_scale, bias, relu, out, ldo, out_scale, quant_group, stream, 0);
    if (rc != 1) return rc;
  }
  int cn = 1;
  if (out_kind == OT_OUT_Q8) {
    if (residual != nullptr || quant_group % kSBN != 0 || quant_group / kSBN > kSMaxCl || N % quant_group != 0) return 1;
    cn = quant_group / kSBN;
    if ((cn & (cn - 1)) != 0) return 1;
  }
  CUtensorMap ta, tb;
  int rc = get_tensor_map(&tb, W, N, K, ldw, kSBN, kSBK, true);
  if (rc) return rc;
  rc = get_tensor_map(&ta, A, M, K, lda, kSBM, kSBK, true);
  if (rc) return rc;
  StreamArgs g = {};
  g.neg0 = 0x8000000080000000ull; g.one = 0x3F8000003F800000ull;
  g.M = M; g.N = N; g.K = K;
  g.row_scale = row_scale; g.col_scale = col_scale; g.bias = bias; g.residual = residual; g.ldr = ldr;
  g.relu = relu; g.out = out; g.ldo = ldo; g.out_scale = out_scale; g.cluster_n = cn;
  CUtensorMap tr_map = ta, to_map = ta;
  {
    static const int tma_epi = getenv("OT_GEMM_STREAM_TMA_EPI") ? atoi(getenv("OT_GEMM_STREAM_TMA_EPI")) : 1;
    const bool aligned = (reinterpret_cast<uintptr_t>(out) % 16 == 0) && (ldo % 4 == 0) &&
                         (residual == nullptr || ((reinterpret_cast<uintptr_t>(residual) % 16 == 0) && (ldr % 4 == 0)));
    // K > 512 (FFN2) keeps the LSU epilogue: its tiles are bound by the TMA operand stream (768 KB per tile), and epilogue boxes queued
    // behind it made the tile 4 % slower (93.7 vs 89.7 us); K <= 512 (O-projection, cross-K/V): 68.7 -> 60.9 us
    if (tma_epi && out_kind == OT_OUT_F32 && aligned && (K <= 512 || tma_epi > 1)) {
      rc = get_tensor_map_sw(&to_map, out, M, static_cast<uint64_t>(N) * 4, static_cast<uint64_t>(ldo) * 4, 32, 64, 2);
      if (rc) return rc;
      if (residual != nullptr) {
        rc = get_tensor_map_sw(&tr_map, residual, M, static_cast<uint64_t>(N) * 4, static_cast<uint64_t>(ldr) * 4, 32, 64, 2);
        if (rc) return rc;
      }
      g.tma_epi = 1;
    }
  }
  {
    // K > 512 (FFN2): the tile is bound by the L2 -> SM operand stream (A 256 KB + W 512 KB per 128 x 256 tile at ~43 B/clk per SM).  A CTA
    // pair shares the weight block (each CTA brings half of it, one cta_group::2 MMA of M = 256 reads both halves): 640 instead of 768 KB
    static const int pair_env = getenv("OT_GEMM_STREAM_PAIR") ? atoi(getenv("OT_GEMM_STREAM_PAIR")) : 1;
    if (pair_env && out_kind == OT_OUT_F32 && (K > 512 || pair_env > 1) && M >= 2 * kSBM * 8) {
      rc = get_tensor_map(&tb, W, N, K, ldw, kSBN / 2, kSBK, true);
      if (rc) return rc;
      g.pair = 1;
    }
  }
  if (const char* tr = getenv("OT_GEMM_STREAM_TRACE")) g.trace = reinterpret_cast<unsigned long long*>(strtoull(tr, nullptr, 16));
  auto kernel = out_kind == OT_OUT_Q8 ? gemm_stream_kernel<true> : (g.pair ? gemm_stream_kernel<false, true> : gemm_stream_kernel<false>);
  static DeviceOnce attr_set[3];
  static int max_clusters[2][kSMaxCl + 1] = {};
  const int ki = out_kind == OT_OUT_Q8 ? 1 : 0;
  if (attr_set[g.pair ? 2 : ki].need()) OT_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSSmem));
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int n_cl;
  if (g.pair) {
    n_cl = sms / 2;      // CTA pairs; launched as clusters of 2 below
  } else if (cn == 1) {
    n_cl = sms;
  } else {
    if (max_clusters[ki][cn] == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(static_cast<unsigned>(sms / cn * cn));
      cfg.blockDim = dim3(kSThreads);
      cfg.dynamicSmemBytes = kSSmem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cn; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = sms / cn / 2 > 0 ? sms / cn / 2 : 1;
      }
      max_clusters[ki][cn] = n;
    }
    n_cl = max_clusters[ki][cn];
  }
  const int m_tiles = (M + kSBM - 1) / kSBM;
  const int total = g.pair ? ((m_tiles + 1) / 2) * (N / kSBN) : m_tiles * ((N / kSBN) / cn);
  if (n_cl > total) n_cl = total;
  const int unit = g.pair ? 2 : cn;
  OT_CHECK_CUDA(launch_kernel(kernel, dim3(static_cast<unsigned>(n_cl * unit)), dim3(kSThreads), kSSmem, stream, unit, ta, tb, tr_map, to_map, g));
  count_launch();
  return OT_OK;
}

}  // namespace ot
