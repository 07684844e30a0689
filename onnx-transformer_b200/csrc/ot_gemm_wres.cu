// Encoder-size int8 GEMM with the requantization fused (out = int8 + per-row-group scales), K <= 512: the WEIGHT-STATIONARY variant.
//
// Why a third kernel.  ot_gemm_stream.cu re-loads a 256 x K weight tile (128 KB at K = 512) with every 128-row activation tile: 192 KB
// of TMA traffic per tile per SM, ~2.3 us at the ~42 B/clk an SM gets from L2 -- more than the 16 MMAs (1.1-1.4 us) or a lean
// epilogue need.  Its epilogue also went through TMEM twice and wrote the int8 tile with 16-byte stores 1.5 KB apart (2048 half-used
// sectors per tile; profiles/r2_gemm_stream_trace_qkv.txt: pass 2 = 3.3 us of a 6.4 us tile).  Here
//   * a CTA keeps ONE 256-column weight tile resident in shared memory for the whole launch (128 KB, loaded once) and walks down the
//     activation rows: per tile only the 128 x K activation block (64 KB) moves, through a 3-stage ring of k-blocks;
//   * warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (two TMEM accumulators: the MMAs of tile i+1 run under the epilogue of
//     tile i), warps 2..17 = epilogue, thread = (row, 64 columns);
//   * the epilogue reads its accumulator slice ONCE: y = fl(fl(float(acc)*sx)*sw)+b (ReLU) stays in 64 registers, the accumulator is
//     handed back to the MMA warp at once; the row abs-max is reduced over the CTA's four column quarters in shared memory and
//     exchanged with the other CTAs of the quant group (cluster of N_group/256 <= 8) by st.async + mbarrier complete_tx;
//   * the quantized tile is staged in shared memory in the 128-byte-swizzled layout and leaves as two TMA tensor stores
//     (cp.async.bulk.tensor ... global.shared::cta): full 128-byte lines, no LSU traffic, the next tile's pass 1 runs under the store.
// Results are bit-identical to ot_gemm_i8.cu / ot_gemm_stream.cu: same fp32 finish order, quantization = rint(RN(y/s)) with the
// quotient computed exactly by FMA residual steps (no division, no redo path; tools/check_div_exact.c).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "ot_common.h"
#include "ot_ptx.cuh"
#include "ot_rowmath.cuh"

namespace ot {

int get_tensor_map(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols, bool swizzle128);

namespace wres {

constexpr int kBM = 128, kBN = 256, kBK = 128, kStages = 3, kMaxKb = 4;
constexpr int kEpiWarps = 16, kEpiThreads = kEpiWarps * 32, kThreads = 64 + kEpiThreads + 32;   // + warp 18: TMA stores
constexpr int kA = kBM * kBK;                       // bytes of one activation k-block (ring stage)
constexpr int kWkb = kBN * kBK;                     // bytes of one weight k-block
constexpr int kMaxCl = 8;
constexpr int kOffW = 0;                            // resident weight tile: kMaxKb k-blocks of [256 rows][128 B]
constexpr int kOffA = kOffW + kMaxKb * kWkb;        // 131072
constexpr int kOffStage = kOffA + kStages * kA;     // 180224: quantized tile, two [128 rows][128 B] halves, swizzled (TMA store source)
constexpr int kOffCol = kOffStage + kBM * kBN;      // 212992: float cs[256] | bias[256]
constexpr int kOffPm = kOffCol + 2 * kBN * 4;       // 215040: float pm[2][4 column quarters][128 rows]
constexpr int kOffRowmax = kOffPm + 2 * 4 * kBM * 4;   // 219136: float rowmax[2][kMaxCl][128]
constexpr int kOffBar = kOffRowmax + 2 * kMaxCl * kBM * 4;   // 227328: wfull, full[3], empty[3], tfull[2], tempty[2], xbar[2][4], sfull, sempty + tmem slot
constexpr int kSmem = kOffBar + 256 + 1024;
static_assert(kSmem <= 232448, "shared memory budget");

struct Args {
  int M, N, K;
  const float* row_scale;
  const float* col_scale;
  const float* bias;
  int relu;
  float* out_scale;
  int cluster_n;   // CTAs per quant group
  int groups;      // quant groups per row = N / (256 * cluster_n)
  unsigned long long* trace;   // profiling aid (OT_GEMM_WRES_TRACE): CTA 0 stamps %globaltimer per tile, epilogue warp and phase
  // {-0,-0}, {1,1}, {1.5*2^23 x 2}: passed as ARGUMENTS so that ptxas cannot see their values -- it rewrites fma(x, y, -0) ->
  // mul and fma(x, 1, b) -> add and then contracts the mul/add pair into ONE FFMA2 (a different rounding), -fmad=false or not
  unsigned long long neg0, one, magic;
};
// trace[(tile < 16) * 16 * 8 + e * 8 + slot]: 0 tile start, 1 accumulator ready, 2 pass 1 done, 3 past the quarter barrier, 4 exchange done,
// 5 pass 2 done, 6 past the staging barrier
__device__ __forceinline__ void wtrace(const Args& g, int tile, int e, int slot) {
  if (g.trace != nullptr && blockIdx.x == 0 && tile < 16 && (threadIdx.x & 31) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g.trace[(tile * 16 + e) * 8 + slot] = t;
  }
}

__device__ __forceinline__ void st_async_f32(uint32_t addr, float v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(__float_as_uint(v)), "r"(mbar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src_smem), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

typedef unsigned long long f2;     // two fp32 values in a 64-bit register (lo = first)
__device__ __forceinline__ f2 pack2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 unpack2(f2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// W4 (cfg4): the weight tile arrives as PACKED int4 (two values per byte, even k in the low nibble: 64 KB instead of 128 KB from HBM / L2),
// is unpacked ONCE into the resident swizzled int8 tile by the epilogue warps, and the launch then runs exactly like the int8 one -- the
// place where 4-bit weights cost nothing at encoder sizes (a separate instantiation: the int8 kernel's register allocation is untouched).
template <bool RELU, bool W4 = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_wres_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_o,
                 const __grid_constant__ Args g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sW = smem + kOffW;
  uint8_t* sA = smem + kOffA;
  uint8_t* stage = smem + kOffStage;
  float* s_cs = reinterpret_cast<float*>(smem + kOffCol);
  float* s_bs = s_cs + kBN;
  float* pm = reinterpret_cast<float*>(smem + kOffPm);
  float* rowmax = reinterpret_cast<float*>(smem + kOffRowmax);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* wfull = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* xbar = tempty_bar + 2;            // [buf][row quarter]
  uint64_t* sfull = xbar + 8;                 // staging tile written (16 epilogue warps)
  uint64_t* sempty = sfull + 1;               // staging tile read by the TMA stores
  uint64_t* wready = sempty + 1;              // W4: the weight tile has been unpacked (16 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wready + 1);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) / 32, 0);
  const int lane = threadIdx.x & 31;
  const int cn = g.cluster_n;
  const uint32_t rank = cn > 1 ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / cn, n_clusters = gridDim.x / cn;
  const int m_tiles = (g.M + kBM - 1) / kBM;
  const int nkb = g.K / kBK;
  // this cluster's quant group (fixed for the launch) and its share of the row tiles: clusters gi, gi + groups, ... serve group gi
  const int gi = cluster_id % g.groups, ci = cluster_id / g.groups;
  const int ng = (n_clusters - gi + g.groups - 1) / g.groups;
  const int n_tile = gi * cn + static_cast<int>(rank);

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_o);
  }
  if (warp_idx == 1) {
    if (elect_one()) {
      mbar_init(smem_u32(wfull), 1);
      mbar_init(smem_u32(wready), kEpiWarps);
      for (int s = 0; s < kStages; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(smem_u32(&tfull_bar[b]), 1);
        mbar_init(smem_u32(&tempty_bar[b]), kEpiWarps);
        for (int q = 0; q < 4; ++q) mbar_init(smem_u32(&xbar[b * 4 + q]), 1);
      }
      mbar_init(smem_u32(sfull), kEpiWarps);
      mbar_init(smem_u32(sempty), 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  if (warp_idx >= 2 && warp_idx < 2 + kEpiWarps) {        // column parameters of this CTA's 256 columns: constant for the launch
    const int t = threadIdx.x - 64;
    if (t < kBN) {
      const int col = n_tile * kBN + t;
      s_cs[t] = g.col_scale ? __ldg(g.col_scale + col) : 1.0f;
      s_bs[t] = g.bias ? __ldg(g.bias + col) : 0.0f;
    }
  }
  tc_fence_before();
  if (cn > 1) cluster_sync_all();   // every CTA of the cluster has initialised its barriers before anyone stores into it
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (g.trace != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g.trace[16 * 16 * 8 + 2 * blockIdx.x] = t;
  }
  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const uint32_t wb = smem_u32(wfull);
      if (W4) {
        // packed k-blocks [256 rows][64 B] into the (still idle) activation ring + staging tile: 16 KB each, unpacked by the epilogue warps
        mbar_arrive_expect_tx(wb, static_cast<uint32_t>(nkb) * (kWkb / 2));
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(smem_u32(sA + kb * (kWkb / 2)), &tmap_w, wb, kb * (kBK / 2), n_tile * kBN);
        mbar_wait(smem_u32(wready), 0);           // the ring is free again
      } else {
        mbar_arrive_expect_tx(wb, static_cast<uint32_t>(nkb) * kWkb);
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(smem_u32(sW + kb * kWkb), &tmap_w, wb, kb * kBK, n_tile * kBN);
      }
      uint32_t cnt = 0;
      for (int i = 0; i < (m_tiles - ci + ng - 1) / ng; ++i) {
        const int m_tile = ci + i * ng;
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const uint32_t s = cnt % kStages, ph = (cnt / kStages) & 1u;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_arrive_expect_tx(fb, kA);
          tma_load_2d(smem_u32(sA + s * kA), &tmap_a, fb, kb * kBK, m_tile * kBM);
        }
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_i8(kBM, kBN);
      mbar_wait(smem_u32(W4 ? wready : wfull), 0);
      uint32_t cnt = 0;
      for (int i = 0; i < (m_tiles - ci + ng - 1) / ng; ++i) {
        const uint32_t buf = i & 1u;
        mbar_wait(smem_u32(&tempty_bar[buf]), ((i >> 1) & 1u) ^ 1u);     // the epilogue has read this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kBN;
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const uint32_t s = cnt % kStages, ph = (cnt / kStages) & 1u;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sA + s * kA));
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(sW + kb * kWkb));
#pragma unroll
          for (int k = 0; k < kBK / 32; ++k)
            mma_i8_ss(d_tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          mma_commit(smem_u32(&empty_bar[s]));
        }
        mma_commit(smem_u32(&tfull_bar[buf]));
      }
    }
    __syncwarp();
  } else if (warp_idx == 2 + kEpiWarps) {
    // ===================== TMA store warp =====================
    // One lane: a TMA instruction costs its issuing warp ~0.3 us, which every epilogue warp waited for at the next barrier when warp 2
    // issued the stores itself (profiles/r2_gemm_wres_trace_*.txt); bulk async-groups are per thread, so issue and wait stay together.
    if (lane == 0) {
      const int n_my = (m_tiles - ci + ng - 1) / ng;
      for (int i = 0; i < n_my; ++i) {
        const int m_tile = ci + i * ng;
        mbar_wait(smem_u32(sfull), i & 1u);
        tma_store_2d(&tmap_o, smem_u32(stage), n_tile * kBN, m_tile * kBM);
        tma_store_2d(&tmap_o, smem_u32(stage + kBM * 128), n_tile * kBN + 128, m_tile * kBM);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(smem_u32(sempty));
      }
      bulk_wait0();
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    // warp e owns TMEM lanes [32*(warp_idx%4), +32) (hardware rule) = 32 rows, and column quarter cq = e/4 = 64 columns
    const int e = warp_idx - 2;
    const int quarter = warp_idx & 3;
    const int cq = e >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const float* cs = s_cs + cq * 64;
    const float* bs = s_bs + cq * 64;
    bool cols_bad = false;
    for (int j = lane; j < 64; j += 32) cols_bad = cols_bad || !(fabsf(cs[j]) < 3.0e38f) || !(fabsf(bs[j]) < 3.0e38f);
    cols_bad = __any_sync(0xffffffffu, cols_bad);
    if (W4) {
      // ---- unpack the packed weight tile once: task = (k-block, row, 16-byte chunk of the int8 row) <- 8 packed bytes
      mbar_wait(smem_u32(wfull), 0);
      const int te = e * 32 + lane;
      for (int task = te; task < nkb * 2048; task += kEpiThreads) {
        const int kb = task >> 11, row = (task >> 3) & 255, c = task & 7;
        const uint2 pk = *reinterpret_cast<const uint2*>(sA + kb * (kWkb / 2) + row * 64 + c * 8);
        uint32_t o[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t w = h ? pk.y : pk.x;
          uint32_t lo = w & 0x0F0F0F0Fu, hi = (w >> 4) & 0x0F0F0F0Fu;               // even-k / odd-k nibbles of 4 packed bytes
          lo |= (lo & 0x08080808u) * 0x1Eu;                                          // sign extension per byte: bit 3 -> bits 4..7
          hi |= (hi & 0x08080808u) * 0x1Eu;
          o[2 * h] = __byte_perm(lo, hi, 0x5140);                                    // k = 0, 1, 2, 3 of this word
          o[2 * h + 1] = __byte_perm(lo, hi, 0x7362);                                // k = 4 .. 7
        }
        *reinterpret_cast<uint4*>(sW + kb * kWkb + row * 128 + ((c ^ (row & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      fence_proxy_async_smem();                   // generic-proxy writes of the weight tile -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(wready));
    }
    // staging address of this thread's 64 bytes: half cq>>1, row, 16-byte chunks 4*(cq&1) .. +3 XOR-swizzled by the row (the layout
    // CU_TENSOR_MAP_SWIZZLE_128B expects of a [128][128 B] box)
    uint8_t* srow = stage + (cq >> 1) * (kBM * 128) + row_in_tile * 128;
    const int ch0 = 4 * (cq & 1), sw = row_in_tile & 7;
    const int n_my = (m_tiles - ci + ng - 1) / ng;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + cq * 64;
    // r holds the thread's 64 accumulator values of the CURRENT tile (then y, in place).  The accumulator of tile i+1 is fetched
    // chunk by chunk during pass 2 of tile i, into the registers pass 2 has just consumed: tcgen05.ld moves ~64 B/clk per SM (1 us per
    // 128 x 256 tile) and would otherwise sit in front of pass 1 with the FP32 pipe idle.  (Fetching half of it during pass 1 instead
    // was slower: pass 1 is short and then waits for the load.)
    uint32_t r[4][16];
    if (n_my > 0) {
      mbar_wait(smem_u32(&tfull_bar[0]), 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x16(tlane + 16 * c, r[c]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[0]));           // the accumulator slice is in registers: back to the MMA warp
    }
    float sx_next = 1.0f;                                             // row scale of the NEXT tile: an L2 / HBM round trip ahead of its use
    if (n_my > 0 && g.row_scale && ci * kBM + row_in_tile < g.M) sx_next = __ldg(g.row_scale + ci * kBM + row_in_tile);
    const f2 kNeg0 = g.neg0, kOne = g.one, kMagic = g.magic;
    for (int i = 0; i < n_my; ++i) {
      const uint32_t buf = i & 1u, par = (i >> 1) & 1u;
      const int m_tile = ci + i * ng;
      const int row = m_tile * kBM + row_in_tile;
      const bool row_ok = row < g.M;
      if (cn > 1 && cq == 0 && lane == 0) mbar_arrive_expect_tx(smem_u32(&xbar[buf * 4 + quarter]), static_cast<uint32_t>(cn) * 32u * 4u);
      const float sx = sx_next;
      {
        const int nrow = row + ng * kBM;
        sx_next = (g.row_scale && i + 1 < n_my && nrow < g.M) ? __ldg(g.row_scale + nrow) : 1.0f;
      }
      wtrace(g, i, e, 1);
      // ---- pass 1: y = fl(fl(float(acc)*sx)*sw)+b (ReLU), in place, and the row abs-max of this thread's 64 columns.  Two columns
      // per FFMA2 (fma.rn.f32x2 = two IEEE fp32 FMAs in one issue slot): x*y = fma(x, y, -0) and x+y = fma(x, 1, y) bit for bit.
      // (mul.rn.f32x2 followed by add.rn.f32x2 is NOT an option: ptxas contracts the pair into one FFMA2, even with -fmad=false.)
      float amax = 0.f;
      {
        const f2 sx2 = pack2(sx, sx);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const ulonglong2 c4 = *reinterpret_cast<const ulonglong2*>(cs + 16 * c + 4 * j);
            const ulonglong2 b4 = *reinterpret_cast<const ulonglong2*>(bs + 16 * c + 4 * j);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const f2 a2 = pack2(__int2float_rn(static_cast<int>(r[c][4 * j + 2 * h])), __int2float_rn(static_cast<int>(r[c][4 * j + 2 * h + 1])));
              const f2 v2 = fma2(fma2(fma2(a2, sx2, kNeg0), h ? c4.y : c4.x, kNeg0), kOne, h ? b4.y : b4.x);
              float2 v = unpack2(v2);
              if (RELU) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); }
              amax = fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y)));
              r[c][4 * j + 2 * h] = __float_as_uint(v.x);
              r[c][4 * j + 2 * h + 1] = __float_as_uint(v.y);
            }
          }
        }
      }
      if (!row_ok) amax = 0.f;
      wtrace(g, i, e, 2);
      // ---- row maximum over the CTA's four column quarters, then over the CTAs of the quant group
      // (the 4 warps of a row quarter meet at their own named barrier; the quarters, one per SM sub-partition, never wait for each other)
      float* pmb = pm + buf * (4 * kBM);
      pmb[cq * kBM + row_in_tile] = amax;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
      wtrace(g, i, e, 3);
      float rmax = fmaxf(fmaxf(pmb[row_in_tile], pmb[kBM + row_in_tile]), fmaxf(pmb[2 * kBM + row_in_tile], pmb[3 * kBM + row_in_tile]));
      if (cn > 1) {
        float* xrow = rowmax + (buf * kMaxCl) * kBM + row_in_tile;
        const uint32_t xb = smem_u32(&xbar[buf * 4 + quarter]);
        if (cq == 0) {
          const uint32_t slot = smem_u32(xrow + rank * kBM);
          for (int peer = 0; peer < cn; ++peer) st_async_f32(mapa_shared(slot, peer), rmax, mapa_shared(xb, peer));
        }
        mbar_wait(xb, par);
        rmax = 0.f;
        for (int p = 0; p < cn; ++p) rmax = fmaxf(rmax, xrow[p * kBM]);
      }
      wtrace(g, i, e, 4);
      const float s = __fdiv_rn(fmaxf(rmax, 1e-5f), 127.0f);
      const float s_rcp = __frcp_rn(s);
      if (row_ok && cq == 0 && rank == 0) g.out_scale[static_cast<int64_t>(row) * g.groups + gi] = s;
      // rows whose inputs are not plain finite numbers take the exact division for every element (same results as ot_gemm_i8.cu)
      const bool row_slow = !(rmax < 3.0e38f) || !(fabsf(sx) < 3.0e38f) || cols_bad;
      // ---- pass 2: quantize from the registers into the staging tile; the registers of a finished chunk receive the next tile's
      // accumulator.  The quotient is EXACT and branch-free: q0 = y*r (r = RN(1/s)), two FMA residual steps q <- q + r*(y - q*s) give
      // q2 = RN(y/s) bit for bit (Markstein; tools/check_div_exact.c: 8e9 random / adversarial cases against the IEEE division), and
      // t = q2 + 1.5*2^23 holds rint(q2), ties to even, in its low mantissa byte.  6 FFMA2 per column pair, no flag, no redo: the
      // earlier "one multiply + redo the elements near a rounding boundary" pass was 2 FFMA2 cheaper, but one warp in eight took the
      // redo and its quarter, and through the row-maximum exchange the whole cluster, waited for it (2 us stragglers).
      const bool has_next = i + 1 < n_my;
      const uint32_t nbuf = buf ^ 1u;
      if (has_next) {
        mbar_wait(smem_u32(&tfull_bar[nbuf]), ((i + 1) >> 1) & 1u);
        tc_fence_after();
      }
      const f2 rr2 = pack2(s_rcp, s_rcp), ns2 = pack2(-s, -s);
      const bool warp_slow = __any_sync(0xffffffffu, row_slow);       // non-finite inputs somewhere in these 32 rows: the true division
      wtrace(g, i, e, 0);
      if (i > 0) mbar_wait(smem_u32(sempty), (i - 1) & 1u);             // the previous tile's TMA stores have read the staging buffer
      wtrace(g, i, e, 6);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t tb[16];
        if (warp_slow) {
#pragma unroll
          for (int j = 0; j < 16; ++j) tb[j] = static_cast<uint32_t>(__float2int_rn(rintf(__fdiv_rn(__uint_as_float(r[c][j]), s))));
        } else {
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const f2 y2 = pack2(__uint_as_float(r[c][j]), __uint_as_float(r[c][j + 1]));
            const f2 q0 = fma2(y2, rr2, kNeg0);
            const f2 q1 = fma2(fma2(q0, ns2, y2), rr2, q0);
            const f2 q2 = fma2(fma2(q1, ns2, y2), rr2, q1);
            const f2 t2 = fma2(q2, kOne, kMagic);
            tb[j] = static_cast<uint32_t>(t2 & 0xffffffffull);
            tb[j + 1] = static_cast<uint32_t>(t2 >> 32);
          }
        }
        if (has_next) tmem_ld_32x16(tlane + nbuf * kBN + 16 * c, r[c]);
        uint32_t packed[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          packed[j] = __byte_perm(__byte_perm(tb[4 * j], tb[4 * j + 1], 0x0040), __byte_perm(tb[4 * j + 2], tb[4 * j + 3], 0x0040), 0x5410);
        *reinterpret_cast<uint4*>(srow + (((ch0 + c) ^ sw) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      }
      if (has_next) {
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[nbuf]));      // the accumulator slice is in registers: back to the MMA warp
      }
      fence_proxy_async_smem();                                       // generic-proxy writes of the staging tile -> visible to the TMA store
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(sfull));
      wtrace(g, i, e, 5);
    }
    if (g.trace != nullptr && e == 0 && lane == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      g.trace[16 * 16 * 8 + 2 * blockIdx.x + 1] = t;
    }
  }

  tc_fence_before();
  if (cn > 1) cluster_sync_all();
  else __syncthreads();
  if (warp_idx == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace wres

// Launch when the problem qualifies; returns 1 if it does not (the caller falls back to the streaming / tile kernels), 0 on success.
int launch_gemm_wres(const int8_t* A, int64_t lda, const int8_t* W, int64_t ldw, int M, int N, int K, const float* row_scale, const float* col_scale,
                     const float* bias, int relu, void* out, int64_t ldo, float* out_scale, int quant_group, cudaStream_t stream, int w4) {
  using namespace wres;
  const char* env = getenv("OT_GEMM_WRES");      // "0": A/B runs against the streaming kernel (tests, tools/bench_gemm.py)
  const int enabled = env ? atoi(env) : 1;
  static const int min_m = getenv("OT_GEMM_STREAM_MIN_M") ? atoi(getenv("OT_GEMM_STREAM_MIN_M")) : 2048;
  if (!enabled || M < min_m || N % kBN != 0 || K % kBK != 0 || K < kBK || K > kMaxKb * kBK) return 1;
  if (quant_group % kBN != 0 || quant_group / kBN > kMaxCl || N % quant_group != 0) return 1;
  if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0 || ldo % 16 != 0) return 1;
  const int cn = quant_group / kBN;
  if ((cn & (cn - 1)) != 0) return 1;
  const int groups = N / quant_group;
  CUtensorMap ta, tw, to;
  // W4: W is the packed matrix [N][K / 2] (row pitch ldw bytes); boxes of 256 rows x 64 bytes, no swizzle (the unpacker re-lays them out)
  int rc = w4 ? get_tensor_map(&tw, W, N, K / 2, ldw, kBN, kBK / 2, false) : get_tensor_map(&tw, W, N, K, ldw, kBN, kBK, true);
  if (rc) return rc;
  rc = get_tensor_map(&ta, A, M, K, lda, kBM, kBK, true);
  if (rc) return rc;
  rc = get_tensor_map(&to, out, M, N, ldo, kBM, 128, true);
  if (rc) return rc;
  Args g = {};
  g.M = M; g.N = N; g.K = K;
  g.row_scale = row_scale; g.col_scale = col_scale; g.bias = bias; g.relu = relu; g.out_scale = out_scale; g.cluster_n = cn; g.groups = groups;
  g.neg0 = 0x8000000080000000ull; g.one = 0x3F8000003F800000ull; g.magic = 0x4B4000004B400000ull;
  if (const char* tr = getenv("OT_GEMM_WRES_TRACE")) g.trace = reinterpret_cast<unsigned long long*>(strtoull(tr, nullptr, 16));
  auto kernel = w4 ? (relu ? gemm_wres_kernel<true, true> : gemm_wres_kernel<false, true>) : (relu ? gemm_wres_kernel<true> : gemm_wres_kernel<false>);
  const int ki = (relu ? 1 : 0) + (w4 ? 2 : 0);
  static DeviceOnce attr_set[4];
  static int max_clusters_k[4][kMaxCl + 1] = {};
  int* max_clusters = max_clusters_k[ki];
  if (attr_set[ki].need()) OT_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int n_cl;
  if (cn == 1) {
    n_cl = sms;
  } else {
    if (max_clusters[cn] == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(static_cast<unsigned>(sms / cn * cn));
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = kSmem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cn; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = sms / cn / 2 > 0 ? sms / cn / 2 : 1;
      }
      max_clusters[cn] = n;
    }
    n_cl = max_clusters[cn];
  }
  const int m_tiles = (M + kBM - 1) / kBM;
  if (n_cl < groups) return 1;
  n_cl = n_cl / groups * groups;                      // the same number of clusters for every quant group
  if (n_cl > m_tiles * groups) n_cl = m_tiles * groups;
  OT_CHECK_CUDA(launch_kernel(kernel, dim3(static_cast<unsigned>(n_cl * cn)), dim3(kThreads), kSmem, stream, cn, ta, tw, to, g));
  count_launch();
  return OT_OK;
}

}  // namespace ot
