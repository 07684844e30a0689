// Shared device code of the quantized attention kernels: argument block, warp reductions and the decode (Tq = 1) body,
// which both attention_decode_kernel (ot_attention.cu) and the persistent decoder kernel (ot_decoder.cu) execute, so that
// the two paths are the same arithmetic instruction for instruction (attention.py:23-36 as exported).
#pragma once
#include "ot_common.h"

namespace ot {

constexpr int kHeads = 8;
constexpr int kDk = 64;
constexpr int kDm = kHeads * kDk;   // 512
constexpr int kKPitch = kDm + 16;   // padded K row pitch: conflict-free 128-bit reads across keys
constexpr int kMaxTk = 192;
constexpr int kKeysPerLane = kMaxTk / 32;
constexpr int kQG = 8;              // queries whose P.V products share one pass over the V rows (register tile 8 x 2 per lane)
constexpr int kDecKeysPerLane = 3;  // decode specialisation: Tk <= 96

struct AttnArgs {
  const int8_t* q; int64_t ldq; const float* sq; int64_t sq_stride;
  int8_t* k; int8_t* v; int64_t ldk; float* sk; float* sv; int64_t skv_stride;
  const int8_t* k_new; const int8_t* v_new; int64_t ld_new; const float* sk_new; const float* sv_new; int64_t snew_stride;
  int B, Tq, Tk, Tk_cap, mask_kind;
  const uint8_t* key_mask; int64_t mask_stride;
  int q_pos0;
  const int32_t* step_dev;
  float* ctx; int64_t ld_ctx;
  int8_t* ctx_q; float* ctx_s;
  uint8_t* probs_q;
  OtFault fault;
  const OtFault* mf_faults;   // batched trials: mf_unit[b] = index of sentence b's fault or -1; indices relative to the sentence
  const int32_t* mf_unit;
};

enum { OPERAND_Q = 0, OPERAND_K = 1, OPERAND_P = 2, OPERAND_V = 3, OPERAND_SCORES = 4, OPERAND_CTX = 5 };

__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The query row of one sentence and (self-attention) this step's new K/V row, already resolved to row pointers: they may
// live in global memory (the stand-alone kernel) or in shared memory (the persistent decoder, which produces them itself).
struct AttnDecRow {
  const int8_t* q;        // [512]
  float sq;
  const int8_t* k_new;    // [512] or nullptr (cross-attention)
  const int8_t* v_new;
  float sk_new, sv_new;
};

// float(int8) of byte B of w, where w holds the int8 bytes XOR 0x80 (i.e. v + 128 in [0, 255]): the float with bits 0x4B000000 | u
// is 2^23 + u exactly, and subtracting 2^23 + 128 is exact -- the value I2F.S8 gives, without the conversion pipe.
template <int B>
__device__ __forceinline__ float s8_as_float(uint32_t w) {
  return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 + B)), 8388736.0f);
}

typedef int8_t (*AttnDecVh)[32 * kDecKeysPerLane][kDk];   // [kHeads][96][64] shared-memory V slices, 16-byte aligned

// Decode attention of sentence b (Tq = 1, Tk <= 96, no fault): 256 threads, warp h = head h; K/V rows are read straight from
// the (L2-resident) cache.  Writes a.ctx (fp32, optional) and a.ctx_q / a.ctx_s (RowQuant for the O-projection) of row b and,
// when r.k_new is given, appends the new K/V row + scales at cache position Tk-1.  Ends with a block barrier.
__device__ __forceinline__ void attention_decode_body(const AttnArgs& a, const AttnDecRow& r, int b, int Tk, int q_pos0, AttnDecVh Vh) {
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int new0 = (r.k_new != nullptr) ? Tk - 1 : Tk;

  // append this step's K/V head slice (and, by head 0, the scales) to the cache
  if (r.k_new != nullptr) {
    const int64_t dst = (static_cast<int64_t>(b) * a.Tk_cap + new0) * a.ldk + h * kDk;
    if (lane < 4) *reinterpret_cast<uint4*>(a.k + dst + lane * 16) = *reinterpret_cast<const uint4*>(r.k_new + h * kDk + lane * 16);
    else if (lane < 8) *reinterpret_cast<uint4*>(a.v + dst + (lane - 4) * 16) = *reinterpret_cast<const uint4*>(r.v_new + h * kDk + (lane - 4) * 16);
    if (h == 0 && lane == 8) {
      const int64_t sdst = (static_cast<int64_t>(b) * a.Tk_cap + new0) * a.skv_stride;
      a.sk[sdst] = r.sk_new;
      a.sv[sdst] = r.sv_new;
    }
  }

  uint32_t qw[16];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(r.q + h * kDk);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint4 t = qp[w];
      qw[4 * w] = t.x; qw[4 * w + 1] = t.y; qw[4 * w + 2] = t.z; qw[4 * w + 3] = t.w;
    }
  }
  const float sqi = r.sq;
  // this head's V slice [Tk][64]: 4 lanes x 16 B per row, 8 rows per instruction; issued now so that their L2 latency
  // overlaps the K loads, the scores and the softmax (stored to shared memory just before the P.V loop)
  uint4 vbuf[4 * kDecKeysPerLane];
#pragma unroll
  for (int it = 0; it < 4 * kDecKeysPerLane; ++it) {
    const int jr = it * 8 + (lane >> 2);
    const int j = min(jr, Tk - 1);
    const int8_t* vp = (j >= new0) ? r.v_new + h * kDk : a.v + (static_cast<int64_t>(b) * a.Tk_cap + j) * a.ldk + h * kDk;
    const uint4 t = *reinterpret_cast<const uint4*>(vp + (lane & 3) * 16);
    vbuf[it] = (jr < Tk) ? t : make_uint4(0, 0, 0, 0);
  }

  // K rows, scales and mask of this lane's keys: branch-free (indices clamped, results masked) so that all loads of all
  // three key rounds are in flight together instead of one L2 round trip per round
  float sc[kDecKeysPerLane], svl[kDecKeysPerLane];
  uint4 kreg[kDecKeysPerLane][4];
  float skl[kDecKeysPerLane];
  uint8_t keepl[kDecKeysPerLane];
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk) {
    const int j = min(kk * 32 + lane, Tk - 1);
    const bool fresh = j >= new0;
    const int8_t* kp = fresh ? r.k_new + h * kDk : a.k + (static_cast<int64_t>(b) * a.Tk_cap + j) * a.ldk + h * kDk;
#pragma unroll
    for (int w = 0; w < 4; ++w) kreg[kk][w] = *reinterpret_cast<const uint4*>(kp + w * 16);
    const int64_t so = (static_cast<int64_t>(b) * a.Tk_cap + (fresh ? 0 : j)) * a.skv_stride;   // clamped: never dereferenced past the cache
    const float skc = a.sk[so], svc = a.sv[so];
    skl[kk] = fresh ? r.sk_new : skc;
    svl[kk] = fresh ? r.sv_new : svc;
    keepl[kk] = (a.mask_kind == 1) ? a.key_mask[static_cast<int64_t>(b) * a.mask_stride + j] : 1;
  }
  float mx = -INFINITY;
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk) {
    const int j = kk * 32 + lane;
    int dot = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint4 t = kreg[kk][w];
      dot = __dp4a(static_cast<int>(qw[4 * w]), static_cast<int>(t.x), dot);
      dot = __dp4a(static_cast<int>(qw[4 * w + 1]), static_cast<int>(t.y), dot);
      dot = __dp4a(static_cast<int>(qw[4 * w + 2]), static_cast<int>(t.z), dot);
      dot = __dp4a(static_cast<int>(qw[4 * w + 3]), static_cast<int>(t.w), dot);
    }
    const float s = __fdiv_rn(__fmul_rn(__fmul_rn(__int2float_rn(dot), sqi), skl[kk]), 8.0f);
    const bool visible = keepl[kk] != 0 && (a.mask_kind != 2 || j <= q_pos0);
    const bool live = j < Tk;
    sc[kk] = live ? (visible ? s : -1e9f) : -INFINITY;
    svl[kk] = live ? svl[kk] : 0.f;
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
  float pq[kDecKeysPerLane];   // quantized probability already divided by 127 (Div(127) of attention.py:35), one division per key
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk)
    pq[kk] = (kk * 32 + lane < Tk) ? __fdiv_rn(rintf(__fmul_rn(__fdiv_rn(sc[kk], sum), 127.0f)), 127.0f) : 0.f;

  // stage this head's V slice [Tk][64] into shared memory: 4 lanes x 16 B per row, 8 rows per instruction, all independent
#pragma unroll
  for (int it = 0; it < 4 * kDecKeysPerLane; ++it) {
    const int j = it * 8 + (lane >> 2);
    *reinterpret_cast<uint4*>(&Vh[h][j][(lane & 3) * 16]) = vbuf[it];   // rows >= Tk are zero-filled
  }
  __syncwarp();
  // context: lane owns features 2*lane, 2*lane+1; keys in order j = 0..Tk-1 (same order as the generic kernel)
  float acc0 = 0.f, acc1 = 0.f;
  const int d0 = 2 * lane;
  // Branch-free and unrolled: a key with p = 0 (masked, or beyond Tk where sv = 0) contributes exactly +0, so skipping it
  // (as the generic kernel does) and adding it give the same sum; 8 independent shuffles / shared loads are in flight.
#pragma unroll
  for (int kk = 0; kk < kDecKeysPerLane; ++kk) {
    if (kk * 32 >= Tk) break;
#pragma unroll 8
    for (int jj = 0; jj < 32; ++jj) {
      const float ph = __shfl_sync(0xffffffffu, pq[kk], jj);
      const float svj = __shfl_sync(0xffffffffu, svl[kk], jj);
      const uint32_t vv = static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&Vh[h][kk * 32 + jj][d0])) ^ 0x8080u;
      acc0 = fmaf(ph, __fmul_rn(s8_as_float<0>(vv), svj), acc0);
      acc1 = fmaf(ph, __fmul_rn(s8_as_float<1>(vv), svj), acc1);
    }
  }
  __syncwarp();
  *reinterpret_cast<float2*>(reinterpret_cast<float*>(&Vh[h][0][0]) + d0) = make_float2(acc0, acc1);
  __syncthreads();
  if (h == 0) {
    const int64_t row = b;
    float4 v[4];
    float amax = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      // feature f = (t*32+lane)*4 .. +3 lives in head f/64 at offset f%64
      const int f0 = (t * 32 + lane) * 4;
      v[t] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(&Vh[f0 >> 6][0][0]) + (f0 & 63));
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[t].x), fabsf(v[t].y)), fmaxf(fabsf(v[t].z), fabsf(v[t].w))));
    }
    if (a.ctx) {
#pragma unroll
      for (int t = 0; t < 4; ++t) *reinterpret_cast<float4*>(a.ctx + row * a.ld_ctx + (t * 32 + lane) * 4) = v[t];
    }
    if (a.ctx_q) {
      const float s = __fdiv_rn(fmaxf(warp_max_f(amax), 1e-5f), 127.0f);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int q0i = __float2int_rn(rintf(__fdiv_rn(v[t].x, s))), q1i = __float2int_rn(rintf(__fdiv_rn(v[t].y, s)));
        const int q2i = __float2int_rn(rintf(__fdiv_rn(v[t].z, s))), q3i = __float2int_rn(rintf(__fdiv_rn(v[t].w, s)));
        const uint32_t w = (static_cast<uint32_t>(q0i) & 0xFFu) | ((static_cast<uint32_t>(q1i) & 0xFFu) << 8) |
                           ((static_cast<uint32_t>(q2i) & 0xFFu) << 16) | ((static_cast<uint32_t>(q3i) & 0xFFu) << 24);
        *reinterpret_cast<uint32_t*>(a.ctx_q + row * kDm + (t * 32 + lane) * 4) = w;
      }
      if (lane == 0) a.ctx_s[row] = s;
    }
  }
  __syncthreads();
}

}  // namespace ot
