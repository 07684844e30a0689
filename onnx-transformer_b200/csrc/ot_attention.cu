// ot_attention_q8: fused quantized multi-head attention (attention.py:23-36 as exported).
//
// One CTA per (sentence b, tile of QT queries); warp h owns head h.  The sentence's int8 K and V rows (all 8
// heads) and their per-token scales are staged once in shared memory; scores are exact int32 dot products
// (dp4a) scaled in the canonical fp32 order  S = fl(fl(float(dot)*sq[i])*sk[j]) / 8 ; masked keys get -1e9;
// softmax uses warp-shuffle row reductions; P is quantized to rint(127 p); the context is accumulated in
// fp32 as  sum_j (pq_j/127) * (sv_j * vq[j,d])  (the V scale sits on the contraction axis, so this product is
// not an int8 GEMM -- SURVEY.md 0.6).  Because the CTA ends up owning complete 512-feature rows it also
// performs the per-token RowQuant that feeds the O-projection, and, for the decoder, appends the step's new
// K/V rows to the persistent KV cache.
#include "ot_attention_decode.cuh"
#include "ot_common.h"

namespace ot {

__device__ __forceinline__ float patch_f32(const OtFault& f, float v) {
  uint32_t bits = __float_as_uint(v);
  if (f.mode == OT_FAULT_RANDOM_BITFLIP) bits ^= (1u << f.bit);
  else bits = f.value_bits;
  const float r = __uint_as_float(bits);
  return (r != r) ? 0.0f : r;
}

template <int QT>
__global__ void __launch_bounds__(256, QT <= 8 ? 2 : 1) attention_q8_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const unsigned int tl = tl_begin(2);
  pdl_wait();      // upstream results are complete and visible from here on
  pdl_trigger();   // now let exactly one successor start its launch + prologue (look-ahead depth 1)
  tl_mark(tl, 2);
  const int step = a.step_dev ? *a.step_dev : 0;
  const int Tk = a.step_dev ? step + a.Tq : a.Tk;
  const int q_pos0 = a.step_dev ? step : a.q_pos0;
  if (a.step_dev && (step < 0 || step + a.Tq > a.Tk_cap)) return;      // a device-side step past the cache capacity: never append out of bounds
  const int Tk_pad = (Tk + 31) & ~31;

  int8_t* Ks = reinterpret_cast<int8_t*>(smem);                        // [Tk][kKPitch]
  int8_t* Vs = Ks + static_cast<size_t>(Tk_pad) * kKPitch;            // [Tk][512]
  int8_t* Qs = Vs + static_cast<size_t>(Tk_pad) * kDm;                // [QT][512]
  float* Cs = reinterpret_cast<float*>(Qs + QT * kDm);                // [QT][512] fp32 context tile
  float* Pw = Cs + QT * kDm;                                          // [8 warps][Tk_pad][kQG] quantized probabilities
  float* sks = Pw + kHeads * Tk_pad * kQG;                                  // [Tk_pad]
  float* svs = sks + Tk_pad;                                          // [Tk_pad]
  uint8_t* keep = reinterpret_cast<uint8_t*>(svs + Tk_pad);           // [Tk_pad] key-padding mask

  const int b = blockIdx.y;
  const int q0 = blockIdx.x * QT;
  const int nq = min(QT, a.Tq - q0);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int new0 = (a.k_new != nullptr) ? Tk - a.Tq : Tk;  // keys >= new0 come from this step's projections
  const bool writer = (blockIdx.x == 0);

  // ---- stage K, V (16-byte chunks), scales, mask, Q tile
  for (int idx = tid; idx < Tk * (kDm / 16); idx += blockDim.x) {
    const int j = idx >> 5, c = idx & 31;
    uint4 kk, vv;
    if (j >= new0) {
      const int64_t src = (static_cast<int64_t>(b) * a.Tq + (j - new0)) * a.ld_new + c * 16;
      kk = *reinterpret_cast<const uint4*>(a.k_new + src);
      vv = *reinterpret_cast<const uint4*>(a.v_new + src);
      if (writer) {
        const int64_t dst = (static_cast<int64_t>(b) * a.Tk_cap + j) * a.ldk + c * 16;
        *reinterpret_cast<uint4*>(a.k + dst) = kk;
        *reinterpret_cast<uint4*>(a.v + dst) = vv;
      }
    } else {
      const int64_t src = (static_cast<int64_t>(b) * a.Tk_cap + j) * a.ldk + c * 16;
      kk = *reinterpret_cast<const uint4*>(a.k + src);
      vv = *reinterpret_cast<const uint4*>(a.v + src);
    }
    *reinterpret_cast<uint4*>(Ks + j * kKPitch + c * 16) = kk;
    *reinterpret_cast<uint4*>(Vs + j * kDm + c * 16) = vv;
  }
  for (int j = tid; j < Tk_pad; j += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    uint8_t kp = 0;
    if (j < Tk) {
      if (j >= new0) {
        const int64_t src = (static_cast<int64_t>(b) * a.Tq + (j - new0)) * a.snew_stride;
        s1 = a.sk_new[src];
        s2 = a.sv_new[src];
        if (writer) {
          const int64_t dst = (static_cast<int64_t>(b) * a.Tk_cap + j) * a.skv_stride;
          a.sk[dst] = s1;
          a.sv[dst] = s2;
        }
      } else {
        const int64_t src = (static_cast<int64_t>(b) * a.Tk_cap + j) * a.skv_stride;
        s1 = a.sk[src];
        s2 = a.sv[src];
      }
      kp = (a.mask_kind == 1) ? a.key_mask[static_cast<int64_t>(b) * a.mask_stride + j] : 1;
    }
    sks[j] = s1;
    svs[j] = s2;
    keep[j] = kp;
  }
  for (int idx = tid; idx < nq * (kDm / 16); idx += blockDim.x) {
    const int i = idx >> 5, c = idx & 31;
    *reinterpret_cast<uint4*>(Qs + i * kDm + c * 16) =
        *reinterpret_cast<const uint4*>(a.q + (static_cast<int64_t>(b) * a.Tq + q0 + i) * a.ldq + c * 16);
  }
  __syncthreads();

  // ---- fault context (App. D); operand selector in fault.reserved
  OtFault f = a.fault;
  if (a.mf_unit != nullptr) {
    const int fi_ = a.mf_unit[b];
    if (fi_ >= 0) f = a.mf_faults[fi_];
    else f.mode = OT_FAULT_NONE;
  }
  const bool has_fault = f.mode != OT_FAULT_NONE;
  int fb = -1, fh = -1, fi = -1, fj = -1, fd = -1, fw0 = 0, fw1 = 0;
  if (has_fault) {
    const int64_t idx = f.flat_index;
    const int wl = f.window_len;
    switch (f.reserved) {
      case OPERAND_Q: {  // Round tensor [B,Tq,512]
        fb = static_cast<int>(idx / (static_cast<int64_t>(a.Tq) * kDm));
        fi = static_cast<int>((idx / kDm) % a.Tq);
        fh = static_cast<int>(idx % kDm) / kDk; fd = static_cast<int>(idx % kDm) % kDk;
        fw0 = wl > 0 ? f.window_start : 0; fw1 = wl > 0 ? min(Tk, f.window_start + wl) : Tk;       // key window
      } break;
      case OPERAND_K: case OPERAND_V: {  // Round tensor [B,Tk,512]
        fb = static_cast<int>(idx / (static_cast<int64_t>(Tk) * kDm));
        fj = static_cast<int>((idx / kDm) % Tk);
        fh = static_cast<int>(idx % kDm) / kDk; fd = static_cast<int>(idx % kDm) % kDk;
        fw0 = wl > 0 ? f.window_start : 0; fw1 = wl > 0 ? min(a.Tq, f.window_start + wl) : a.Tq;   // query window
      } break;
      case OPERAND_P: case OPERAND_SCORES: {  // [B,8,Tq,Tk]
        fj = static_cast<int>(idx % Tk);
        fi = static_cast<int>((idx / Tk) % a.Tq);
        fh = static_cast<int>((idx / (static_cast<int64_t>(Tk) * a.Tq)) % kHeads);
        fb = static_cast<int>(idx / (static_cast<int64_t>(Tk) * a.Tq * kHeads));
        fw0 = wl > 0 ? f.window_start : 0; fw1 = wl > 0 ? min(kDk, f.window_start + wl) : kDk;     // feature window
      } break;
      default: {  // OPERAND_CTX: [B,8,Tq,64]
        fd = static_cast<int>(idx % kDk);
        fi = static_cast<int>((idx / kDk) % a.Tq);
        fh = static_cast<int>((idx / (static_cast<int64_t>(kDk) * a.Tq)) % kHeads);
        fb = static_cast<int>(idx / (static_cast<int64_t>(kDk) * a.Tq * kHeads));
      } break;
    }
  }
  if (a.mf_unit != nullptr) fb = b;   // batched faults address the sentence's own tensors
  const bool fault_here = has_fault && fb == b && fh == warp;

  // ---- per head (warp) and query: scores -> softmax -> quantized P -> context
  const int h = warp;
  float* P = Pw + h * Tk_pad * kQG;
  if (!has_fault) {
    // ---- fast path: groups of kQG queries.  Phase 1 computes each query's quantized probabilities into a transposed
    // [key][kQG] tile; phase 2 walks the V rows ONCE for the whole group (v is converted and scaled once per key, the
    // kQG probabilities arrive as two 128-bit broadcast loads), same summation order over keys as the per-query path.
    const int d0 = 2 * lane;
    for (int g0 = 0; g0 < nq; g0 += kQG) {
      const int ng = min(kQG, nq - g0);
#pragma unroll 1
      for (int qq = 0; qq < kQG; ++qq) {
        if (qq >= ng) {
#pragma unroll
          for (int kk = 0; kk < kKeysPerLane; ++kk)
            if (kk * 32 < Tk_pad) P[(kk * 32 + lane) * kQG + qq] = 0.f;
          continue;
        }
        const int iq = g0 + qq, i = q0 + iq;
        const float sqi = a.sq[(static_cast<int64_t>(b) * a.Tq + i) * a.sq_stride];
        uint32_t qw[16];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const uint4 t = *reinterpret_cast<const uint4*>(Qs + iq * kDm + h * kDk + w * 16);
          qw[4 * w] = t.x; qw[4 * w + 1] = t.y; qw[4 * w + 2] = t.z; qw[4 * w + 3] = t.w;
        }
        float sc[kKeysPerLane];
        float mx = -INFINITY;
#pragma unroll
        for (int kk = 0; kk < kKeysPerLane; ++kk) {
          const int j = kk * 32 + lane;
          sc[kk] = -INFINITY;
          if (kk * 32 < Tk && j < Tk) {
            int dot = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const uint4 t = *reinterpret_cast<const uint4*>(Ks + j * kKPitch + h * kDk + w * 16);
              dot = __dp4a(static_cast<int>(qw[4 * w]), static_cast<int>(t.x), dot);
              dot = __dp4a(static_cast<int>(qw[4 * w + 1]), static_cast<int>(t.y), dot);
              dot = __dp4a(static_cast<int>(qw[4 * w + 2]), static_cast<int>(t.z), dot);
              dot = __dp4a(static_cast<int>(qw[4 * w + 3]), static_cast<int>(t.w), dot);
            }
            const float sv_ = __fdiv_rn(__fmul_rn(__fmul_rn(__int2float_rn(dot), sqi), sks[j]), 8.0f);
            const bool visible = keep[j] && (a.mask_kind != 2 || j <= q_pos0 + i);
            sc[kk] = visible ? sv_ : -1e9f;
            mx = fmaxf(mx, sc[kk]);
          }
        }
        mx = warp_max_f(mx);
        float sum = 0.f;
#pragma unroll
        for (int kk = 0; kk < kKeysPerLane; ++kk) {
          const int j = kk * 32 + lane;
          if (kk * 32 < Tk && j < Tk) {
            sc[kk] = expf(__fsub_rn(sc[kk], mx));
            sum += sc[kk];
          }
        }
        sum = warp_sum_f(sum);
#pragma unroll
        for (int kk = 0; kk < kKeysPerLane; ++kk) {
          const int j = kk * 32 + lane;
          if (kk * 32 < Tk_pad) {
            float ph = 0.f;
            if (j < Tk) {
              const float pq = rintf(__fmul_rn(__fdiv_rn(sc[kk], sum), 127.0f));   // Round(Mul(p,127))
              if (a.probs_q) a.probs_q[((static_cast<int64_t>(b) * kHeads + h) * a.Tq + i) * Tk + j] = static_cast<uint8_t>(pq);
              ph = __fdiv_rn(pq, 127.0f);                                          // Div(127)
            }
            P[j * kQG + qq] = ph;
          }
        }
      }
      __syncwarp();
      float acc[kQG][2];
#pragma unroll
      for (int qq = 0; qq < kQG; ++qq) acc[qq][0] = acc[qq][1] = 0.f;
#pragma unroll 2
      for (int j = 0; j < Tk; ++j) {
        const char2 vv = *reinterpret_cast<const char2*>(Vs + j * kDm + h * kDk + d0);
        const float svj = svs[j];
        const float v0 = __fmul_rn(__int2float_rn(vv.x), svj), v1 = __fmul_rn(__int2float_rn(vv.y), svj);
        const float4 pa = *reinterpret_cast<const float4*>(P + j * kQG);
        const float4 pb = *reinterpret_cast<const float4*>(P + j * kQG + 4);
        acc[0][0] = fmaf(pa.x, v0, acc[0][0]); acc[0][1] = fmaf(pa.x, v1, acc[0][1]);
        acc[1][0] = fmaf(pa.y, v0, acc[1][0]); acc[1][1] = fmaf(pa.y, v1, acc[1][1]);
        acc[2][0] = fmaf(pa.z, v0, acc[2][0]); acc[2][1] = fmaf(pa.z, v1, acc[2][1]);
        acc[3][0] = fmaf(pa.w, v0, acc[3][0]); acc[3][1] = fmaf(pa.w, v1, acc[3][1]);
        acc[4][0] = fmaf(pb.x, v0, acc[4][0]); acc[4][1] = fmaf(pb.x, v1, acc[4][1]);
        acc[5][0] = fmaf(pb.y, v0, acc[5][0]); acc[5][1] = fmaf(pb.y, v1, acc[5][1]);
        acc[6][0] = fmaf(pb.z, v0, acc[6][0]); acc[6][1] = fmaf(pb.z, v1, acc[6][1]);
        acc[7][0] = fmaf(pb.w, v0, acc[7][0]); acc[7][1] = fmaf(pb.w, v1, acc[7][1]);
      }
#pragma unroll
      for (int qq = 0; qq < kQG; ++qq)
        if (qq < ng) *reinterpret_cast<float2*>(Cs + (g0 + qq) * kDm + h * kDk + d0) = make_float2(acc[qq][0], acc[qq][1]);
      __syncwarp();
    }
  } else
  for (int iq = 0; iq < nq; ++iq) {
    const int i = q0 + iq;
    const float sqi = a.sq[(static_cast<int64_t>(b) * a.Tq + i) * a.sq_stride];
    uint32_t qw[16];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint4 t = *reinterpret_cast<const uint4*>(Qs + iq * kDm + h * kDk + w * 16);
      qw[4 * w] = t.x; qw[4 * w + 1] = t.y; qw[4 * w + 2] = t.z; qw[4 * w + 3] = t.w;
    }
    float sc[kKeysPerLane];
    float mx = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < kKeysPerLane; ++kk) {
      const int j = kk * 32 + lane;
      sc[kk] = -INFINITY;
      if (kk * 32 < Tk && j < Tk) {
        int dot = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const uint4 t = *reinterpret_cast<const uint4*>(Ks + j * kKPitch + h * kDk + w * 16);
          dot = __dp4a(static_cast<int>(qw[4 * w]), static_cast<int>(t.x), dot);
          dot = __dp4a(static_cast<int>(qw[4 * w + 1]), static_cast<int>(t.y), dot);
          dot = __dp4a(static_cast<int>(qw[4 * w + 2]), static_cast<int>(t.z), dot);
          dot = __dp4a(static_cast<int>(qw[4 * w + 3]), static_cast<int>(t.w), dot);
        }
        if (fault_here) {
          if (f.reserved == OPERAND_Q && f.mode == OT_FAULT_INPUT && i == fi && j >= fw0 && j < fw1) {
            const int qv = Qs[iq * kDm + h * kDk + fd];
            dot += (flip_int8_bit(qv, f.bit) - qv) * Ks[j * kKPitch + h * kDk + fd];
          } else if (f.reserved == OPERAND_K && f.mode == OT_FAULT_WEIGHT && j == fj && i >= fw0 && i < fw1) {
            const int kv = Ks[j * kKPitch + h * kDk + fd];
            dot += Qs[iq * kDm + h * kDk + fd] * (flip_int8_bit(kv, f.bit) - kv);
          }
        }
        float s = __fmul_rn(__fmul_rn(__int2float_rn(dot), sqi), sks[j]);   // MatMul_k_out0
        if (fault_here && f.reserved == OPERAND_SCORES && i == fi && j == fj) s = patch_f32(f, s);
        s = __fdiv_rn(s, 8.0f);                                              // / sqrt(d_k)
        const bool visible = keep[j] && (a.mask_kind != 2 || j <= q_pos0 + i);
        sc[kk] = visible ? s : -1e9f;                                        // masked_fill(mask == 0, -1e9)
        mx = fmaxf(mx, sc[kk]);
      }
    }
    mx = warp_max_f(mx);
    float sum = 0.f;
#pragma unroll
    for (int kk = 0; kk < kKeysPerLane; ++kk) {
      const int j = kk * 32 + lane;
      if (kk * 32 < Tk && j < Tk) {
        sc[kk] = expf(__fsub_rn(sc[kk], mx));
        sum += sc[kk];
      }
    }
    sum = warp_sum_f(sum);
#pragma unroll
    for (int kk = 0; kk < kKeysPerLane; ++kk) {
      const int j = kk * 32 + lane;
      if (kk * 32 < Tk_pad) {
        float pq = 0.f;
        if (j < Tk) {
          pq = rintf(__fmul_rn(__fdiv_rn(sc[kk], sum), 127.0f));             // Round(Mul(p,127))
          if (a.probs_q) a.probs_q[((static_cast<int64_t>(b) * kHeads + h) * a.Tq + i) * Tk + j] = static_cast<uint8_t>(pq);
        }
        P[j] = pq;                                                          // integer-valued, 0..127
      }
    }
    __syncwarp();

    // context: lane owns features 2*lane, 2*lane+1 of this head
    float acc0 = 0.f, acc1 = 0.f;
    const int d0 = 2 * lane;
    const bool p_fault = fault_here && f.reserved == OPERAND_P && f.mode == OT_FAULT_INPUT && i == fi;
    const bool v_fault = fault_here && f.reserved == OPERAND_V && f.mode == OT_FAULT_WEIGHT && i >= fw0 && i < fw1;
    for (int j = 0; j < Tk; ++j) {
      float pq = P[j];
      float pq0 = pq, pq1 = pq;
      if (p_fault && j == fj) {
        const float pf = static_cast<float>(flip_int8_bit(static_cast<int>(pq), f.bit));
        if (d0 >= fw0 && d0 < fw1) pq0 = pf;
        if (d0 + 1 >= fw0 && d0 + 1 < fw1) pq1 = pf;
      } else if (pq == 0.f) {
        continue;  // exact zero contribution (warp-uniform branch)
      }
      const char2 vv = *reinterpret_cast<const char2*>(Vs + j * kDm + h * kDk + d0);
      int v0 = vv.x, v1 = vv.y;
      if (v_fault && j == fj) {
        if (d0 == fd) v0 = flip_int8_bit(v0, f.bit);
        if (d0 + 1 == fd) v1 = flip_int8_bit(v1, f.bit);
      }
      const float svj = svs[j];
      acc0 = fmaf(__fdiv_rn(pq0, 127.0f), __fmul_rn(__int2float_rn(v0), svj), acc0);
      acc1 = fmaf(__fdiv_rn(pq1, 127.0f), __fmul_rn(__int2float_rn(v1), svj), acc1);
    }
    if (fault_here && f.reserved == OPERAND_CTX && i == fi) {
      if (d0 == fd) acc0 = patch_f32(f, acc0);
      if (d0 + 1 == fd) acc1 = patch_f32(f, acc1);
    }
    *reinterpret_cast<float2*>(Cs + iq * kDm + h * kDk + d0) = make_float2(acc0, acc1);
    __syncwarp();
  }
  __syncthreads();

  // ---- merge heads (Transpose + Reshape are free: Cs is already [query][512]); optional RowQuant (a7)
  for (int iq = warp; iq < nq; iq += 8) {
    const int64_t row = static_cast<int64_t>(b) * a.Tq + q0 + iq;
    float4 v[4];
    float amax = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      v[t] = *reinterpret_cast<const float4*>(Cs + iq * kDm + (t * 32 + lane) * 4);
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
}

// ------------------------------------------------------------------------------------------------
// Encoder-size specialisation (Tq >= 32, fault-free, fp32 context requested): one CTA per (sentence, HEAD).  The kernel above
// stages all 8 heads' K / V of a sentence (132 KB at Tk = 128: one CTA of 8 warps per SM) and is latency-bound at 0.25 IPC; a
// head's K / V slice is 18 KB, so three of these CTAs (24 warps) share an SM and a sentence's K / V bytes are read once instead of
// once per query tile.  Same arithmetic, op for op, as the fast path of attention_q8_kernel (groups of kQG queries; P.V walks the
// keys in order for the whole group).  The per-token RowQuant of the merged context needs all 8 heads of a row, i.e. 8 CTAs:
// ot_attention_q8 runs rowquant_kernel on the fp32 context afterwards (same quant_scale / quant_one instructions).
constexpr int kKhPitch = kDk + 16;     // 80-byte K rows: conflict-free 128-bit reads with lane = key

__global__ void __launch_bounds__(256) attention_heads_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  pdl_wait();
  pdl_trigger();
  const int Tk = a.Tk, Tq = a.Tq;
  const int Tk_pad = (Tk + 31) & ~31;
  const int h = blockIdx.x, b = blockIdx.y;
  int8_t* Kh = reinterpret_cast<int8_t*>(smem);                              // [Tk_pad][80]
  int8_t* Vh = Kh + static_cast<size_t>(Tk_pad) * kKhPitch;                  // [Tk_pad][64]
  int8_t* Qh = Vh + static_cast<size_t>(Tk_pad) * kDk;                       // [Tq][64]
  float* Pw = reinterpret_cast<float*>(Qh + ((static_cast<size_t>(Tq) * kDk + 15) & ~static_cast<size_t>(15)));   // [8 warps][Tk_pad][kQG]
  float* sks = Pw + static_cast<size_t>(8) * Tk_pad * kQG;                   // [Tk_pad]
  float* svs = sks + Tk_pad;
  uint8_t* keep = reinterpret_cast<uint8_t*>(svs + Tk_pad);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int idx = tid; idx < Tk * 4; idx += 256) {
    const int j = idx >> 2, c = idx & 3;
    const int64_t src = (static_cast<int64_t>(b) * a.Tk_cap + j) * a.ldk + h * kDk + c * 16;
    *reinterpret_cast<uint4*>(Kh + j * kKhPitch + c * 16) = *reinterpret_cast<const uint4*>(a.k + src);
    *reinterpret_cast<uint4*>(Vh + j * kDk + c * 16) = *reinterpret_cast<const uint4*>(a.v + src);
  }
  for (int idx = tid; idx < Tq * 4; idx += 256) {
    const int i = idx >> 2, c = idx & 3;
    *reinterpret_cast<uint4*>(Qh + i * kDk + c * 16) =
        *reinterpret_cast<const uint4*>(a.q + (static_cast<int64_t>(b) * Tq + i) * a.ldq + h * kDk + c * 16);
  }
  for (int j = tid; j < Tk_pad; j += 256) {
    float s1 = 0.f, s2 = 0.f;
    uint8_t kp = 0;
    if (j < Tk) {
      const int64_t src = (static_cast<int64_t>(b) * a.Tk_cap + j) * a.skv_stride;
      s1 = a.sk[src];
      s2 = a.sv[src];
      kp = (a.mask_kind == 1) ? a.key_mask[static_cast<int64_t>(b) * a.mask_stride + j] : 1;
    }
    sks[j] = s1; svs[j] = s2; keep[j] = kp;
  }
  __syncthreads();

  float* P = Pw + static_cast<size_t>(warp) * Tk_pad * kQG;
  const int d0 = 2 * lane;
  for (int g0 = warp * kQG; g0 < Tq; g0 += 8 * kQG) {
    const int ng = min(kQG, Tq - g0);
#pragma unroll 1
    for (int qq = 0; qq < kQG; ++qq) {
      if (qq >= ng) {
#pragma unroll
        for (int kk = 0; kk < kKeysPerLane; ++kk)
          if (kk * 32 < Tk_pad) P[(kk * 32 + lane) * kQG + qq] = 0.f;
        continue;
      }
      const int i = g0 + qq;
      const float sqi = a.sq[(static_cast<int64_t>(b) * Tq + i) * a.sq_stride];
      uint32_t qw[16];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint4 t = *reinterpret_cast<const uint4*>(Qh + i * kDk + w * 16);
        qw[4 * w] = t.x; qw[4 * w + 1] = t.y; qw[4 * w + 2] = t.z; qw[4 * w + 3] = t.w;
      }
      float sc[kKeysPerLane];
      float mx = -INFINITY;
#pragma unroll
      for (int kk = 0; kk < kKeysPerLane; ++kk) {
        const int j = kk * 32 + lane;
        sc[kk] = -INFINITY;
        if (kk * 32 < Tk && j < Tk) {
          int dot = 0;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const uint4 t = *reinterpret_cast<const uint4*>(Kh + j * kKhPitch + w * 16);
            dot = __dp4a(static_cast<int>(qw[4 * w]), static_cast<int>(t.x), dot);
            dot = __dp4a(static_cast<int>(qw[4 * w + 1]), static_cast<int>(t.y), dot);
            dot = __dp4a(static_cast<int>(qw[4 * w + 2]), static_cast<int>(t.z), dot);
            dot = __dp4a(static_cast<int>(qw[4 * w + 3]), static_cast<int>(t.w), dot);
          }
          const float sv_ = __fdiv_rn(__fmul_rn(__fmul_rn(__int2float_rn(dot), sqi), sks[j]), 8.0f);
          const bool visible = keep[j] && (a.mask_kind != 2 || j <= a.q_pos0 + i);
          sc[kk] = visible ? sv_ : -1e9f;
          mx = fmaxf(mx, sc[kk]);
        }
      }
      mx = warp_max_f(mx);
      float sum = 0.f;
#pragma unroll
      for (int kk = 0; kk < kKeysPerLane; ++kk) {
        const int j = kk * 32 + lane;
        if (kk * 32 < Tk && j < Tk) {
          sc[kk] = expf(__fsub_rn(sc[kk], mx));
          sum += sc[kk];
        }
      }
      sum = warp_sum_f(sum);
#pragma unroll
      for (int kk = 0; kk < kKeysPerLane; ++kk) {
        const int j = kk * 32 + lane;
        if (kk * 32 < Tk_pad) {
          float ph = 0.f;
          if (j < Tk) ph = __fdiv_rn(rintf(__fmul_rn(__fdiv_rn(sc[kk], sum), 127.0f)), 127.0f);   // Round(Mul(p,127)) then Div(127)
          P[j * kQG + qq] = ph;
        }
      }
    }
    __syncwarp();
    float acc[kQG][2];
#pragma unroll
    for (int qq = 0; qq < kQG; ++qq) acc[qq][0] = acc[qq][1] = 0.f;
#pragma unroll 2
    for (int j = 0; j < Tk; ++j) {
      const char2 vv = *reinterpret_cast<const char2*>(Vh + j * kDk + d0);
      const float svj = svs[j];
      const float v0 = __fmul_rn(__int2float_rn(vv.x), svj), v1 = __fmul_rn(__int2float_rn(vv.y), svj);
      const float4 pa = *reinterpret_cast<const float4*>(P + j * kQG);
      const float4 pb = *reinterpret_cast<const float4*>(P + j * kQG + 4);
      acc[0][0] = fmaf(pa.x, v0, acc[0][0]); acc[0][1] = fmaf(pa.x, v1, acc[0][1]);
      acc[1][0] = fmaf(pa.y, v0, acc[1][0]); acc[1][1] = fmaf(pa.y, v1, acc[1][1]);
      acc[2][0] = fmaf(pa.z, v0, acc[2][0]); acc[2][1] = fmaf(pa.z, v1, acc[2][1]);
      acc[3][0] = fmaf(pa.w, v0, acc[3][0]); acc[3][1] = fmaf(pa.w, v1, acc[3][1]);
      acc[4][0] = fmaf(pb.x, v0, acc[4][0]); acc[4][1] = fmaf(pb.x, v1, acc[4][1]);
      acc[5][0] = fmaf(pb.y, v0, acc[5][0]); acc[5][1] = fmaf(pb.y, v1, acc[5][1]);
      acc[6][0] = fmaf(pb.z, v0, acc[6][0]); acc[6][1] = fmaf(pb.z, v1, acc[6][1]);
      acc[7][0] = fmaf(pb.w, v0, acc[7][0]); acc[7][1] = fmaf(pb.w, v1, acc[7][1]);
    }
#pragma unroll
    for (int qq = 0; qq < kQG; ++qq)
      if (qq < ng)
        *reinterpret_cast<float2*>(a.ctx + (static_cast<int64_t>(b) * Tq + g0 + qq) * a.ld_ctx + h * kDk + d0) = make_float2(acc[qq][0], acc[qq][1]);
    __syncwarp();
  }
}

static size_t attn_heads_smem_bytes(int Tq, int Tk) {
  const size_t Tk_pad = (Tk + 31) & ~31;
  return Tk_pad * kKhPitch + Tk_pad * kDk + ((static_cast<size_t>(Tq) * kDk + 15) & ~static_cast<size_t>(15)) + 8 * Tk_pad * kQG * 4 + 2 * Tk_pad * 4 +
         Tk_pad + 16;
}

// ------------------------------------------------------------------------------------------------
// Decode specialisation (Tq = 1, Tk <= 96, no fault, no probability dump): one CTA per sentence, warp h = head h,
// K/V rows are read straight from the (L2-resident) cache -- no shared-memory staging, no block barrier before the
// head merge.  Same arithmetic, op for op, as attention_q8_kernel; the body lives in ot_attention_decode.cuh.
__global__ void __launch_bounds__(256) attention_decode_kernel(const AttnArgs a) {
  // per-head V slices (48 KB): low-latency P.V operands; a warp's finished context row (64 floats) overwrites the head of its own slice
  __shared__ __align__(16) int8_t Vh[kHeads][32 * kDecKeysPerLane][kDk];
  const unsigned int tl = tl_begin(3);
  pdl_wait();      // upstream results are complete and visible from here on
  pdl_trigger();   // now let exactly one successor start its launch + prologue (look-ahead depth 1)
  tl_mark(tl, 2);
  const int step = a.step_dev ? *a.step_dev : 0;
  const int Tk = a.step_dev ? step + 1 : a.Tk;
  const int q_pos0 = a.step_dev ? step : a.q_pos0;
  if (a.step_dev && (step < 0 || step + 1 > a.Tk_cap)) return;         // a device-side step past the cache capacity: never append out of bounds
  const int b = blockIdx.x;
  AttnDecRow r;
  r.q = a.q + static_cast<int64_t>(b) * a.ldq;
  r.sq = a.sq[static_cast<int64_t>(b) * a.sq_stride];
  r.k_new = r.v_new = nullptr;
  r.sk_new = r.sv_new = 0.f;
  if (a.k_new != nullptr) {
    r.k_new = a.k_new + static_cast<int64_t>(b) * a.ld_new;
    r.v_new = a.v_new + static_cast<int64_t>(b) * a.ld_new;
    r.sk_new = a.sk_new[static_cast<int64_t>(b) * a.snew_stride];
    r.sv_new = a.sv_new[static_cast<int64_t>(b) * a.snew_stride];
  }
  attention_decode_body(a, r, b, Tk, q_pos0, Vh);
  tl_mark(tl, 3);
}

template <int QT>
static size_t attn_smem_bytes(int Tk) {
  const int Tk_pad = (Tk + 31) & ~31;
  return static_cast<size_t>(Tk_pad) * kKPitch + static_cast<size_t>(Tk_pad) * kDm + QT * kDm + QT * kDm * 4 +
         static_cast<size_t>(kHeads) * Tk_pad * kQG * 4 + 2 * Tk_pad * 4 + Tk_pad;
}

template <int QT>
static int launch_attention(const AttnArgs& a, int tk_max, cudaStream_t stream) {
  const size_t smem = attn_smem_bytes<QT>(tk_max);
  OT_REQUIRE(smem <= 227 * 1024, "attention tile does not fit in shared memory");
  auto kernel = attention_q8_kernel<QT>;
  static DeviceOnce configured;
  if (configured.need(smem)) OT_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid((a.Tq + QT - 1) / QT, a.B, 1);
  OT_CHECK_CUDA(launch_kernel(kernel, grid, dim3(256), smem, stream, 1, a));
  count_launch();
  return OT_OK;
}

OT_DEFINE_TL_SETTER(tl_set_attention)

int launch_attention_tc(const AttnArgs& a, cudaStream_t stream, bool* fused_q);   // ot_attention_tc.cu

}  // namespace ot

using namespace ot;

extern "C" int ot_attention_q8_mf(const int8_t* q, int64_t ldq, const float* sq, int64_t sq_stride, int8_t* k, int8_t* v, int64_t ldk, float* sk,
                                  float* sv, int64_t skv_stride, const int8_t* k_new, const int8_t* v_new, int64_t ld_new, const float* sk_new,
                                  const float* sv_new, int64_t snew_stride, int B, int H, int Tq, int Tk, int Tk_cap, int mask_kind,
                                  const uint8_t* key_mask, int64_t mask_stride, int q_pos0, const int32_t* step_dev, float* ctx, int64_t ld_ctx,
                                  int8_t* ctx_q, float* ctx_s, uint8_t* probs_q, const OtFault* fault, const OtFault* faults_dev,
                                  const int32_t* unit_fault_dev, void* stream);

extern "C" int ot_attention_q8(const int8_t* q, int64_t ldq, const float* sq, int64_t sq_stride,
                               int8_t* k, int8_t* v, int64_t ldk, float* sk, float* sv, int64_t skv_stride,
                               const int8_t* k_new, const int8_t* v_new, int64_t ld_new, const float* sk_new,
                               const float* sv_new, int64_t snew_stride,
                               int B, int H, int Tq, int Tk, int Tk_cap, int mask_kind, const uint8_t* key_mask,
                               int64_t mask_stride, int q_pos0, const int32_t* step_dev,
                               float* ctx, int64_t ld_ctx, int8_t* ctx_q, float* ctx_s, uint8_t* probs_q,
                               const OtFault* fault, void* stream) {
  return ot_attention_q8_mf(q, ldq, sq, sq_stride, k, v, ldk, sk, sv, skv_stride, k_new, v_new, ld_new, sk_new, sv_new, snew_stride, B, H, Tq, Tk,
                            Tk_cap, mask_kind, key_mask, mask_stride, q_pos0, step_dev, ctx, ld_ctx, ctx_q, ctx_s, probs_q, fault, nullptr, nullptr,
                            stream);
}

extern "C" int ot_attention_q8_mf(const int8_t* q, int64_t ldq, const float* sq, int64_t sq_stride,
                                  int8_t* k, int8_t* v, int64_t ldk, float* sk, float* sv, int64_t skv_stride,
                                  const int8_t* k_new, const int8_t* v_new, int64_t ld_new, const float* sk_new,
                                  const float* sv_new, int64_t snew_stride,
                                  int B, int H, int Tq, int Tk, int Tk_cap, int mask_kind, const uint8_t* key_mask,
                                  int64_t mask_stride, int q_pos0, const int32_t* step_dev,
                                  float* ctx, int64_t ld_ctx, int8_t* ctx_q, float* ctx_s, uint8_t* probs_q,
                                  const OtFault* fault, const OtFault* faults_dev, const int32_t* unit_fault_dev, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(q && sq && k && v && sk && sv, "null operand");
  OT_REQUIRE(H == kHeads, "this build is specialised for 8 heads of 64 features (model.py:15-16)");
  OT_REQUIRE(B > 0 && Tq > 0, "empty problem");
  OT_REQUIRE(ldq % 16 == 0 && ldk % 16 == 0 && (k_new == nullptr || ld_new % 16 == 0), "row pitches must be multiples of 16");
  OT_REQUIRE(mask_kind >= 0 && mask_kind <= 2 && (mask_kind != 1 || key_mask), "bad mask");
  OT_REQUIRE((ctx_q == nullptr) == (ctx_s == nullptr), "ctx_q and ctx_s go together");
  OT_REQUIRE(ctx || ctx_q, "no output requested");
  OT_REQUIRE((k_new == nullptr) == (v_new == nullptr) && (k_new == nullptr || (sk_new && sv_new)), "k_new/v_new/sk_new/sv_new go together");
  const int tk_max = step_dev ? Tk_cap : Tk;  // with a device-side step the launch is sized for the cache capacity
  OT_REQUIRE(tk_max > 0 && tk_max <= kMaxTk && Tk <= Tk_cap, "Tk must be in (0, 192] and <= Tk_cap");
  OT_REQUIRE(probs_q == nullptr || step_dev == nullptr, "probs_q needs a host-side Tk");
  AttnArgs a = {};
  a.q = q; a.ldq = ldq; a.sq = sq; a.sq_stride = sq_stride;
  a.k = k; a.v = v; a.ldk = ldk; a.sk = sk; a.sv = sv; a.skv_stride = skv_stride;
  a.k_new = k_new; a.v_new = v_new; a.ld_new = ld_new; a.sk_new = sk_new; a.sv_new = sv_new; a.snew_stride = snew_stride;
  a.B = B; a.Tq = Tq; a.Tk = Tk; a.Tk_cap = Tk_cap; a.mask_kind = mask_kind;
  a.key_mask = key_mask; a.mask_stride = mask_stride; a.q_pos0 = q_pos0; a.step_dev = step_dev;
  a.ctx = ctx; a.ld_ctx = ld_ctx; a.ctx_q = ctx_q; a.ctx_s = ctx_s; a.probs_q = probs_q;
  if (fault) a.fault = *fault; else a.fault.mode = OT_FAULT_NONE;
  OT_REQUIRE((faults_dev == nullptr) == (unit_fault_dev == nullptr) && !(fault && unit_fault_dev), "bad batched-fault arguments");
  a.mf_faults = faults_dev; a.mf_unit = unit_fault_dev;
  cudaStream_t s = as_stream(stream);
  if (Tq == 1 && tk_max <= 32 * kDecKeysPerLane && a.fault.mode == OT_FAULT_NONE && a.mf_unit == nullptr && probs_q == nullptr) {
    OT_CHECK_CUDA(launch_kernel(attention_decode_kernel, dim3(B), dim3(256), 0, s, 1, a));
    count_launch();
    return OT_OK;
  }
  // encoder-size blocks (32 <= Tq <= 128, Tk <= 128), faulty or not: tensor-core kernel, one CTA per (sentence, head); the merged rows
  // are complete only across the 8 head CTAs, so the RowQuant for the O-projection is a second launch (or, without an fp32 context
  // buffer / with OT_ATTN_FUSE_Q=1, a cluster exchange inside the kernel: slower, see ot_attention_tc.cu)
  {
    bool fused_q = false;
    const int rc = launch_attention_tc(a, s, &fused_q);
    if (rc < 0) return rc;
    if (rc == 0) {
      if (ctx_q != nullptr && !fused_q) return ot_rowquant(ctx, ld_ctx, static_cast<int64_t>(B) * Tq, kDm, kDm, ctx_q, ctx_s, nullptr, stream);
      return OT_OK;
    }
  }
  if (Tq >= 32 && ctx != nullptr && a.fault.mode == OT_FAULT_NONE && a.mf_unit == nullptr && probs_q == nullptr && k_new == nullptr &&
      step_dev == nullptr && attn_heads_smem_bytes(Tq, Tk) <= 100 * 1024) {
    const size_t smem = attn_heads_smem_bytes(Tq, Tk);
    static DeviceOnce configured;
    if (configured.need(smem))
      OT_CHECK_CUDA(cudaFuncSetAttribute(attention_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    OT_CHECK_CUDA(launch_kernel(attention_heads_kernel, dim3(kHeads, B), dim3(256), smem, s, 1, a));
    count_launch();
    if (ctx_q != nullptr)     // the merged rows are complete only across the 8 head CTAs: RowQuant as a second launch
      return ot_rowquant(ctx, ld_ctx, static_cast<int64_t>(B) * Tq, kDm, kDm, ctx_q, ctx_s, nullptr, stream);
    return OT_OK;
  }
  if (Tq == 1) return launch_attention<1>(a, tk_max, s);
  // short sequences: tiles of 8 queries keep the CTA under half of the shared memory, so two CTAs (16 warps) share an SM -- the
  // kernel is latency-bound at 8 warps per SM -- and the grid has 4x more CTAs to fill the 148 SMs with
  if (Tq > 1 && attn_smem_bytes<8>(tk_max) <= 113 * 1024 && static_cast<int64_t>(B) * ((Tq + 31) / 32) < 4 * 148)
    return launch_attention<8>(a, tk_max, s);
  if (Tq <= 16 || attn_smem_bytes<32>(tk_max) > 227 * 1024) return launch_attention<16>(a, tk_max, s);
  return launch_attention<32>(a, tk_max, s);
}
