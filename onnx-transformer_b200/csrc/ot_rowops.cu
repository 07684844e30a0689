// Row-local, HBM-bound kernels: LayerNorm(+quant), RowQuant, residual Add, embedding+PE, int4 unpack,
// greedy-loop token append.  One warp per row, 128-bit coalesced accesses, warp-shuffle reductions.
// fp32 op order follows SURVEY.md App. A (the order of the ops the reference exports); compiled with
// -fmad=false, explicit *_rn intrinsics where the order matters for bit-exact integer results.
#include "ot_common.h"
#include "ot_rowmath.cuh"

namespace ot {

// ------------------------------------------------------------------------------------------------
// LayerNorm (+ optional RowQuant of y); the row math lives in ot_rowmath.cuh.   VEC = n / 128 float4 per lane.
template <int VEC>
__global__ void __launch_bounds__(256) layernorm_quant_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int64_t rows, int n, float eps,
                                                              float* __restrict__ y_out, int8_t* __restrict__ q_out,
                                                              float* __restrict__ s_out) {
  const unsigned int tl = tl_begin(4);
  pdl_wait();      // upstream results are complete and visible from here on
  pdl_trigger();   // now let exactly one successor start its launch + prologue (look-ahead depth 1)
  tl_mark(tl, 2);
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * n);
  float4 v[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = __ldg(xr + i * 32 + lane);
  const float amax = layernorm_row<VEC>(v, lane, n, gamma, beta, eps);
  if (y_out) {
    float4* yr = reinterpret_cast<float4*>(y_out + row * n);
#pragma unroll
    for (int i = 0; i < VEC; ++i) yr[i * 32 + lane] = v[i];
  }
  if (q_out) {
    const float s = quant_scale_x(warp_max_nonneg(amax));      // exact equivalents of the division / shuffle-tree forms (ot_rowmath.cuh)
    const float s_rcp = __frcp_rn(s);
    uint32_t* qr = reinterpret_cast<uint32_t*>(q_out + row * n);
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      qr[i * 32 + lane] = quant4_pack(v[i], s, s_rcp);
    if (lane == 0) s_out[row] = s;
  }
  tl_mark(tl, 3);
}

// ------------------------------------------------------------------------------------------------
// RowQuant over groups of `group` columns; one warp per (row, group); two passes over L1/L2-resident data.
__global__ void __launch_bounds__(256) rowquant_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int n, int group,
                                                       int8_t* __restrict__ q, float* __restrict__ s_out,
                                                       float* __restrict__ xhat) {
  pdl_wait();      // upstream results are complete and visible from here on
  pdl_trigger();   // now let exactly one successor start its launch + prologue (look-ahead depth 1)
  const int lane = threadIdx.x & 31;
  const int groups = n / group;
  const int64_t item = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= rows * groups) return;
  const int64_t row = item / groups;
  const int gidx = static_cast<int>(item % groups);
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx + static_cast<int64_t>(gidx) * group);
  const int nvec = group >> 2;
  float amax = 0.f;
  for (int i = lane; i < nvec; i += 32) {
    const float4 v = __ldg(xr + i);
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  const float s = quant_scale_x(warp_max_nonneg(amax));
  const float s_rcp = __frcp_rn(s);
  uint32_t* qr = reinterpret_cast<uint32_t*>(q + row * n + static_cast<int64_t>(gidx) * group);
  float4* hr = xhat ? reinterpret_cast<float4*>(xhat + row * n + static_cast<int64_t>(gidx) * group) : nullptr;
  for (int i = lane; i < nvec; i += 32) {
    const float4 v = __ldg(xr + i);
    const uint32_t packed = quant4_pack(v, s, s_rcp);
    const int a = static_cast<int8_t>(packed & 0xFF), b = static_cast<int8_t>((packed >> 8) & 0xFF);
    const int c = static_cast<int8_t>((packed >> 16) & 0xFF), d = static_cast<int8_t>(packed >> 24);
    qr[i] = packed;
    if (hr) hr[i] = make_float4(__fmul_rn(__int2float_rn(a), s), __fmul_rn(__int2float_rn(b), s),
                                __fmul_rn(__int2float_rn(c), s), __fmul_rn(__int2float_rn(d), s));
  }
  if (lane == 0) s_out[row * groups + gidx] = s;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) residual_add_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                           float4* __restrict__ out, int64_t nvec) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 u = __ldg(a + i), v = __ldg(b + i);
    out[i] = make_float4(__fadd_rn(u.x, v.x), __fadd_rn(u.y, v.y), __fadd_rn(u.z, v.z), __fadd_rn(u.w, v.w));
  }
}

// ------------------------------------------------------------------------------------------------
// out[r,:] = table[id_r,:] * scale + pe[pos_r,:]   (embeddings.py:13, positional_encodings.py:24)
__global__ void __launch_bounds__(128) embed_pe_kernel(const int64_t* __restrict__ ids, int64_t ids_stride,
                                                       const float* __restrict__ table, const float* __restrict__ pe,
                                                       int64_t rows, int seq_len, int d, int pos0,
                                                       const int32_t* __restrict__ pos_dev, float scale, float* __restrict__ out) {
  const unsigned int tl = tl_begin(6);
  pdl_wait();      // upstream results are complete and visible from here on
  pdl_trigger();   // now let exactly one successor start its launch + prologue (look-ahead depth 1)
  tl_mark(tl, 2);
  const int64_t row = blockIdx.x;
  if (row >= rows) return;
  const int dyn = pos_dev ? *pos_dev : 0;
  const int64_t id = pos_dev ? ids[row * ids_stride + dyn] : ids[row * ids_stride];
  const int pos = (pos_dev ? dyn : pos0) + static_cast<int>(row % seq_len);
  const float4* t = reinterpret_cast<const float4*>(table + id * d);
  const float4* p = reinterpret_cast<const float4*>(pe + static_cast<int64_t>(pos) * d);
  float4* o = reinterpret_cast<float4*>(out + row * d);
  for (int i = threadIdx.x; i < (d >> 2); i += blockDim.x) {
    const float4 e = __ldg(t + i), q = __ldg(p + i);
    o[i] = make_float4(__fadd_rn(__fmul_rn(e.x, scale), q.x), __fadd_rn(__fmul_rn(e.y, scale), q.y),
                       __fadd_rn(__fmul_rn(e.z, scale), q.z), __fadd_rn(__fmul_rn(e.w, scale), q.w));
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unpack_int4_kernel(const uint8_t* __restrict__ w4, int8_t* __restrict__ w8, int64_t nbytes) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nbytes;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint8_t b = w4[i];
    const int lo = ((b & 0xF) ^ 8) - 8, hi = ((b >> 4) ^ 8) - 8;
    reinterpret_cast<uint16_t*>(w8)[i] = static_cast<uint16_t>((lo & 0xFF) | ((hi & 0xFF) << 8));
  }
}

__global__ void __launch_bounds__(256) pack_int4_kernel(const int8_t* __restrict__ w8, uint8_t* __restrict__ w4, int64_t nbytes) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nbytes;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint16_t pair = reinterpret_cast<const uint16_t*>(w8)[i];       // even k in the low byte
    w4[i] = static_cast<uint8_t>((pair & 0xF) | (((pair >> 8) & 0xF) << 4));
  }
}

// Row sums of an int8 matrix (zero-point correction terms of ONNX MatMulInteger / QLinearMatMul): warp per row, 16-byte loads where
// the row allows, dp4a against 0x01010101.
__global__ void __launch_bounds__(256) rowsum_i8_kernel(const int8_t* __restrict__ x, int64_t ld, int64_t rows, int cols, int32_t* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int8_t* r = x + row * ld;
  int acc = 0;
  const bool vec = ((reinterpret_cast<uintptr_t>(r) & 15) == 0);
  const int nv = vec ? (cols >> 4) : 0;
  for (int i = lane; i < nv; i += 32) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(r) + i);
    acc = __dp4a(static_cast<int>(v.x), 0x01010101, acc);
    acc = __dp4a(static_cast<int>(v.y), 0x01010101, acc);
    acc = __dp4a(static_cast<int>(v.z), 0x01010101, acc);
    acc = __dp4a(static_cast<int>(v.w), 0x01010101, acc);
  }
  for (int i = nv * 16 + lane; i < cols; i += 32) acc += r[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = acc;
}

// ys[b, step+1] = next[b]; step++  (greedy_decode: parallelized_inject_onnx_transformer.py:753-758)
__global__ void append_token_kernel(int64_t* __restrict__ ys, int64_t ld, const int64_t* __restrict__ next, int B,
                                    int32_t* __restrict__ step_dev) {
  const unsigned int tl = tl_begin(9);
  pdl_wait();      // upstream results are complete and visible from here on
  pdl_trigger();   // now let exactly one successor start its launch + prologue (look-ahead depth 1)
  tl_mark(tl, 2);
  const int step = *step_dev;
  for (int b = threadIdx.x; b < B; b += blockDim.x) ys[static_cast<int64_t>(b) * ld + step + 1] = next[b];
  __syncthreads();
  if (threadIdx.x == 0) *step_dev = step + 1;
  tl_mark(tl, 3);
}

OT_DEFINE_TL_SETTER(tl_set_rowops)

}  // namespace ot

using namespace ot;

extern "C" int ot_layernorm_quant(const float* x, const float* gamma, const float* beta, int64_t rows, int n, float eps,
                                  float* y_out, int8_t* q_out, float* s_out, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(x && gamma && beta && rows >= 0, "null operand");
  OT_REQUIRE(n % 128 == 0 && n >= 128 && n <= 2048, "n must be a multiple of 128 in [128, 2048]");
  OT_REQUIRE((q_out == nullptr) == (s_out == nullptr), "q_out and s_out go together");
  if (rows == 0) return OT_OK;
  const int warps = 8;
  const unsigned grid = static_cast<unsigned>((rows + warps - 1) / warps);
  cudaStream_t s = as_stream(stream);
#define OT_LN_CASE(V)                                                                                            \
  case V:                                                                                                        \
    OT_CHECK_CUDA(launch_kernel(layernorm_quant_kernel<V>, dim3(grid), dim3(warps * 32), 0, s, 1, x, gamma, beta, rows, n, eps, y_out, q_out, s_out)); \
    break;
  switch (n / 128) {
    OT_LN_CASE(1) OT_LN_CASE(2) OT_LN_CASE(3) OT_LN_CASE(4) OT_LN_CASE(5) OT_LN_CASE(6) OT_LN_CASE(7) OT_LN_CASE(8)
    OT_LN_CASE(9) OT_LN_CASE(10) OT_LN_CASE(11) OT_LN_CASE(12) OT_LN_CASE(13) OT_LN_CASE(14) OT_LN_CASE(15) OT_LN_CASE(16)
    default: OT_REQUIRE(false, "unsupported n");
  }
#undef OT_LN_CASE
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_rowquant(const float* x, int64_t ldx, int64_t rows, int n, int group, int8_t* q, float* s, float* xhat,
                           void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(x && q && s && rows >= 0, "null operand");
  OT_REQUIRE(group > 0 && n % group == 0 && group % 4 == 0 && ldx % 4 == 0, "group must divide n and be a multiple of 4");
  if (rows == 0) return OT_OK;
  const int64_t items = rows * (n / group);
  const int warps = 8;
  OT_CHECK_CUDA(launch_kernel(rowquant_kernel, dim3(static_cast<unsigned>((items + warps - 1) / warps)), dim3(warps * 32), 0, as_stream(stream), 1,
                              x, ldx, rows, n, group, q, s, xhat));
  count_launch();
  return OT_OK;
}

extern "C" int ot_rowsum_i8(const int8_t* X, int64_t ld, int64_t rows, int cols, int32_t* out, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(X && out && rows >= 0 && cols > 0 && ld >= cols, "bad rowsum arguments");
  if (rows == 0) return OT_OK;
  OT_CHECK_CUDA(launch_kernel(rowsum_i8_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, as_stream(stream), 1, X, ld, rows, cols, out));
  count_launch();
  return OT_OK;
}

extern "C" int ot_residual_add(const float* a, const float* b, float* out, int64_t n, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(a && b && out && n >= 0 && n % 4 == 0, "n must be a multiple of 4");
  if (n == 0) return OT_OK;
  const int64_t nvec = n / 4;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((nvec + 255) / 256, 148 * 16));
  residual_add_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                           reinterpret_cast<float4*>(out), nvec);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_embed_pe(const int64_t* ids, int64_t ids_stride, const float* table, const float* pe, int64_t rows, int seq_len,
                           int d, int pos0, const int32_t* pos_dev, float scale, float* out, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(ids && table && pe && out && rows >= 0 && seq_len > 0 && d % 4 == 0, "bad embed_pe arguments");
  if (rows == 0) return OT_OK;
  OT_CHECK_CUDA(launch_kernel(embed_pe_kernel, dim3(static_cast<unsigned>(rows)), dim3(128), 0, as_stream(stream), 1, ids, ids_stride, table, pe,
                              rows, seq_len, d, pos0, pos_dev, scale, out));
  count_launch();
  return OT_OK;
}

extern "C" int ot_unpack_int4(const uint8_t* W4, int8_t* W8, int64_t rows, int64_t cols, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(W4 && W8 && rows >= 0 && cols % 2 == 0, "cols must be even");
  const int64_t nbytes = rows * cols / 2;
  if (nbytes == 0) return OT_OK;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((nbytes + 255) / 256, 148 * 16));
  unpack_int4_kernel<<<grid, 256, 0, as_stream(stream)>>>(W4, W8, nbytes);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_pack_int4(const int8_t* W8, uint8_t* W4, int64_t rows, int64_t cols, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(W4 && W8 && rows >= 0 && cols % 2 == 0, "cols must be even");
  const int64_t nbytes = rows * cols / 2;
  if (nbytes == 0) return OT_OK;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((nbytes + 255) / 256, 148 * 16));
  pack_int4_kernel<<<grid, 256, 0, as_stream(stream)>>>(W8, W4, nbytes);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_append_token(int64_t* ys, int64_t ld_ys, const int64_t* next_ids, int B, int32_t* step_dev, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(ys && next_ids && step_dev && B > 0, "bad append_token arguments");
  OT_CHECK_CUDA(launch_kernel(append_token_kernel, dim3(1), dim3(256), 0, as_stream(stream), 1, ys, ld_ys, next_ids, B, step_dev));
  count_launch();
  return OT_OK;
}
