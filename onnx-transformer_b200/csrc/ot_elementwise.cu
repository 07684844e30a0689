// One CUDA handler per remaining ONNX op name of the reference graph (SURVEY.md 8a / App. B histogram):
// Abs, Relu, Sqrt, Round, Add, Sub, Mul, Div, Clip, ReduceMax, ReduceMean, Softmax, Where, Equal, Cast,
// Transpose and the float MatMul.  These serve the un-fused node-by-node walk of execute_node
// (onnx_optimized_inference.py:18-57), which the fault traces need (they re-run single nodes on delta_4d,
// :84-104).  All are HBM-bound streaming kernels with IEEE fp32 semantics (no fast math, no FMA contraction).
#include "ot_common.h"

namespace ot {

static inline unsigned grid_for(int64_t n, int threads = 256) {
  return static_cast<unsigned>(std::min<int64_t>((n + threads - 1) / threads, 148 * 32));
}

__global__ void __launch_bounds__(256) unary_kernel(int op, const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    float r;
    switch (op) {
      case 0: r = fabsf(v); break;
      case 1: r = fmaxf(v, 0.f); break;   // NaN-propagation differs from ORT only for NaN inputs
      case 2: r = __fsqrt_rn(v); break;
      case 3: r = rintf(v); break;        // ONNX Round: half to even
      case 4: r = -v; break;
      case 5: r = expf(v); break;
      default: r = v; break;
    }
    y[i] = r;
  }
}

struct Bcast4 {
  int64_t out[4];
  int64_t sa[4];
  int64_t sb[4];
};

__device__ __forceinline__ float apply_binary(int op, float a, float b) {
  switch (op) {
    case 0: return __fadd_rn(a, b);
    case 1: return __fsub_rn(a, b);
    case 2: return __fmul_rn(a, b);
    case 3: return __fdiv_rn(a, b);
    case 4: return fmaxf(a, b);
    default: return fminf(a, b);
  }
}

__global__ void __launch_bounds__(256) binary_kernel(int op, const float* __restrict__ a, const float* __restrict__ b,
                                                     float* __restrict__ out, Bcast4 s, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t r = i;
    const int64_t i3 = r % s.out[3]; r /= s.out[3];
    const int64_t i2 = r % s.out[2]; r /= s.out[2];
    const int64_t i1 = r % s.out[1]; r /= s.out[1];
    const int64_t i0 = r;
    const float va = a[i0 * s.sa[0] + i1 * s.sa[1] + i2 * s.sa[2] + i3 * s.sa[3]];
    const float vb = b[i0 * s.sb[0] + i1 * s.sb[1] + i2 * s.sb[2] + i3 * s.sb[3]];
    out[i] = apply_binary(op, va, vb);
  }
}

__global__ void __launch_bounds__(256) clip_kernel(const float* __restrict__ x, float lo, float hi, float* __restrict__ y, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[i] = fminf(fmaxf(x[i], lo), hi);
}

// one warp per row
__global__ void __launch_bounds__(256) reduce_last_kernel(int op, const float* __restrict__ x, int64_t rows, int n, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * n;
  float acc = op == 0 ? -INFINITY : 0.f;
  for (int i = lane; i < n; i += 32) acc = op == 0 ? fmaxf(acc, xr[i]) : acc + xr[i];
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, acc, o);
    acc = op == 0 ? fmaxf(acc, t) : acc + t;
  }
  if (lane == 0) y[row] = op == 0 ? acc : __fdiv_rn(acc, static_cast<float>(n));
}

__global__ void __launch_bounds__(256) softmax_kernel(const float* __restrict__ x, int64_t rows, int n, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * n;
  float* yr = y + row * n;
  float mx = -INFINITY;
  for (int i = lane; i < n; i += 32) mx = fmaxf(mx, xr[i]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int i = lane; i < n; i += 32) sum += expf(__fsub_rn(xr[i], mx));
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  for (int i = lane; i < n; i += 32) yr[i] = __fdiv_rn(expf(__fsub_rn(xr[i], mx)), sum);
}

__global__ void __launch_bounds__(256) where_kernel(const uint8_t* __restrict__ cond, float a_scalar, const float* __restrict__ x,
                                                    float* __restrict__ out, Bcast4 s, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t r = i;
    const int64_t i3 = r % s.out[3]; r /= s.out[3];
    const int64_t i2 = r % s.out[2]; r /= s.out[2];
    const int64_t i1 = r % s.out[1]; r /= s.out[1];
    const int64_t i0 = r;
    const uint8_t c = cond[i0 * s.sa[0] + i1 * s.sa[1] + i2 * s.sa[2] + i3 * s.sa[3]];
    out[i] = c ? a_scalar : x[i];
  }
}

__global__ void __launch_bounds__(256) equal_i64_kernel(const int64_t* __restrict__ x, int64_t scalar, uint8_t* __restrict__ out, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = x[i] == scalar ? 1 : 0;
}

template <typename S>
__device__ __forceinline__ void store_as(int kind, void* dst, int64_t i, S v) {
  switch (kind) {
    case 0: reinterpret_cast<float*>(dst)[i] = static_cast<float>(v); break;
    case 1: reinterpret_cast<int64_t*>(dst)[i] = static_cast<int64_t>(v); break;
    case 2: reinterpret_cast<uint8_t*>(dst)[i] = (v != S(0)) ? 1 : 0; break;   // Cast to bool
    case 3: reinterpret_cast<int8_t*>(dst)[i] = static_cast<int8_t>(v); break;
    case 5: reinterpret_cast<uint8_t*>(dst)[i] = static_cast<uint8_t>(v); break;   // numeric uint8 (QuantizeLinear results)
    default: reinterpret_cast<int32_t*>(dst)[i] = static_cast<int32_t>(v); break;
  }
}

__global__ void __launch_bounds__(256) cast_kernel(int sk, const void* __restrict__ src, int dk, void* __restrict__ dst, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    switch (sk) {
      case 0: store_as<float>(dk, dst, i, reinterpret_cast<const float*>(src)[i]); break;
      case 1: store_as<int64_t>(dk, dst, i, reinterpret_cast<const int64_t*>(src)[i]); break;
      case 2: store_as<int>(dk, dst, i, reinterpret_cast<const uint8_t*>(src)[i]); break;
      case 3: store_as<int>(dk, dst, i, reinterpret_cast<const int8_t*>(src)[i]); break;
      default: store_as<int>(dk, dst, i, reinterpret_cast<const int32_t*>(src)[i]); break;
    }
  }
}

struct Perm4 {
  int64_t out[4];
  int64_t stride_in[4];  // input stride of each OUTPUT dimension
};

__global__ void __launch_bounds__(256) transpose4_kernel(const uint32_t* __restrict__ x, uint32_t* __restrict__ y, Perm4 p, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t r = i;
    const int64_t i3 = r % p.out[3]; r /= p.out[3];
    const int64_t i2 = r % p.out[2]; r /= p.out[2];
    const int64_t i1 = r % p.out[1]; r /= p.out[1];
    const int64_t i0 = r;
    y[i] = x[i0 * p.stride_in[0] + i1 * p.stride_in[1] + i2 * p.stride_in[2] + i3 * p.stride_in[3]];
  }
}

// Batched fp32 MatMul, 64x64 tiles, 16-deep k-chunks, 4x4 micro-tile per thread, explicit fmaf.
__global__ void __launch_bounds__(256) matmul_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                         int M, int N, int K, int64_t sA, int64_t sB, int64_t sC) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const float* a = A + static_cast<int64_t>(blockIdx.z) * sA;
  const float* b = B + static_cast<int64_t>(blockIdx.z) * sB;
  float* c = C + static_cast<int64_t>(blockIdx.z) * sC;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
      const int m = idx >> 4, k = idx & 15;
      As[k][m] = (m0 + m < M && k0 + k < K) ? a[static_cast<int64_t>(m0 + m) * K + k0 + k] : 0.f;
    }
    for (int idx = threadIdx.x; idx < 16 * 64; idx += 256) {
      const int k = idx >> 6, n = idx & 63;
      Bs[k][n] = (k0 + k < K && n0 + n < N) ? b[static_cast<int64_t>(k0 + k) * N + n0 + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[k][ty * 4 + i]; bv[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) c[static_cast<int64_t>(m) * N + n] = acc[i][j];
    }
}

static int make_bcast(const int64_t a_shape[4], const int64_t b_shape[4], const int64_t out_shape[4], Bcast4* s) {
  for (int d = 0; d < 4; ++d) {
    s->out[d] = out_shape[d];
    if (!(a_shape[d] == out_shape[d] || a_shape[d] == 1)) return -1;
    if (b_shape && !(b_shape[d] == out_shape[d] || b_shape[d] == 1)) return -1;
  }
  int64_t sa = 1, sb = 1;
  for (int d = 3; d >= 0; --d) {
    s->sa[d] = a_shape[d] == 1 ? 0 : sa;
    sa *= a_shape[d];
    if (b_shape) {
      s->sb[d] = b_shape[d] == 1 ? 0 : sb;
      sb *= b_shape[d];
    } else {
      s->sb[d] = 0;
    }
  }
  return 0;
}

}  // namespace ot

using namespace ot;

extern "C" int ot_unary_f32(int op, const float* x, float* y, int64_t n, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(x && y && n >= 0 && op >= 0 && op <= 6, "bad unary arguments");
  if (n == 0) return OT_OK;
  unary_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(op, x, y, n);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_binary_f32(int op, const float* a, const int64_t a_shape[4], const float* b, const int64_t b_shape[4], float* out,
                             const int64_t out_shape[4], void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(a && b && out && op >= 0 && op <= 5, "bad binary arguments");
  Bcast4 s;
  OT_REQUIRE(make_bcast(a_shape, b_shape, out_shape, &s) == 0, "shapes are not broadcast-compatible");
  const int64_t n = out_shape[0] * out_shape[1] * out_shape[2] * out_shape[3];
  if (n == 0) return OT_OK;
  binary_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(op, a, b, out, s, n);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_clip_f32(const float* x, float lo, float hi, float* y, int64_t n, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(x && y && n >= 0, "bad clip arguments");
  if (n == 0) return OT_OK;
  clip_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(x, lo, hi, y, n);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_reduce_last_f32(int op, const float* x, int64_t rows, int n, float* y, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(x && y && rows >= 0 && n > 0 && (op == 0 || op == 1), "bad reduce arguments");
  if (rows == 0) return OT_OK;
  reduce_last_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, as_stream(stream)>>>(op, x, rows, n, y);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_softmax_f32(const float* x, int64_t rows, int n, float* y, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(x && y && rows >= 0 && n > 0, "bad softmax arguments");
  if (rows == 0) return OT_OK;
  softmax_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, as_stream(stream)>>>(x, rows, n, y);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_where_f32(const uint8_t* cond, const int64_t c_shape[4], float a_scalar, const float* x, float* out,
                            const int64_t out_shape[4], void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(cond && x && out, "bad where arguments");
  Bcast4 s;
  OT_REQUIRE(make_bcast(c_shape, nullptr, out_shape, &s) == 0, "condition is not broadcast-compatible");
  const int64_t n = out_shape[0] * out_shape[1] * out_shape[2] * out_shape[3];
  if (n == 0) return OT_OK;
  where_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(cond, a_scalar, x, out, s, n);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_equal_i64(const int64_t* x, int64_t scalar, uint8_t* out, int64_t n, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(x && out && n >= 0, "bad equal arguments");
  if (n == 0) return OT_OK;
  equal_i64_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(x, scalar, out, n);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_cast(int src_kind, const void* src, int dst_kind, void* dst, int64_t n, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(src && dst && n >= 0 && src_kind >= 0 && src_kind <= 4 && dst_kind >= 0 && dst_kind <= 5, "bad cast arguments");
  if (n == 0) return OT_OK;
  cast_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(src_kind, src, dst_kind, dst, n);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_transpose4_b32(const void* x, const int64_t shape[4], const int perm[4], void* y, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(x && y, "bad transpose arguments");
  int64_t in_stride[4];
  int64_t st = 1;
  for (int d = 3; d >= 0; --d) { in_stride[d] = st; st *= shape[d]; }
  Perm4 p;
  bool seen[4] = {false, false, false, false};
  for (int d = 0; d < 4; ++d) {
    OT_REQUIRE(perm[d] >= 0 && perm[d] < 4 && !seen[perm[d]], "perm is not a permutation");
    seen[perm[d]] = true;
    p.out[d] = shape[perm[d]];
    p.stride_in[d] = in_stride[perm[d]];
  }
  if (st == 0) return OT_OK;
  transpose4_kernel<<<grid_for(st), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint32_t*>(x), reinterpret_cast<uint32_t*>(y), p, st);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}

extern "C" int ot_matmul_f32(const float* A, const float* B, float* C, int batch, int M, int N, int K, int64_t strideA, int64_t strideB,
                             int64_t strideC, void* stream) {
  OT_REQUIRE_DEVICE();
  OT_REQUIRE(A && B && C && batch > 0 && M > 0 && N > 0 && K > 0 && batch <= 65535, "bad matmul arguments");
  dim3 grid((N + 63) / 64, (M + 63) / 64, batch);
  matmul_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(A, B, C, M, N, K, strideA, strideB, strideC);
  OT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OT_OK;
}
