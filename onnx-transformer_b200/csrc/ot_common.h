// Shared host-side helpers of libot_b200.so: error slot, launch counter, argument checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ot_b200.h"

namespace ot {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
bool device_is_sm100();
void set_pdl(int on);
void tl_set_gemm(unsigned long long*, unsigned int);
void tl_set_attention(unsigned long long*, unsigned int);
void tl_set_rowops(unsigned long long*, unsigned int);
void tl_set_generator(unsigned long long*, unsigned int);

#define OT_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::ot::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return OT_ECUDA;                                                                       \
    }                                                                                        \
  } while (0)

#define OT_REQUIRE(cond, msg)                                        \
  do {                                                               \
    if (!(cond)) {                                                   \
      ::ot::set_error("%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond); \
      return OT_EINVAL;                                              \
    }                                                                \
  } while (0)

#define OT_REQUIRE_DEVICE()                                          \
  do {                                                               \
    if (!::ot::device_is_sm100()) {                                  \
      ::ot::set_error("no sm_100 CUDA device available (no CPU fallback exists)"); \
      return OT_ENODEV;                                              \
    }                                                                \
  } while (0)

// Programmatic dependent launch (PDL): when enabled, a kernel is launched with programmaticStreamSerialization so that its
// launch latency and prologue (barrier init, TMEM alloc, descriptor prefetch) overlap the tail of its predecessor; every
// PDL-launched kernel executes pdl_wait() before it touches global memory produced upstream.
bool pdl_enabled();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Optional device-side timeline (profiling aid, ot_set_timeline): block 0 / thread 0 of every kernel records
// {kernel id, t_start, t_after_pdl_wait, t_end} from %globaltimer.  Each translation unit holds its own copy of the pointer.
static __device__ unsigned long long* tl_buf = nullptr;
static __device__ unsigned int tl_cap = 0;
struct TlMark {
  unsigned int slot;
};
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool tl_on() { return tl_buf != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0; }
__device__ __forceinline__ unsigned int tl_begin(int kernel_id) {
  if (!tl_on()) return 0xffffffffu;
  const unsigned int slot = static_cast<unsigned int>(atomicAdd(tl_buf, 1ull));
  if (slot >= tl_cap) return 0xffffffffu;
  tl_buf[1 + slot * 4] = static_cast<unsigned long long>(kernel_id);
  tl_buf[2 + slot * 4] = tl_now();
  return slot;
}
__device__ __forceinline__ void tl_mark(unsigned int slot, int which) {
  if (slot != 0xffffffffu) tl_buf[1 + slot * 4 + which] = tl_now();
}
#define OT_DEFINE_TL_SETTER(name)                                                        \
  void name(unsigned long long* buf, unsigned int cap) {                                 \
    cudaMemcpyToSymbol(tl_buf, &buf, sizeof(buf));                                       \
    cudaMemcpyToSymbol(tl_cap, &cap, sizeof(cap));                                       \
  }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                                        Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// "Configure this kernel once": cudaFuncSetAttribute applies to the CURRENT device only, so the once-flag is per device
// (one process per GPU is the normal deployment; a process that drives several devices must not skip the attribute on the second).
struct DeviceOnce {
  size_t level[64] = {};
  // true when the attribute has not been set on the current device yet, or was set for less than `want`
  bool need(size_t want = 1) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (level[dev] >= want) return false;
    level[dev] = want;
    return true;
  }
};
static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// 8-bit two's-complement bit flip with wrap, as inject_utils/layers.py:61-68 (flip_int8_bit).
__host__ __device__ inline int flip_int8_bit(int value, int bit) {
  int flipped = value ^ (1 << bit);
  if (flipped > 127) flipped -= 256;
  if (flipped < -128) flipped += 256;
  return flipped;
}

// 4-bit two's-complement bit flip with wrap to [-8, 7], as inject_utils/layers.py:48-59 (flip_int4_bit).
__host__ __device__ inline int flip_int4_bit(int value, int bit) {
  int flipped = value ^ (1 << bit);
  if (flipped > 7) flipped -= 16;
  if (flipped < -8) flipped += 16;
  return flipped;
}

}  // namespace ot
