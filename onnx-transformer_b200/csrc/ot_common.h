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

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// 8-bit two's-complement bit flip with wrap, as inject_utils/layers.py:61-68 (flip_int8_bit).
__host__ __device__ inline int flip_int8_bit(int value, int bit) {
  int flipped = value ^ (1 << bit);
  if (flipped > 127) flipped -= 256;
  if (flipped < -128) flipped += 256;
  return flipped;
}

}  // namespace ot
