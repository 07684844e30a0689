// Library-level entry points of libot_b200.so: version, error slot, device probe, launch counter.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "ot_common.h"

namespace ot {

static thread_local char g_error[1024] = "";
static std::atomic<int64_t> g_launches{0};
static std::atomic<int> g_pdl{0};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }
void set_pdl(int on) { g_pdl.store(on); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool device_is_sm100() {
  static int cached = -1;
  if (cached >= 0) return cached == 1;
  int dev = 0, major = 0, count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    cached = 0;
    return false;
  }
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
  cached = (major == 10) ? 1 : 0;
  return cached == 1;
}

}  // namespace ot

extern "C" int ot_set_timeline(unsigned long long* buf, unsigned int capacity) {
  ot::tl_set_gemm(buf, capacity);
  ot::tl_set_attention(buf, capacity);
  ot::tl_set_rowops(buf, capacity);
  ot::tl_set_generator(buf, capacity);
  return cudaGetLastError() == cudaSuccess ? OT_OK : OT_ECUDA;
}
extern "C" int ot_set_pdl(int enable) { ot::set_pdl(enable ? 1 : 0); return OT_OK; }
extern "C" int ot_version(void) { return 100; }
extern "C" const char* ot_last_error(void) { return ot::g_error; }
extern "C" int ot_device_ok(void) { return ot::device_is_sm100() ? 1 : 0; }
extern "C" int64_t ot_launch_count(void) { return ot::g_launches.load(std::memory_order_relaxed); }
