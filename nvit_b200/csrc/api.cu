// Library-level plumbing of the C ABI declared in include/nvit_b200.h: error string, version, device info.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void nvit_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<int> g_sm_budget{0};  // 0 = all SMs

// Number of SMs the persistent kernels size their grids for: the device's SM count, or the budget set through
// nvit_set_sm_budget (data-parallel runs leave a few SMs to the NCCL kernels that overlap the backward pass).
int nvit_num_sms() {
  static std::atomic<int> cached[128];      // per device; 0 = not queried yet
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) dev = -1;
  if (dev >= 0) n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (dev < 0 || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    if (dev >= 0) cached[dev].store(n, std::memory_order_relaxed);
  }
  const int budget = g_sm_budget.load(std::memory_order_relaxed);
  return (budget > 0 && budget < n) ? budget : n;
}

// Programmatic dependent launch (common.cuh: pdl_wait / nvit::launch).  Off unless nvit_set_pdl(1).
static std::atomic<int> g_pdl{0};
int nvit_pdl_enabled() { return g_pdl.load(std::memory_order_relaxed); }
extern "C" int nvit_set_pdl(int on) {
  g_pdl.store(on ? 1 : 0, std::memory_order_relaxed);
  return NVIT_OK;
}

extern "C" const char* nvit_last_error(void) { return g_err; }
extern "C" int nvit_version(void) { return 101; }   // 101: + nvit_augment_u8, nvit_tmap_cache_stats
extern "C" int nvit_sm_count(void) { return nvit_num_sms(); }
extern "C" int nvit_set_sm_budget(int n) {
  if (n < 0) {
    nvit_set_error("nvit_set_sm_budget: negative budget");
    return NVIT_ERR_ARG;
  }
  g_sm_budget.store(n & ~1, std::memory_order_relaxed);  // CTA pairs: keep it even
  return NVIT_OK;
}
