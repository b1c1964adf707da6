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

int nvit_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

extern "C" const char* nvit_last_error(void) { return g_err; }
extern "C" int nvit_version(void) { return 100; }
extern "C" int nvit_sm_count(void) { return nvit_num_sms(); }
