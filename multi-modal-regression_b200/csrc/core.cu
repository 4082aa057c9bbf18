// Library plumbing: ABI version, thread-local error string, device attribute cache.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void bdp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int bdp_num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached_dev = dev;
    cached_sms = n;
  }
  return cached_sms;
}

extern "C" int bdp_abi_version(void) { return BDP_ABI_VERSION; }
extern "C" const char* bdp_last_error(void) { return g_err; }
extern "C" int bdp_sm_count(void) { return bdp_num_sms(); }
