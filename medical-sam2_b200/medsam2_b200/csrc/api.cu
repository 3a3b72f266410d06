// Error reporting + version for the C-ABI (include/medsam2_b200.h).
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

extern "C" void ms2_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* ms2_last_error(void) { return g_err; }
extern "C" int ms2_version(void) { return 100; }

extern "C" int ms2_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}
