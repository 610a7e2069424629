// Error reporting and library identity for the C ABI declared in include/fpmatch.h.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

extern "C" void fpm_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

extern "C" const char* fpm_last_error(void) { return g_err; }

extern "C" int fpm_abi_version(void) { return 1; }

// 1 when a CUDA device of compute capability 10.x is current, 0 otherwise (no fallback exists:
// callers must fail loudly on 0).
extern "C" int fpm_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
