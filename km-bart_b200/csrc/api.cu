// Library-level entry points: version, architecture gate, last-error text.
#include <stdio.h>
#include <string.h>
#include "common.cuh"
#include "../../include/kmbart.h"

static thread_local char g_err[512] = "";

void kmb_set_last_error(const char* msg, const char* file, int line) {
  snprintf(g_err, sizeof g_err, "%s (%s:%d)", msg, file, line);
}

extern "C" const char* kmb_last_error(void) { return g_err; }
extern "C" int kmb_version(void) { return 100; }

extern "C" int kmb_arch_check(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    kmb_set_last_error("no CUDA device", __FILE__, __LINE__);
    return KMB_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    char m[96];
    snprintf(m, sizeof m, "device is sm_%d%d; libkmbart_sm100 requires sm_100 (B200)", major, minor);
    kmb_set_last_error(m, __FILE__, __LINE__);
    return KMB_ERR_ARCH;
  }
  return KMB_OK;
}
