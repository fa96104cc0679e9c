// Library-level entry points of libsynthsr_b200: error reporting, launch accounting, device query.
#include "common.cuh"
#include <cstdarg>

unsigned long long g_ssr_launch_count = 0;

static thread_local char g_err[1024] = "";

void ssr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

const char* ssr_last_error(void) { return g_err; }

unsigned long long ssr_launch_count(void) { return g_ssr_launch_count; }

int ssr_abi_version(void) { return 1; }

// returns compute capability major*10+minor of the current device, or a negative error code
int ssr_device_arch(void) {
  int dev = 0;
  SSR_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  SSR_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SSR_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}

}  // extern "C"
