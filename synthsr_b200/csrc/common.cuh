// Shared helpers for libsynthsr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#define SSR_OK 0
#define SSR_ERR_ARG -1
#define SSR_ERR_CUDA -2
#define SSR_ERR_UNSUPPORTED -3

void ssr_set_error(const char* fmt, ...);

#define SSR_CHECK_ARG(cond, msg)                                                                  \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      ssr_set_error("%s:%d: invalid argument: %s (%s)", __FILE__, __LINE__, msg, #cond);          \
      return SSR_ERR_ARG;                                                                         \
    }                                                                                             \
  } while (0)

#define SSR_CHECK_CUDA(expr)                                                                      \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      ssr_set_error("%s:%d: CUDA error %d (%s) in %s", __FILE__, __LINE__, (int)e__,              \
                    cudaGetErrorString(e__), #expr);                                              \
      return SSR_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

#define SSR_CHECK_LAUNCH() SSR_CHECK_CUDA(cudaGetLastError())

static inline int ssr_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// Number of kernel launches issued by this library since process start (bench.py reports it as gpu_launches).
extern unsigned long long g_ssr_launch_count;
#define SSR_COUNT_LAUNCH() (++g_ssr_launch_count)
