// Shared host-side plumbing for the C-ABI: per-thread error string, launch accounting, checks.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/xdet_b200.h"

namespace xdet {

char* tls_error_buf();                       // 512-byte per-thread buffer
int fail(int code, const char* fmt, ...);    // formats into the buffer, returns `code`
extern std::atomic<long long> g_launches;    // kernels launched by this library

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return XDET_OK;
  return fail(XDET_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

// Call right after a <<<>>> launch.
inline int after_launch(const char* kernel) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), kernel);
}

constexpr int kNumSMs = 148;  // B200

}  // namespace xdet

#define XDET_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != XDET_OK) return _rc; \
  } while (0)
