// Library-wide entry points of the C-ABI (include/xdet_b200.h).
#include "common.cuh"

namespace xdet {

std::atomic<long long> g_launches{0};

char* tls_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace xdet

extern "C" const char* xdet_last_error(void) { return xdet::tls_error_buf(); }
extern "C" const char* xdet_version(void) { return "xdet_b200 0.1 sm_100a"; }
extern "C" long long xdet_launch_count(void) { return xdet::g_launches.load(); }
