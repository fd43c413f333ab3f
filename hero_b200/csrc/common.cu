#include "common.h"

namespace hero {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return status;
}

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -HERO_ERR_NO_DEVICE;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return -HERO_ERR_NO_DEVICE;
  cached = n;
  return n;
}

}  // namespace hero

extern "C" {

const char* hero_last_error(void) { return hero::error_buffer(); }

int hero_version(void) { return 100; }

int hero_sm_count(void) {
  int n = hero::sm_count();
  if (n < 0) hero::set_error(HERO_ERR_NO_DEVICE, "no CUDA device available");
  return n;
}

}  // extern "C"
