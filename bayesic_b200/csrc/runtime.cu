// Error text, device info and launch counting shared by all translation units.
#include "common.cuh"

namespace bb {

thread_local int64_t g_launch_count = 0;
static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

const char* get_error() { return g_error; }

int device_sm_count() {
  static int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached_dev = dev;
    cached = n;
  }
  return cached;
}

}  // namespace bb
