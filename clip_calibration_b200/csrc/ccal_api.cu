// Library-wide pieces of the C ABI: version, thread-local error text, device check.
#include "ccal_common.cuh"

#include <math.h>
#include <string.h>

#include <atomic>

namespace ccal {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static std::atomic<long long> g_kernel_launches{0};
void note_launch(int n) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

float ceil_to_f32(double t) {
  float f = (float)t;                       // round to nearest
  if ((double)f < t) f = nextafterf(f, INFINITY);
  return f;
}

}  // namespace ccal

extern "C" int ccal_version(void) { return CCAL_VERSION; }

extern "C" long long ccal_launch_count(void) { return ccal::g_kernel_launches.load(std::memory_order_relaxed); }

extern "C" const char* ccal_last_error(void) { return ccal::error_buffer(); }

extern "C" int ccal_check_device(void) {
  int dev = 0;
  CCAL_CUDA_OK(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  CCAL_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CCAL_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10)
    return ccal::fail(CCAL_ERR_UNSUPPORTED,
                      "libccal is built for sm_100a only; device %d is sm_%d%d (no fallback path exists)",
                      dev, major, minor);
  return CCAL_OK;
}
