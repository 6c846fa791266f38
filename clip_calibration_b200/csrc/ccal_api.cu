// Library-wide pieces of the C ABI: version, thread-local error text, device check.
#include "ccal_common.cuh"

#include <math.h>
#include <string.h>

#include <atomic>
#include <mutex>

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

namespace {
struct TraceMark { cudaEvent_t ev; int id; long long seq; unsigned long long stream; };
constexpr int kTraceMarks = 256;
TraceMark g_marks[kTraceMarks];
std::atomic<long long> g_mark_seq{0};
std::mutex g_mark_mu;
}  // namespace

void trace_mark(cudaStream_t stream, int id) {
  static const bool on = getenv("CCAL_TRACE_MARKS") != nullptr;
  if (!on) return;
  std::lock_guard<std::mutex> lock(g_mark_mu);
  const long long seq = g_mark_seq.fetch_add(1);
  TraceMark& m = g_marks[seq % kTraceMarks];
  if (m.ev == nullptr && cudaEventCreateWithFlags(&m.ev, cudaEventDisableTiming) != cudaSuccess) { m.ev = nullptr; return; }
  m.id = id; m.seq = seq; m.stream = (unsigned long long)(uintptr_t)stream;
  cudaEventRecord(m.ev, stream);
}

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

extern "C" int ccal_trace_marks_report(char* buf, int cap) {
  if (buf == nullptr || cap <= 0) return 0;
  std::lock_guard<std::mutex> lock(ccal::g_mark_mu);
  const long long end = ccal::g_mark_seq.load();
  const long long begin = end > 96 ? end - 96 : 0;
  int n = 0;
  for (long long q = begin; q < end && n < cap - 48; ++q) {
    const ccal::TraceMark& m = ccal::g_marks[q % ccal::kTraceMarks];
    if (m.ev == nullptr || m.seq != q) continue;
    const bool done = cudaEventQuery(m.ev) == cudaSuccess;
    n += snprintf(buf + n, (size_t)(cap - n), "%lld:s%llx:%d%s ", q, (m.stream >> 4) & 0xfff, m.id, done ? "" : "*PENDING*");
  }
  cudaGetLastError();
  buf[n < cap ? n : cap - 1] = 0;
  return n;
}

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
