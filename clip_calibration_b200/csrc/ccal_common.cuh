// Shared host/device helpers for libccal (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/ccal.h"

namespace ccal {

// ----------------------------------------------------------------------------------------
// host: thread-local error string
// ----------------------------------------------------------------------------------------
char* error_buffer();
int fail(int code, const char* fmt, ...);

#define CCAL_CUDA_OK(expr)                                                             \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess)                                                             \
      return ::ccal::fail(CCAL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                 \
  } while (0)

#define CCAL_REQUIRE(cond, ...)                                                        \
  do {                                                                                 \
    if (!(cond)) return ::ccal::fail(CCAL_ERR_BAD_ARG, __VA_ARGS__);                   \
  } while (0)

int num_sms();
void note_launch(int n = 1);   // process-wide count of kernels this library launched
// Development aid (CCAL_TRACE_MARKS=1): record an event named `id` on `stream`; ccal_trace_marks_report() lists which
// of the most recent marks the device has reached - the way to see WHERE a stream stopped making progress.
void trace_mark(cudaStream_t stream, int id);

// TMA descriptor of a [rows, d] row-major 16-bit matrix (dtype CCAL_BF16 / CCAL_F16), box =
// {64 features, box_rows}, 128-byte swizzle, zero fill out of range (defined in score_fused.cu).
int make_map(CUtensorMap* map, const void* base, long long rows, int d, int box_rows, int dtype);

// Transient, stream-ordered workspace (cudaMallocAsync); released on every exit path.
struct AsyncWorkspace {
  unsigned char* ptr = nullptr;
  cudaStream_t stream = nullptr;
  cudaError_t alloc(size_t bytes, cudaStream_t s) {
    stream = s;
    // keep freed workspace cached in the device's default pool instead of returning it to the driver at
    // every synchronisation (the default release threshold is 0, which costs milliseconds per call)
    static thread_local int tuned_dev = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev != tuned_dev) {
      cudaMemPool_t pool;
      if ((e = cudaDeviceGetDefaultMemPool(&pool, dev)) != cudaSuccess) return e;
      unsigned long long keep = ~0ull;
      if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) return e;
      tuned_dev = dev;
    }
    return cudaMallocAsync(reinterpret_cast<void**>(&ptr), bytes, s);
  }
  ~AsyncWorkspace() {
    if (ptr) cudaFreeAsync(ptr, stream);
  }
  AsyncWorkspace() = default;
  AsyncWorkspace(const AsyncWorkspace&) = delete;
  AsyncWorkspace& operator=(const AsyncWorkspace&) = delete;
};

// Thresholds are given as doubles (np.linspace edges are float64 and np.digitize compares
// in double).  For a float key x and a double threshold t:  x >= t  <=>  x >= ceil_f32(t),
// the smallest float32 that is >= t.  So the device compares floats only.
float ceil_to_f32(double t);

// ----------------------------------------------------------------------------------------
// device: bin table accumulation with warp-aggregated shared-memory atomics
// ----------------------------------------------------------------------------------------
struct BinCell {                 // shared-memory cell
  unsigned int count;
  unsigned int correct;
  unsigned long long sum_fx;     // sum of round(conf * 2^40)
};

__device__ __forceinline__ unsigned long long conf_to_fx(float conf) {
  // conf in [0,1]; for conf >= 2^-16 (always true for softmax maxima over <= 65536 classes)
  // conf * 2^40 is an exact integer, so the sum is order independent AND exact.
  return (unsigned long long)__float2ull_rn(conf * 1099511627776.0f);
}
__device__ __forceinline__ unsigned long long conf_to_fx(double conf) {
  return (unsigned long long)__double2ull_rn(conf * 1099511627776.0);
}

// All 32 lanes must call.  Lanes with valid == false contribute nothing.  One leader lane per
// distinct bin present in the warp issues the three shared-memory atomics.
__device__ __forceinline__ void warp_bin_add(BinCell* cells, int bin, bool correct,
                                             unsigned long long fx, bool valid) {
  const unsigned full = 0xffffffffu;
  unsigned todo = __ballot_sync(full, valid);
  const int lane = threadIdx.x & 31;
  while (todo) {
    const int leader = __ffs(todo) - 1;
    const int b = __shfl_sync(full, bin, leader);
    const bool mine = valid && (bin == b);
    const unsigned grp = __ballot_sync(full, mine);
    const unsigned ncorrect = __popc(__ballot_sync(full, mine && correct));
    // 41-bit fixed point split in two halves so the 32-bit warp reduction cannot overflow
    unsigned lo = mine ? (unsigned)(fx & 0xFFFFFu) : 0u;
    unsigned hi = mine ? (unsigned)(fx >> 20) : 0u;
    lo = __reduce_add_sync(full, lo);
    hi = __reduce_add_sync(full, hi);
    if (lane == leader) {
      atomicAdd(&cells[b].count, (unsigned)__popc(grp));
      atomicAdd(&cells[b].correct, ncorrect);
      atomicAdd(&cells[b].sum_fx, ((unsigned long long)hi << 20) + lo);
    }
    todo &= ~grp;
  }
}

__device__ __forceinline__ int bin_of(float x, const float* thr, int n_thr) {
  int b = 0;
  for (int j = 0; j < n_thr; ++j) b += (x >= thr[j]) ? 1 : 0;
  return b;
}

}  // namespace ccal
