// K3: per-bin {count, n_correct, sum conf} tables for ECE / MCE / ACE / PIECE, and the 16-bit
// radix histograms that give exact global order statistics (quantile bin edges).
//
// HBM-bound streaming kernels: 16 B per image (conf f32 + pred i32 + label i64), read once.
// Grid = a multiple of the SM count, grid-stride loop, warp-aggregated shared-memory atomics,
// one set of global atomics per CTA at the end.  Everything is integer / fixed point, so the
// result does not depend on the grid, the order, or how images are sharded over GPUs.
#include "ccal_common.cuh"

#include <math.h>
#include <math_constants.h>

namespace ccal {

constexpr int kBinThreads = 256;

// small by-value kernel parameter blocks: no device allocation, no extra copies
struct Thr64 { double t[CCAL_MAX_THRESHOLDS]; };
struct Thr32 { float t[CCAL_MAX_THRESHOLDS]; };
struct Prefixes { unsigned int p[64]; };

template <typename ConfT, typename PredT>
__global__ void __launch_bounds__(kBinThreads)
bin_stats_kernel(const ConfT* __restrict__ conf, const PredT* __restrict__ pred,
                 const long long* __restrict__ gt, long long n,
                 const __grid_constant__ Thr64 thr, int n_thr,
                 const float* __restrict__ key2, const __grid_constant__ Thr32 thr2, int n_thr2,
                 unsigned long long* __restrict__ table) {
  extern __shared__ unsigned char smem_raw[];
  const int n_cells = (n_thr + 1) * (n_thr2 + 1);
  BinCell* cells = reinterpret_cast<BinCell*>(smem_raw);
  double* s_thr = reinterpret_cast<double*>(cells + n_cells);
  float* s_thr2 = reinterpret_cast<float*>(s_thr + n_thr);
  for (int i = threadIdx.x; i < n_cells; i += blockDim.x) cells[i] = BinCell{0u, 0u, 0ull};
  for (int i = threadIdx.x; i < n_thr; i += blockDim.x) s_thr[i] = thr.t[i];
  for (int i = threadIdx.x; i < n_thr2; i += blockDim.x) s_thr2[i] = thr2.t[i];
  __syncthreads();

  const long long stride = (long long)gridDim.x * blockDim.x;
  // every lane of a warp runs the same number of iterations (warp_bin_add is warp-collective)
  const long long n_round = ((n + 31) / 32) * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    const bool valid = i < n;
    int bin = 0;
    bool correct = false;
    unsigned long long fx = 0;
    if (valid) {
      const ConfT x = conf[i];
      const double xd = (double)x;
      for (int j = 0; j < n_thr; ++j) bin += (xd >= s_thr[j]) ? 1 : 0;
      if (n_thr2 > 0) bin += bin_of(key2[i], s_thr2, n_thr2) * (n_thr + 1);
      correct = ((long long)pred[i] == gt[i]);
      fx = conf_to_fx(x);
    }
    warp_bin_add(cells, bin, correct, fx, valid);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_cells; i += blockDim.x) {
    const BinCell c = cells[i];
    if (c.count) {
      atomicAdd(&table[3 * i + 0], (unsigned long long)c.count);
      atomicAdd(&table[3 * i + 1], (unsigned long long)c.correct);
      atomicAdd(&table[3 * i + 2], c.sum_fx);
    }
  }
}


// Fast path (float confidences, 1-D table with <= kFastCells bins - every ECE/MCE/ACE call):
// every lane owns a private column of cells in shared memory, [warp][bin][lane], so the update
// is a plain 16-byte load / add / store: no atomics, no bank conflicts (a quarter warp of 128-bit
// accesses covers all 32 banks whatever the bins are), no inter-lane dependency chains.  Each
// thread streams 4 consecutive images per iteration with 128-bit loads.
constexpr int kFastCells = 16;
constexpr int kFastWarps = kBinThreads / 32;

// per lane and bin: {count (low 16 bits) | correct (high 16 bits)} and the 64-bit fixed-point confidence sum, in two
// arrays so that an update is one 4-byte and one 8-byte conflict-free read-modify-write.  A thread sees fewer than
// 65,536 images per launch (n < 2^32 over >= 592 x 256 threads), so the packed 16-bit counters cannot overflow.
struct LaneCounts { unsigned int count_correct; };

// kUniform: the thresholds are within half a bin of (i+1)/n_thr (every ECE/MCE table), so the bin is
// floor(x * n_thr) corrected by at most one step against the exact thresholds: 2 compares instead of n_thr.
template <typename PredT, bool kUniform>
__global__ void __launch_bounds__(kBinThreads)
bin_stats_fast_kernel(const float* __restrict__ conf, const PredT* __restrict__ pred,
                      const long long* __restrict__ gt, long long n,
                      const __grid_constant__ Thr32 thr, int n_thr, unsigned long long* __restrict__ table) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n_cells = n_thr + 1;
  unsigned long long* sums = reinterpret_cast<unsigned long long*>(smem_raw);          // [kFastWarps][n_cells][32]
  unsigned int* counts = reinterpret_cast<unsigned int*>(sums + kFastWarps * n_cells * 32);   // same shape
  __shared__ float s_thr[kFastCells];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kFastWarps * n_cells * 32; i += blockDim.x) { sums[i] = 0ull; counts[i] = 0u; }
  if (threadIdx.x < kFastCells) s_thr[threadIdx.x] = threadIdx.x < n_thr ? thr.t[threadIdx.x] : 0.f;
  __syncthreads();
  unsigned long long* my_sum = sums + (size_t)warp * n_cells * 32 + lane;
  unsigned int* my_cnt = counts + (size_t)warp * n_cells * 32 + lane;

  const float fn = (float)n_thr;
  auto add = [&](float x, long long p, long long g) {
    int b = 0;
    if (kUniform) {
      // guess g = floor(x * n) clamped to [0, n_thr]; exact answer = #(thr <= x) is g-1, g or g+1
      int gss = min(max(__float2int_rd(x * fn), 0), n_thr);
      const float lo = gss > 0 ? s_thr[gss - 1] : -CUDART_INF_F;       // bin gss = [thr[gss-1], thr[gss])
      const float hi = gss < n_thr ? s_thr[gss] : CUDART_INF_F;
      b = gss + (x >= hi ? 1 : 0) - (x < lo ? 1 : 0);
    } else {
      for (int j = 0; j < n_thr; ++j) b += (x >= s_thr[j]) ? 1 : 0;
    }
    my_cnt[b * 32] += (p == g) ? 0x10001u : 1u;
    my_sum[b * 32] += conf_to_fx(x);
  };

  const long long n4 = n / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(conf) | reinterpret_cast<uintptr_t>(pred) |
                         reinterpret_cast<uintptr_t>(gt)) & 15) == 0;
  long long done = 0;
  if (aligned) {
    struct Quad { float4 x; long long p[4]; longlong2 g0, g1; };
    auto load = [&](long long i) {
      Quad q;
      q.x = __ldcs(reinterpret_cast<const float4*>(conf) + i);
      if (sizeof(PredT) == 4) {
        const int4 v = __ldcs(reinterpret_cast<const int4*>(pred) + i);
        q.p[0] = v.x; q.p[1] = v.y; q.p[2] = v.z; q.p[3] = v.w;
      } else {
        const longlong2 v0 = __ldcs(reinterpret_cast<const longlong2*>(pred) + 2 * i);
        const longlong2 v1 = __ldcs(reinterpret_cast<const longlong2*>(pred) + 2 * i + 1);
        q.p[0] = v0.x; q.p[1] = v0.y; q.p[2] = v1.x; q.p[3] = v1.y;
      }
      q.g0 = __ldcs(reinterpret_cast<const longlong2*>(gt) + 2 * i);
      q.g1 = __ldcs(reinterpret_cast<const longlong2*>(gt) + 2 * i + 1);
      return q;
    };
    auto consume = [&](const Quad& q) {
      add(q.x.x, q.p[0], q.g0.x); add(q.x.y, q.p[1], q.g0.y); add(q.x.z, q.p[2], q.g1.x); add(q.x.w, q.p[3], q.g1.y);
    };
    // two independent 4-image groups per iteration: 128 B of loads in flight per thread
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n4; i += 2 * stride) {
      const Quad a = load(i), b = load(i + stride);
      consume(a);
      consume(b);
    }
    if (i < n4) consume(load(i));
    done = n4 * 4;
  }
  for (long long i = done + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    add(conf[i], (long long)pred[i], gt[i]);
  __syncthreads();
  // column sums: one warp per cell, lanes stride over the kFastWarps*32 private copies
  for (int cell = warp; cell < n_cells; cell += kFastWarps) {
    unsigned long long cnt = 0, cor = 0, sm = 0;
    for (int w = 0; w < kFastWarps; ++w) {
      const unsigned int cc = counts[((size_t)w * n_cells + cell) * 32 + lane];
      cnt += cc & 0xFFFFu; cor += cc >> 16;
      sm += sums[((size_t)w * n_cells + cell) * 32 + lane];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
      cor += __shfl_xor_sync(0xffffffffu, cor, off);
      sm += __shfl_xor_sync(0xffffffffu, sm, off);
    }
    if (lane == 0 && cnt) {
      atomicAdd(&table[3 * cell + 0], cnt);
      atomicAdd(&table[3 * cell + 1], cor);
      atomicAdd(&table[3 * cell + 2], sm);
    }
  }
}

// order-preserving key of a non-negative float = its bit pattern
// Keys of values in [2^-31, 2) - every confidence / proximity that matters - fall in the 4096-key
// window [0x3000, 0x4000): that window is privatised per CTA in shared memory (warp-aggregated
// shared atomics), everything else goes straight to global atomics.
constexpr unsigned kWinLo = 0x3000u, kWinSize = 0x1000u;

__global__ void __launch_bounds__(kBinThreads)
radix_hist_level0(const float* __restrict__ keys, long long n, unsigned int* __restrict__ hist) {
  __shared__ unsigned int win[kWinSize];
  for (int i = threadIdx.x; i < (int)kWinSize; i += blockDim.x) win[i] = 0u;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned key = __float_as_uint(__ldcs(keys + i)) >> 16;
    // confidences cluster (e.g. saturated 1.0): one atomic per distinct key per warp
    const unsigned peers = __match_any_sync(__activemask(), key);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) {
      const unsigned w = key - kWinLo;
      if (w < kWinSize) atomicAdd(&win[w], (unsigned)__popc(peers));
      else atomicAdd(&hist[key], (unsigned)__popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (int)kWinSize; i += blockDim.x)
    if (win[i]) atomicAdd(&hist[kWinLo + i], win[i]);
}

__global__ void __launch_bounds__(kBinThreads)
radix_hist_level1(const float* __restrict__ keys, long long n, const __grid_constant__ Prefixes prefixes,
                  int n_prefix, unsigned int* __restrict__ hist) {
  __shared__ unsigned int s_pref[64];
  for (int i = threadIdx.x; i < n_prefix; i += blockDim.x) s_pref[i] = prefixes.p[i];
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned bits = __float_as_uint(__ldcs(keys + i));
    const unsigned hi = bits >> 16;
    int slot = -1;
    for (int p = 0; p < n_prefix; ++p) slot = (s_pref[p] == hi) ? p : slot;
    // saturated confidences (thousands of identical keys) would serialise on one address: aggregate equal keys per warp
    const unsigned key = slot >= 0 ? ((unsigned)slot << 16) | (bits & 0xFFFFu) : 0xFFFFFFFFu;
    const unsigned peers = __match_any_sync(__activemask(), key);
    if (slot >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1)
      atomicAdd(&hist[(size_t)slot * 65536u + (bits & 0xFFFFu)], (unsigned)__popc(peers));
  }
}

static int grid_for(long long n, int threads, int per_sm) {
  long long want = (n + threads - 1) / threads;
  long long cap = (long long)num_sms() * per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// ----------------------------------------------------------------------------------------
// per-class confusion counts {tp, fp, fn} behind macro-F1 (evaluators/vl_evaluator.py:74-79)
// ----------------------------------------------------------------------------------------
constexpr int kClassSmemMax = 2048;        // classes privatised in shared memory (3 x u32 each)

template <typename PredT, bool kSmem>
__global__ void __launch_bounds__(kBinThreads)
class_counts_kernel(const PredT* __restrict__ pred, const long long* __restrict__ gt, long long n, int c,
                    unsigned long long* __restrict__ counts) {
  extern __shared__ unsigned int s_cnt[];   // [c][3] when kSmem
  if (kSmem) {
    for (int j = threadIdx.x; j < 3 * c; j += blockDim.x) s_cnt[j] = 0;
    __syncthreads();
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long p = (long long)pred[i], g = gt[i];
    const bool p_ok = p >= 0 && p < c, g_ok = g >= 0 && g < c;
    if (kSmem) {
      if (p == g) { if (g_ok) atomicAdd(&s_cnt[3 * g], 1u); }
      else {
        if (p_ok) atomicAdd(&s_cnt[3 * p + 1], 1u);
        if (g_ok) atomicAdd(&s_cnt[3 * g + 2], 1u);
      }
    } else {
      if (p == g) { if (g_ok) atomicAdd(&counts[3 * g], 1ull); }
      else {
        if (p_ok) atomicAdd(&counts[3 * p + 1], 1ull);
        if (g_ok) atomicAdd(&counts[3 * g + 2], 1ull);
      }
    }
  }
  if (kSmem) {
    __syncthreads();
    for (int j = threadIdx.x; j < 3 * c; j += blockDim.x)
      if (s_cnt[j]) atomicAdd(&counts[j], (unsigned long long)s_cnt[j]);
  }
}

}  // namespace ccal

using namespace ccal;

extern "C" int ccal_bin_stats(const void* conf, int conf_f64, const void* pred, int pred_i64,
                              const int64_t* gt, int64_t n, const double* thresholds_host, int n_thr,
                              const float* key2, const double* thresholds2_host, int n_thr2,
                              unsigned long long* table, ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE(n >= 0, "ccal_bin_stats: n < 0");
  CCAL_REQUIRE(n < (1ll << 32), "ccal_bin_stats: n must be < 2^32 per call (32-bit per-CTA counters)");
  CCAL_REQUIRE(n_thr >= 0 && n_thr <= CCAL_MAX_THRESHOLDS, "ccal_bin_stats: n_thr out of range");
  CCAL_REQUIRE(n_thr2 >= 0 && n_thr2 <= CCAL_MAX_THRESHOLDS, "ccal_bin_stats: n_thr2 out of range");
  CCAL_REQUIRE((n_thr + 1) * (n_thr2 + 1) <= 1024, "ccal_bin_stats: more than 1024 cells");
  CCAL_REQUIRE(table != nullptr, "ccal_bin_stats: table is NULL");
  CCAL_REQUIRE((n_thr2 == 0) == (key2 == nullptr), "ccal_bin_stats: key2 / n_thr2 mismatch");
  CCAL_REQUIRE(n_thr == 0 || thresholds_host, "ccal_bin_stats: thresholds NULL");
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(conf && pred && gt, "ccal_bin_stats: NULL input");
  CCAL_CUDA_OK(cudaPeekAtLastError());

  Thr64 t1;
  Thr32 t2;
  for (int i = 0; i < CCAL_MAX_THRESHOLDS; ++i) {
    t1.t[i] = i < n_thr ? thresholds_host[i] : 0.0;
    t2.t[i] = i < n_thr2 ? ceil_to_f32(thresholds2_host[i]) : 0.0f;
  }
  const int n_cells = (n_thr + 1) * (n_thr2 + 1);
  const long long* g = reinterpret_cast<const long long*>(gt);
  if (!conf_f64 && n_thr2 == 0 && n_cells <= kFastCells) {
    // float keys: compare against ceil_f32(threshold) (equivalent to the double comparison)
    Thr32 tf;
    for (int i = 0; i < CCAL_MAX_THRESHOLDS; ++i) tf.t[i] = i < n_thr ? ceil_to_f32(thresholds_host[i]) : 0.0f;
    const size_t fsmem = (size_t)12 * kFastWarps * n_cells * 32;
    // >= 592 CTAs whenever there is that much work keeps every thread below 65,536 images (packed 16-bit counters)
    const int fgrid = grid_for((n + 3) / 4, kBinThreads, 4);
    CCAL_REQUIRE((n + (long long)fgrid * kBinThreads - 1) / ((long long)fgrid * kBinThreads) < 65536,
                 "ccal_bin_stats: too many images per thread for the packed counters (n=%lld on %d CTAs)", (long long)n, fgrid);
    bool uniform = n_thr >= 1;
    for (int i = 0; i < n_thr; ++i)
      uniform = uniform && fabs(thresholds_host[i] - (double)(i + 1) / n_thr) < 0.25 / n_thr;
#define CCAL_LAUNCH_FAST(PT, U)                                                                                        \
  do {                                                                                                                 \
    CCAL_CUDA_OK(cudaFuncSetAttribute(bin_stats_fast_kernel<PT, U>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                      (int)fsmem));                                                                    \
    bin_stats_fast_kernel<PT, U><<<fgrid, kBinThreads, fsmem, stream>>>((const float*)conf, (const PT*)pred, g,        \
                                                                        (long long)n, tf, n_thr, table);               \
  } while (0)
    if (pred_i64) { if (uniform) CCAL_LAUNCH_FAST(long long, true); else CCAL_LAUNCH_FAST(long long, false); }
    else { if (uniform) CCAL_LAUNCH_FAST(int, true); else CCAL_LAUNCH_FAST(int, false); }
#undef CCAL_LAUNCH_FAST
    note_launch();
    CCAL_CUDA_OK(cudaGetLastError());
    return CCAL_OK;
  }
  const size_t smem = sizeof(BinCell) * n_cells + sizeof(double) * n_thr + sizeof(float) * n_thr2;
  const int grid = grid_for(n, kBinThreads, 8);
#define CCAL_LAUNCH_BIN(CT, PT)                                                                   \
  bin_stats_kernel<CT, PT><<<grid, kBinThreads, smem, stream>>>(                                  \
      (const CT*)conf, (const PT*)pred, g, (long long)n, t1, n_thr, key2, t2, n_thr2, table)
  if (conf_f64) {
    if (pred_i64) CCAL_LAUNCH_BIN(double, long long); else CCAL_LAUNCH_BIN(double, int);
  } else {
    if (pred_i64) CCAL_LAUNCH_BIN(float, long long); else CCAL_LAUNCH_BIN(float, int);
  }
  note_launch();
#undef CCAL_LAUNCH_BIN
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

extern "C" int ccal_radix_hist(const float* keys, int64_t n, int level, const uint32_t* prefixes_host,
                               int n_prefix, uint32_t* hist, ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE(n >= 0 && n < (1ll << 32), "ccal_radix_hist: n out of range");
  CCAL_REQUIRE(level == 0 || level == 1, "ccal_radix_hist: level must be 0 or 1");
  CCAL_REQUIRE(hist != nullptr, "ccal_radix_hist: hist is NULL");
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(keys != nullptr, "ccal_radix_hist: keys is NULL");
  const int grid = grid_for(n, kBinThreads, 8);
  if (level == 0) {
    radix_hist_level0<<<grid, kBinThreads, 0, stream>>>(keys, (long long)n, hist);
    note_launch();
  } else {
    CCAL_REQUIRE(n_prefix >= 1 && n_prefix <= 64 && prefixes_host, "ccal_radix_hist: 1..64 prefixes required");
    Prefixes pf;
    for (int i = 0; i < 64; ++i) pf.p[i] = i < n_prefix ? prefixes_host[i] : 0xFFFFFFFFu;
    radix_hist_level1<<<grid, kBinThreads, 0, stream>>>(keys, (long long)n, pf, n_prefix, hist);
    note_launch();
  }
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

extern "C" int ccal_class_counts(const void* pred, int pred_i64, const int64_t* gt, int64_t n, int c,
                                 unsigned long long* counts, ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE(n >= 0 && n < (1ll << 32), "ccal_class_counts: n out of range");
  CCAL_REQUIRE(c >= 1, "ccal_class_counts: c must be positive (got %d)", c);
  CCAL_REQUIRE(counts != nullptr, "ccal_class_counts: counts is NULL");
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(pred && gt, "ccal_class_counts: NULL input");
  const long long* g = reinterpret_cast<const long long*>(gt);
  const int grid = grid_for(n, kBinThreads, 8);
  const size_t smem = (size_t)3 * c * sizeof(unsigned int);
  if (c <= kClassSmemMax) {
    if (pred_i64) class_counts_kernel<long long, true><<<grid, kBinThreads, smem, stream>>>((const long long*)pred, g, n, c, counts);
    else class_counts_kernel<int, true><<<grid, kBinThreads, smem, stream>>>((const int*)pred, g, n, c, counts);
  } else {
    if (pred_i64) class_counts_kernel<long long, false><<<grid, kBinThreads, 0, stream>>>((const long long*)pred, g, n, c, counts);
    else class_counts_kernel<int, false><<<grid, kBinThreads, 0, stream>>>((const int*)pred, g, n, c, counts);
  }
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}
