// f-4: multi-class isotonic-regression calibrator, reference trainers/calibration/multi_isotonic_regression.py
// (`IsotonicRegression(out_of_bounds='clip').fit_transform(p.flatten(), onehot.flatten())`, scikit-learn) and its
// proximity-binned wrapper BinMeanShift (multi_proximity_isotonic.py:130-247).
//
//   exp_normalise_rows   p = exp(v) / sum(exp(v)) per row in float64, no max shift - the reference's own formula
//                        (multi_isotonic_regression.py:26, :33) - and the flattened one-hot targets.
//   isotonic_fit_binary  scikit-learn's `_build_y` for 0/1 targets: sort by x (radix sort, sort_scan.cuh), merge equal x
//                        (`_make_unique`: a new value starts where x - first x of the current value >= 1e-15), pool adjacent violators,
//                        drop interior points of constant stretches -> knots (X_thresholds_, y_thresholds_).
//   isotonic_transform   clip to [X_min_, X_max_], linear interpolation between knots (scipy interp1d), plus the
//                        reference's `+ 1e-9 * p` term.  HBM-bound, 16 B per element.
//
// Pool-adjacent-violators is sequential as written in scikit-learn (_isotonic.pyx); here it runs in ROUNDS: every
// maximal run of blocks whose means do not strictly increase is pooled into one block, until no run is left.  The
// fixed point is the same isotonic fit (pooling a non-increasing run is forced); ~10 rounds for 2M points.  Targets
// are 0/1, so a block is the integer pair (ones, count): comparisons are exact cross-multiplications and the fitted
// value is one correctly rounded division - there is no summation order to worry about.
#include "ccal_common.cuh"
#include "sort_scan.cuh"

#include <algorithm>
#include <math_constants.h>

namespace ccal {

constexpr int kIsoThreads = 256;
constexpr double kUniqueEps = 1e-15;        // np.finfo(np.float64).resolution, sklearn _make_unique

static inline unsigned iso_grid(long long n) {
  return (unsigned)std::min<long long>((n + kIsoThreads - 1) / kIsoThreads, 1 << 20);
}

// ---------------------------------------------------------------------------------------- rows
template <typename T, int GROUP>
__global__ void __launch_bounds__(256)
exp_normalise_rows_kernel(const T* __restrict__ v, long long n, int c, double* __restrict__ out,
                          const long long* __restrict__ labels, unsigned char* __restrict__ onehot) {
  __shared__ double s_sum[8];
  constexpr int kRowsPerCta = 256 / GROUP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = threadIdx.x % GROUP;
  const long long stride = (long long)gridDim.x * kRowsPerCta;
  const long long sweeps = (n + stride - 1) / stride;
  for (long long sweep = 0; sweep < sweeps; ++sweep) {
    const long long row = (long long)blockIdx.x * kRowsPerCta + threadIdx.x / GROUP + sweep * stride;
    const bool live = row < n;
    const T* x = v + (live ? row : 0) * (long long)c;
    double sum = 0.0;
    double* o = out + (live ? row : 0) * (long long)c;
    if (live)
      for (int j = t; j < c; j += GROUP) {              // exp once: park it in the output row, rescale below
        const double e = exp((double)x[j]);
        o[j] = e;
        sum += e;
      }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (GROUP > 32) {
      if (lane == 0) s_sum[warp] = sum;
      __syncthreads();
      sum = 0.0;
      for (int w = 0; w < 8; ++w) sum += s_sum[w];
      __syncthreads();
    }
    if (!live) continue;
    const long long lab = labels ? labels[row] : -1;
    for (int j = t; j < c; j += GROUP) {                // every thread re-reads only what it wrote itself
      o[j] = o[j] / sum;
      if (onehot) onehot[row * (long long)c + j] = (unsigned char)(lab == j);
    }
  }
}

// ---------------------------------------------------------------------------------------- fit
__global__ void iso_unique_flags_kernel(const double* __restrict__ xs, long long n, int* __restrict__ flag) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    flag[i] = (i == 0 || xs[i] - xs[i - 1] >= kUniqueEps) ? 1 : 0;
}

// scikit-learn's `_make_unique` is ANCHORED: a new value starts where x - x_first_of_current_group >= eps, not where
// two neighbours differ by eps.  Every neighbour-rule flag above is also an anchored start (the anchor is never
// above the left neighbour), but a run of neighbours closer than eps (dense softmax tails) may hold further starts:
// next[i] = first j > i with xs[j] - xs[i] >= eps (binary search; the difference is monotone in j), and the starts
// are the orbit of the flagged positions under `next`.  The orbit is marked by pointer doubling: round r marks
// next^(2^r) of everything marked so far and squares the jump table, so a run holding g starts needs log2(g) rounds;
// data without such runs needs one round that marks nothing.
__global__ void iso_next_kernel(const double* __restrict__ xs, long long n, const int* __restrict__ flag,
                                int* __restrict__ next) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long long j = n;
    if (i + 1 < n) {
      if (flag[i + 1]) {
        j = i + 1;
      } else {
        const double a = xs[i];
        long long lo = i + 1, hi = n;                     // first j in (i, n) with xs[j] - a >= eps, else n
        while (lo < hi) {
          const long long mid = (lo + hi) >> 1;
          if (xs[mid] - a >= kUniqueEps) hi = mid; else lo = mid + 1;
        }
        j = lo;
      }
    }
    next[i] = (int)j;
  }
}

// flag_out is a copy of flag_in on entry; only positions that become marked are written
__global__ void iso_mark_round_kernel(const int* __restrict__ flag_in, const int* __restrict__ next, long long n,
                                      int* __restrict__ flag_out, int* __restrict__ next_out, int* __restrict__ changed) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int j = next[i];
    if (flag_in[i] && j < n && !flag_in[j]) { flag_out[j] = 1; *changed = 1; }
    next_out[i] = j < n ? next[j] : (int)n;
  }
}

// gid = inclusive scan of the flags (1-based group number)
__global__ void iso_group_kernel(const double* __restrict__ xs, const unsigned char* __restrict__ ys,
                                 const int* __restrict__ flag, const int* __restrict__ gid, long long n,
                                 double* __restrict__ ux, unsigned long long* __restrict__ ones,
                                 unsigned long long* __restrict__ cnt, int* __restrict__ start) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int g = gid[i] - 1;
    if (flag[i]) { ux[g] = xs[i]; start[g] = g; }
    if (ys[i]) atomicAdd(&ones[g], 1ull);
    atomicAdd(&cnt[g], 1ull);
  }
}

// head[j] = block j starts a new pooled block: j == 0 or mean(j-1) < mean(j) strictly
__global__ void iso_heads_kernel(const unsigned long long* __restrict__ ones, const unsigned long long* __restrict__ cnt,
                                 int nb, int* __restrict__ head, int* __restrict__ any_violation) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nb) return;
  int h = 1;
  if (j > 0) {
    // mean(j-1) >= mean(j)  <=>  ones[j-1] * cnt[j] >= ones[j] * cnt[j-1]   (exact: both factors < 2^31)
    const bool violation = ones[j - 1] * cnt[j] >= ones[j] * cnt[j - 1];
    if (violation) { h = 0; *any_violation = 1; }
  }
  head[j] = h;
}

__global__ void iso_pool_kernel(const unsigned long long* __restrict__ ones, const unsigned long long* __restrict__ cnt,
                                const int* __restrict__ start, const int* __restrict__ head, const int* __restrict__ bid,
                                int nb, unsigned long long* __restrict__ ones_out, unsigned long long* __restrict__ cnt_out,
                                int* __restrict__ start_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nb) return;
  const int b = bid[j] - 1;
  if (head[j]) start_out[b] = start[j];
  atomicAdd(&ones_out[b], ones[j]);
  atomicAdd(&cnt_out[b], cnt[j]);
}

// fitted value of every unique x: its block = the last one whose first group is <= g
__global__ void iso_expand_kernel(const unsigned long long* __restrict__ ones, const unsigned long long* __restrict__ cnt,
                                  const int* __restrict__ start, int nb, int n_groups, double* __restrict__ fy) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  int lo = 0, hi = nb - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (start[mid] <= g) lo = mid; else hi = mid - 1;
  }
  fy[g] = (double)ones[lo] / (double)cnt[lo];
}

// sklearn _build_y, trim_duplicates: drop interior points equal to both neighbours
__global__ void iso_keep_kernel(const double* __restrict__ fy, int n_groups, unsigned char* __restrict__ keep) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  keep[g] = (g == 0 || g == n_groups - 1 || fy[g] != fy[g - 1] || fy[g] != fy[g + 1]) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------- transform
constexpr int kIsoSmemKnots = 3072;           // knots staged in shared memory (48 KB); more stay in global memory / L1

// scikit-learn IsotonicRegression.transform: out_of_bounds='clip', then scipy interp1d(kind='linear') on the knots
__device__ __forceinline__ double iso_eval(const double* __restrict__ kx, const double* __restrict__ ky, int nk, double v) {
  if (nk == 1) return ky[0];
  const double x = fmin(fmax(v, kx[0]), kx[nk - 1]);
  int lo = 0, hi = nk;                                         // searchsorted(kx, x, side='left')
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (kx[mid] < x) lo = mid + 1; else hi = mid;
  }
  const int idx = min(max(lo, 1), nk - 1);
  const double x0 = kx[idx - 1], x1 = kx[idx], y0 = ky[idx - 1], y1 = ky[idx];
  const double slope = (y1 - y0) / (x1 - x0);
  return slope * (x - x0) + y0;
}

template <bool kSmem>
__global__ void __launch_bounds__(kIsoThreads)
iso_transform_kernel(const double* __restrict__ gx, const double* __restrict__ gy, int nk, const double* __restrict__ t,
                     long long n, double residual_scale, double* __restrict__ out) {
  extern __shared__ double s_knots[];            // [nk] x then [nk] y
  const double* kx = gx;
  const double* ky = gy;
  if (kSmem) {
    for (int j = threadIdx.x; j < nk; j += kIsoThreads) { s_knots[j] = gx[j]; s_knots[nk + j] = gy[j]; }
    __syncthreads();
    kx = s_knots;
    ky = s_knots + nk;
  }
  auto f = [&](double v) { return iso_eval(kx, ky, nk, v) + residual_scale * v; };
  // four independent loads in flight per thread
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (i + u * stride < n) ? t[i + u * stride] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < n) out[i + u * stride] = f(v[u]);
  }
}

struct HostPair { int any_violation; int last_bid; };

// ---------------------------------------------------------------------------------------- one-vs-all calibrators
// netcal's multi-class scheme (AbstractCalibration._create_one_vs_all_models / _calibrate_multiclass): class j gets its
// own BINARY calibrator fitted on (X[:, j], y == j); transform applies calibrator j to column j and divides every row
// by its sum.  The matrix is row-major, so one warp walks a row with consecutive lanes on consecutive classes.
constexpr int kOvaMaxBins = 64;

__device__ __forceinline__ int ova_bin(double x, const double* edges, int n_bins) {
  // np.linspace edges; bin i <=> edges[i] <= x < edges[i+1], the last bin closed on the right (scipy
  // binned_statistic_dd / np.digitize - 1 clipped), values below edges[0] fall into bin 0
  int b = 0;
  for (int j = 1; j < n_bins; ++j) b += (x >= edges[j]) ? 1 : 0;
  return b;
}

// counts / hits [c_tile][n_bins] live in shared memory for a tile of classes; grid = (row groups, class tiles)
constexpr int kOvaFitThreads = 1024;           // 32 warps = 32 rows in flight per CTA, four loads in flight per lane

template <typename T>
__global__ void __launch_bounds__(kOvaFitThreads)
ova_hist_fit_kernel(const T* __restrict__ p, long long n, int c, const long long* __restrict__ labels,
                    const double* __restrict__ edges_g, int n_bins, int tile_c, unsigned* __restrict__ count,
                    unsigned* __restrict__ hits) {
  extern __shared__ unsigned s_ova[];               // [tile_c * n_bins] counts, then [tile_c * n_bins] hits
  __shared__ double s_edges[kOvaMaxBins + 1];
  constexpr int kWarps = kOvaFitThreads / 32;
  const int c0 = blockIdx.y * tile_c;
  const int cw = min(tile_c, c - c0);
  const int cells = cw * n_bins;
  for (int j = threadIdx.x; j < 2 * cells; j += kOvaFitThreads) s_ova[j] = 0;
  if (threadIdx.x <= n_bins) s_edges[threadIdx.x] = edges_g[threadIdx.x];
  __syncthreads();
  unsigned* s_cnt = s_ova;
  unsigned* s_hit = s_ova + cells;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long row = (long long)blockIdx.x * kWarps + warp; row < n; row += (long long)gridDim.x * kWarps) {
    const T* x = p + row * (long long)c + c0;
    const long long lab = labels[row] - c0;
    // the walk over a row is a chain of dependent-latency loads unless several are issued before the first is used
    for (int j0 = lane; j0 < cw; j0 += 128) {
      T v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (j0 + 32 * u < cw) ? x[j0 + 32 * u] : T(0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 32 * u;
        if (j < cw) {
          const int b = ova_bin((double)v[u], s_edges, n_bins);
          // softmax tails put nearly every (row, class) pair into bin 0: that bin is not counted here but derived as
          // n - (all other bins) by ova_hist_finish_kernel
          if (b > 0) atomicAdd(&s_cnt[j * n_bins + b], 1u);
          if (j == lab) atomicAdd(&s_hit[j * n_bins + b], 1u);
        }
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < cells; j += kOvaFitThreads) {
    if (s_cnt[j]) atomicAdd(&count[(long long)c0 * n_bins + j], s_cnt[j]);
    if (s_hit[j]) atomicAdd(&hits[(long long)c0 * n_bins + j], s_hit[j]);
  }
}

// count[j][0] = n - sum_{b >= 1} count[j][b]: every row falls into exactly one bin of every class
__global__ void ova_hist_finish_kernel(unsigned* __restrict__ count, int c, int n_bins, unsigned n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c) return;
  unsigned rest = 0;
  for (int b = 1; b < n_bins; ++b) rest += count[(long long)j * n_bins + b];
  count[(long long)j * n_bins] = n - rest;
}

// kMode 0: histogram binning (out = bin_map[class][bin]); 1: isotonic (out = f_class(x), knots of class j are
// knots[off[j] .. off[j+1]), a class without knots gives 0).  One warp per row; the row sum is accumulated in class
// order per lane and combined by a fixed butterfly, so the result does not depend on the launch shape.
template <typename T, int kMode>
__global__ void __launch_bounds__(256)
ova_apply_kernel(const T* __restrict__ p, long long n, int c, const double* __restrict__ edges_g, int n_bins,
                 const double* __restrict__ bin_map, const double* __restrict__ kx, const double* __restrict__ ky,
                 const int* __restrict__ off, int normalise, double* __restrict__ out) {
  __shared__ double s_edges[kOvaMaxBins + 1];
  if (kMode == 0 && threadIdx.x <= n_bins) s_edges[threadIdx.x] = edges_g[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < n; row += (long long)gridDim.x * 8) {
    const T* x = p + row * (long long)c;
    double* o = out + row * (long long)c;
    double sum = 0.0;
    for (int j = lane; j < c; j += 32) {
      const double v = (double)x[j];
      double r;
      if (kMode == 0) {
        r = bin_map[(long long)j * n_bins + ova_bin(v, s_edges, n_bins)];
      } else {
        const int a = off[j], nk = off[j + 1] - a;
        r = nk > 0 ? iso_eval(kx + a, ky + a, nk, v) : 0.0;
      }
      o[j] = r;
      sum += r;
    }
    if (!normalise) continue;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    for (int j = lane; j < c; j += 32) o[j] = o[j] / sum;      // every lane re-reads only what it wrote itself
  }
}

}  // namespace ccal

using namespace ccal;

extern "C" int ccal_exp_normalise_rows(const float* v_f32, const double* v_f64, int64_t n, int c, double* out,
                                       const int64_t* labels, unsigned char* onehot_out, ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE((v_f32 != nullptr) != (v_f64 != nullptr), "ccal_exp_normalise_rows: exactly one of v_f32 / v_f64 must be given");
  CCAL_REQUIRE(n >= 0 && c >= 1, "ccal_exp_normalise_rows: bad shape n=%lld c=%d", (long long)n, c);
  CCAL_REQUIRE(out != nullptr, "ccal_exp_normalise_rows: out is NULL");
  CCAL_REQUIRE((labels == nullptr) == (onehot_out == nullptr), "ccal_exp_normalise_rows: labels and onehot_out go together");
  if (n == 0) return CCAL_OK;
  const long long cap = (long long)num_sms() * 8;
  const long long* lab = reinterpret_cast<const long long*>(labels);
#define CCAL_LAUNCH_EXPN(T, ptr)                                                                                        \
  do {                                                                                                                  \
    if (c <= 2048)                                                                                                      \
      exp_normalise_rows_kernel<T, 32><<<(int)std::min<long long>((n + 7) / 8, cap), 256, 0, stream>>>(ptr, n, c, out, lab, onehot_out); \
    else                                                                                                                \
      exp_normalise_rows_kernel<T, 256><<<(int)std::min<long long>(n, cap), 256, 0, stream>>>(ptr, n, c, out, lab, onehot_out);          \
  } while (0)
  if (v_f32) CCAL_LAUNCH_EXPN(float, v_f32); else CCAL_LAUNCH_EXPN(double, v_f64);
#undef CCAL_LAUNCH_EXPN
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

extern "C" int ccal_isotonic_fit_binary(const double* x, const unsigned char* y, int64_t n, double* knots_x,
                                        double* knots_y, int64_t* n_knots_host, ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE(n >= 1 && n < 2147483647ll, "ccal_isotonic_fit_binary: n must be in [1, 2^31) (got %lld)", (long long)n);
  CCAL_REQUIRE(x && y && knots_x && knots_y && n_knots_host, "ccal_isotonic_fit_binary: NULL pointer");

  // ---- workspace layout
  const SortPlan plan = sort_plan(n);
  const size_t scan_bytes = scan_workspace_bytes(n);
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t sz_d = up(sizeof(double) * n), sz_u8 = up(n), sz_i = up(sizeof(int) * n), sz_u64 = up(8 * (size_t)n);
  // xs, ux, fy | ys, keep | flag, gid, start0, start1 | ones0, cnt0, ones1, cnt1 | small | prefix-sum tile sums | sort scratch
  const size_t total = 3 * sz_d + 2 * sz_u8 + 4 * sz_i + 4 * sz_u64 + 256 + up(scan_bytes) + up(plan.total);
  AsyncWorkspace ws;
  CCAL_CUDA_OK(ws.alloc(total, stream));
  unsigned char* p = ws.ptr;
  auto take = [&](size_t b) { unsigned char* r = p; p += b; return r; };
  double* xs = (double*)take(sz_d); double* ux = (double*)take(sz_d); double* fy = (double*)take(sz_d);
  unsigned char* ys = take(sz_u8); unsigned char* keep = take(sz_u8);
  int* flag = (int*)take(sz_i); int* gid = (int*)take(sz_i);
  int* start[2] = {(int*)take(sz_i), (int*)take(sz_i)};
  unsigned long long* ones[2]; unsigned long long* cnt[2];
  ones[0] = (unsigned long long*)take(sz_u64); cnt[0] = (unsigned long long*)take(sz_u64);
  ones[1] = (unsigned long long*)take(sz_u64); cnt[1] = (unsigned long long*)take(sz_u64);
  int* small = (int*)take(256);                 // [0] any_violation
  void* scan_ws = take(up(scan_bytes));
  unsigned char* sort_ws = take(up(plan.total));

  // ---- sort by x, merge equal x into groups (ones, count)
  CCAL_CUDA_OK(sort_pairs_f64_u8(x, y, n, xs, ys, sort_ws, stream));
  iso_unique_flags_kernel<<<iso_grid(n), kIsoThreads, 0, stream>>>(xs, n, flag);
  note_launch();
  // anchored starts inside runs of close neighbours (buffers: gid = second flag array, start[] = jump tables)
  iso_next_kernel<<<iso_grid(n), kIsoThreads, 0, stream>>>(xs, n, flag, start[0]);
  note_launch();
  {
    int* f_cur = flag; int* f_oth = gid;
    int cur_next = 0;
    for (int round = 0; round < 40; ++round) {
      CCAL_CUDA_OK(cudaMemcpyAsync(f_oth, f_cur, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
      CCAL_CUDA_OK(cudaMemsetAsync(small, 0, sizeof(int), stream));
      iso_mark_round_kernel<<<iso_grid(n), kIsoThreads, 0, stream>>>(f_cur, start[cur_next], n, f_oth, start[cur_next ^ 1], small);
      note_launch();
      int changed = 0;
      CCAL_CUDA_OK(cudaMemcpyAsync(&changed, small, sizeof(int), cudaMemcpyDeviceToHost, stream));
      CCAL_CUDA_OK(cudaStreamSynchronize(stream));
      if (!changed) break;                         // f_oth == f_cur: the orbit is closed
      std::swap(f_cur, f_oth);
      cur_next ^= 1;
    }
    if (f_cur != flag) CCAL_CUDA_OK(cudaMemcpyAsync(flag, f_cur, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
  }
  CCAL_CUDA_OK(prefix_sum_i32(flag, gid, n, true, scan_ws, stream));
  CCAL_CUDA_OK(cudaMemsetAsync(ones[0], 0, 8 * (size_t)n, stream));
  CCAL_CUDA_OK(cudaMemsetAsync(cnt[0], 0, 8 * (size_t)n, stream));
  iso_group_kernel<<<iso_grid(n), kIsoThreads, 0, stream>>>(xs, ys, flag, gid, n, ux, ones[0], cnt[0], start[0]);
  note_launch();
  int n_groups = 0;
  CCAL_CUDA_OK(cudaMemcpyAsync(&n_groups, gid + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
  CCAL_CUDA_OK(cudaStreamSynchronize(stream));

  // ---- pool adjacent violators in rounds
  int nb = n_groups, cur = 0;
  int* head = flag;            // reused
  int* bid = gid;
  for (int round = 0; nb > 1; ++round) {
    CCAL_REQUIRE(round < 100000, "ccal_isotonic_fit_binary: pooling did not converge");
    CCAL_CUDA_OK(cudaMemsetAsync(small, 0, sizeof(int), stream));
    iso_heads_kernel<<<(nb + kIsoThreads - 1) / kIsoThreads, kIsoThreads, 0, stream>>>(ones[cur], cnt[cur], nb, head, small);
    note_launch();
    CCAL_CUDA_OK(prefix_sum_i32(head, bid, nb, true, scan_ws, stream));
    HostPair hp;
    CCAL_CUDA_OK(cudaMemcpyAsync(&hp.any_violation, small, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CCAL_CUDA_OK(cudaMemcpyAsync(&hp.last_bid, bid + (nb - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    CCAL_CUDA_OK(cudaStreamSynchronize(stream));
    if (!hp.any_violation) break;
    const int nxt = cur ^ 1;
    CCAL_CUDA_OK(cudaMemsetAsync(ones[nxt], 0, 8 * (size_t)hp.last_bid, stream));
    CCAL_CUDA_OK(cudaMemsetAsync(cnt[nxt], 0, 8 * (size_t)hp.last_bid, stream));
    iso_pool_kernel<<<(nb + kIsoThreads - 1) / kIsoThreads, kIsoThreads, 0, stream>>>(ones[cur], cnt[cur], start[cur], head, bid, nb,
                                                                                    ones[nxt], cnt[nxt], start[nxt]);
    note_launch();
    nb = hp.last_bid;
    cur = nxt;
  }

  // ---- fitted value per unique x, then drop interior points of constant stretches
  iso_expand_kernel<<<(n_groups + kIsoThreads - 1) / kIsoThreads, kIsoThreads, 0, stream>>>(ones[cur], cnt[cur], start[cur], nb,
                                                                                           n_groups, fy);
  note_launch();
  iso_keep_kernel<<<(n_groups + kIsoThreads - 1) / kIsoThreads, kIsoThreads, 0, stream>>>(fy, n_groups, keep);
  note_launch();
  // compaction: slot of a kept point = number of kept points up to and including it (flag / gid are free again)
  const unsigned kb = (n_groups + kIsoThreads - 1) / kIsoThreads;
  compact_flags_kernel<<<kb, kIsoThreads, 0, stream>>>(keep, n_groups, flag);
  note_launch();
  CCAL_CUDA_OK(prefix_sum_i32(flag, gid, n_groups, true, scan_ws, stream));
  compact_pair_kernel<<<kb, kIsoThreads, 0, stream>>>(ux, fy, keep, gid, n_groups, knots_x, knots_y);
  note_launch();
  int n_knots = 0;
  CCAL_CUDA_OK(cudaMemcpyAsync(&n_knots, gid + (n_groups - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
  CCAL_CUDA_OK(cudaStreamSynchronize(stream));
  CCAL_CUDA_OK(cudaGetLastError());
  *n_knots_host = n_knots;
  return CCAL_OK;
}

extern "C" int ccal_isotonic_transform(const double* knots_x, const double* knots_y, int64_t n_knots, const double* t,
                                       int64_t n, double residual_scale, double* out, ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE(n_knots >= 1 && n_knots < 2147483647ll, "ccal_isotonic_transform: n_knots out of range (got %lld)", (long long)n_knots);
  CCAL_REQUIRE(n >= 0, "ccal_isotonic_transform: negative n");
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(knots_x && knots_y && t && out, "ccal_isotonic_transform: NULL pointer");
  const long long grid = std::min<long long>((n + kIsoThreads - 1) / kIsoThreads, (long long)num_sms() * 16);
  if (n_knots <= kIsoSmemKnots)
    iso_transform_kernel<true><<<(int)grid, kIsoThreads, (size_t)n_knots * 16, stream>>>(knots_x, knots_y, (int)n_knots, t, n,
                                                                                         residual_scale, out);
  else
    iso_transform_kernel<false><<<(int)grid, kIsoThreads, 0, stream>>>(knots_x, knots_y, (int)n_knots, t, n, residual_scale, out);
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

// ---- the device primitives above, exported so that tests can drive them directly (tests/test_gpu_parity.py)
extern "C" int ccal_sort_pairs_f64_u8(const double* keys, const unsigned char* vals, int64_t n, double* keys_out,
                                      unsigned char* vals_out, ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE(n >= 0 && n < 2147483647ll, "ccal_sort_pairs_f64_u8: n must be in [0, 2^31) (got %lld)", (long long)n);
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(keys && vals && keys_out && vals_out, "ccal_sort_pairs_f64_u8: NULL pointer");
  CCAL_REQUIRE((const void*)keys != (const void*)keys_out && vals != vals_out, "ccal_sort_pairs_f64_u8: in-place sort is not supported");
  AsyncWorkspace ws;
  CCAL_CUDA_OK(ws.alloc(sort_plan(n).total, stream));
  CCAL_CUDA_OK(sort_pairs_f64_u8(keys, vals, n, keys_out, vals_out, ws.ptr, stream));
  return CCAL_OK;
}

extern "C" int ccal_prefix_sum_i32(const int* in, int* out, int64_t n, int inclusive, ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE(n >= 0, "ccal_prefix_sum_i32: negative n");
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(in && out, "ccal_prefix_sum_i32: NULL pointer");
  AsyncWorkspace ws;
  CCAL_CUDA_OK(ws.alloc(scan_workspace_bytes(n), stream));
  CCAL_CUDA_OK(prefix_sum_i32(in, out, n, inclusive != 0, ws.ptr, stream));
  return CCAL_OK;
}

// ---- one-vs-all (netcal-style) calibrators
extern "C" int ccal_ova_hist_fit(const float* p_f32, const double* p_f64, int64_t n, int c, const int64_t* labels,
                                 const double* edges, int n_bins, uint32_t* count, uint32_t* hits, ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE((p_f32 != nullptr) != (p_f64 != nullptr), "ccal_ova_hist_fit: exactly one of p_f32 / p_f64 must be given");
  CCAL_REQUIRE(n >= 0 && n < (1ll << 32) && c >= 1, "ccal_ova_hist_fit: bad shape n=%lld c=%d (n < 2^32: 32-bit counters)", (long long)n, c);
  CCAL_REQUIRE(n_bins >= 1 && n_bins <= kOvaMaxBins, "ccal_ova_hist_fit: n_bins must be in [1, %d] (got %d)", kOvaMaxBins, n_bins);
  CCAL_REQUIRE(edges && count && hits && (labels || n == 0), "ccal_ova_hist_fit: NULL pointer");
  CCAL_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(uint32_t) * (size_t)c * n_bins, stream));
  CCAL_CUDA_OK(cudaMemsetAsync(hits, 0, sizeof(uint32_t) * (size_t)c * n_bins, stream));
  if (n == 0) return CCAL_OK;
  const int tile_c = std::min(c, 20480 / n_bins);              // 2 x tile_c x n_bins counters <= 160 KB of shared memory
  const size_t smem = 2 * sizeof(unsigned) * (size_t)tile_c * n_bins;
  const int tiles = (c + tile_c - 1) / tile_c;
  const unsigned gx = (unsigned)std::min<long long>((n + 31) / 32, std::max(1, num_sms() * 2 / tiles));   // two 1024-thread CTAs per SM (32 registers, <= 80 KB of counters each)
  const long long* lab = reinterpret_cast<const long long*>(labels);
#define CCAL_LAUNCH_OVA_FIT(T, ptr)                                                                                   \
  do {                                                                                                                \
    CCAL_CUDA_OK(cudaFuncSetAttribute(ova_hist_fit_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    ova_hist_fit_kernel<T><<<dim3(gx, tiles), kOvaFitThreads, smem, stream>>>(ptr, n, c, lab, edges, n_bins, tile_c, count, hits); \
  } while (0)
  if (p_f32) CCAL_LAUNCH_OVA_FIT(float, p_f32); else CCAL_LAUNCH_OVA_FIT(double, p_f64);
#undef CCAL_LAUNCH_OVA_FIT
  ova_hist_finish_kernel<<<(c + 255) / 256, 256, 0, stream>>>(count, c, n_bins, (unsigned)n);
  note_launch(2);
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

extern "C" int ccal_ova_apply(const float* p_f32, const double* p_f64, int64_t n, int c, const double* edges, int n_bins,
                              const double* bin_map, const double* knots_x, const double* knots_y, const int32_t* knot_off,
                              int normalise, double* out, ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE((p_f32 != nullptr) != (p_f64 != nullptr), "ccal_ova_apply: exactly one of p_f32 / p_f64 must be given");
  CCAL_REQUIRE(n >= 0 && c >= 1, "ccal_ova_apply: bad shape n=%lld c=%d", (long long)n, c);
  const bool hist = bin_map != nullptr;
  CCAL_REQUIRE(hist != (knot_off != nullptr), "ccal_ova_apply: give either bin_map (+ edges) or the knots (+ knot_off)");
  if (hist) CCAL_REQUIRE(edges && n_bins >= 1 && n_bins <= kOvaMaxBins, "ccal_ova_apply: edges / n_bins (1..%d) are required with bin_map", kOvaMaxBins);
  else CCAL_REQUIRE(knots_x && knots_y, "ccal_ova_apply: knots_x / knots_y are required with knot_off");
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(out != nullptr, "ccal_ova_apply: out is NULL");
  const unsigned gx = (unsigned)std::min<long long>((n + 7) / 8, (long long)num_sms() * 8);
#define CCAL_LAUNCH_OVA_APPLY(T, ptr)                                                                                  \
  do {                                                                                                                 \
    if (hist) ova_apply_kernel<T, 0><<<gx, 256, 0, stream>>>(ptr, n, c, edges, n_bins, bin_map, nullptr, nullptr, nullptr, normalise, out); \
    else ova_apply_kernel<T, 1><<<gx, 256, 0, stream>>>(ptr, n, c, nullptr, 0, nullptr, knots_x, knots_y, knot_off, normalise, out);        \
  } while (0)
  if (p_f32) CCAL_LAUNCH_OVA_APPLY(float, p_f32); else CCAL_LAUNCH_OVA_APPLY(double, p_f64);
#undef CCAL_LAUNCH_OVA_APPLY
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}
