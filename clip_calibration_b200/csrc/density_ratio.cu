// f-4: density-ratio (proximity-informed) calibration, reference trainers/calibration/density_ratio_calibration.py.
//   kde2_pdf             : product-Gaussian kernel density of (confidence, proximity) points, evaluated at n query
//                          points against m data points (statsmodels KDEMultivariate 'cc' as the reference uses it,
//                          :66, :70, :104-105).  MUFU/FP64-bound: one ex2 and ~6 double operations per pair.
//   density_ratio_apply  : conf_cal = t / max(t + f * ratio, 1e-10); the predicted class of every row gets conf_cal,
//                          the other classes are rescaled to sum to 1 - conf_cal (:108-117).  HBM-bound: the row is
//                          read once from HBM (later sweeps hit L1/L2) and written once as float64.
//
// Range: proximity bandwidths are ~1e-3, so a query a few percent away from every validation point has a density of
// e^-500; the reference works in float64 and still forms the ratio t / (t + f * ratio) from such values.  The kernel
// therefore keeps every partial sum as (m, s) with  sum_i exp(e_i) = exp(m) * s,  m = running maximum exponent, the
// exponents e_i in double, exp(e_i - m) through one ex2.approx on the fp32 fraction (relative error 2e-7), s in double.
#include "ccal_common.cuh"
#include <math_constants.h>

#include <algorithm>

namespace ccal {

constexpr int kKdeThreads = 256;
constexpr int kKdeTile = 512;        // data points staged per shared-memory tile

// exp(x) for x <= 0 (double argument, fp32 result): split x*log2(e) = n + f in double, 2^f by ex2.approx, scale by 2^n
__device__ __forceinline__ float exp_neg(double x) {
  const double t = x * 1.4426950408889634;
  if (t < -149.0) return 0.f;
  const double nd = rint(t);
  const float f = (float)(t - nd);
  float r;
  asm("ex2.approx.f32 %0, %1;" : "=f"(r) : "f"(f));
  return ldexpf(r, (int)nd);
}

// grid (query tiles, data splits): partial (m, s) of every query over the split's data range
__global__ void __launch_bounds__(kKdeThreads)
kde2_partial_kernel(const double* __restrict__ dx, const double* __restrict__ dy, long long m_data,
                    const double* __restrict__ qx, const double* __restrict__ qy, long long n, double ax, double ay,
                    double* __restrict__ part_m, double* __restrict__ part_s) {
  __shared__ double2 tile[kKdeTile];
  const long long q = (long long)blockIdx.x * kKdeThreads + threadIdx.x;
  const bool live = q < n;
  const double x = live ? qx[q] : 0.0, y = live ? qy[q] : 0.0;
  const long long per = (m_data + gridDim.y - 1) / gridDim.y;
  const long long lo = per * blockIdx.y, hi = min(m_data, lo + per);
  double mx = -CUDART_INF, s = 0.0;
  for (long long base = lo; base < hi; base += kKdeTile) {
    const int cnt = (int)min((long long)kKdeTile, hi - base);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += kKdeThreads) tile[j] = make_double2(dx[base + j], dy[base + j]);
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const double2 p = tile[j];
      const double u = p.x - x, v = p.y - y;
      const double e = -(u * u * ax + v * v * ay);
      if (e > mx) {                       // new maximum: rescale what has been summed so far (rare after the first points)
        s = (mx == -CUDART_INF) ? 0.0 : s * (double)exp_neg(mx - e);
        mx = e;
      }
      s += (double)exp_neg(e - mx);
    }
  }
  if (live) {
    part_m[(long long)blockIdx.y * n + q] = mx;
    part_s[(long long)blockIdx.y * n + q] = s;
  }
}

// pdf = norm * sum over splits (in split order) of exp(m_p) * s_p, formed around the largest m_p
__global__ void kde2_finish_kernel(const double* __restrict__ part_m, const double* __restrict__ part_s, int splits,
                                   long long n, double norm, double* __restrict__ pdf_out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  double mx = -CUDART_INF;
  for (int p = 0; p < splits; ++p) mx = fmax(mx, part_m[(long long)p * n + q]);
  double s = 0.0;
  for (int p = 0; p < splits; ++p) {
    const double mp = part_m[(long long)p * n + q];
    if (mp != -CUDART_INF) s += part_s[(long long)p * n + q] * exp(mp - mx);
  }
  pdf_out[q] = (mx == -CUDART_INF) ? 0.0 : exp(mx) * s * norm;
}

// ----------------------------------------------------------------------------------------
// density_ratio_apply: GROUP threads per row (32 = one warp, 256 = the CTA)
// ----------------------------------------------------------------------------------------
template <typename T, int GROUP>
__global__ void __launch_bounds__(256)
density_ratio_rows_kernel(const T* __restrict__ probs, long long n, int c, const double* __restrict__ pdf_true,
                          const double* __restrict__ pdf_false, double ratio, double* __restrict__ out,
                          double* __restrict__ conf_cal_out, int* __restrict__ pred_out) {
  __shared__ double s_v[8];
  __shared__ int s_i[8];
  __shared__ double s_sum[8];
  constexpr int kRowsPerCta = 256 / GROUP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = threadIdx.x % GROUP;
  const long long stride = (long long)gridDim.x * kRowsPerCta;
  const long long sweeps = (n + stride - 1) / stride;
  for (long long sweep = 0; sweep < sweeps; ++sweep) {
    const long long row = (long long)blockIdx.x * kRowsPerCta + threadIdx.x / GROUP + sweep * stride;
    const bool live = row < n;
    const T* x = probs + (live ? row : 0) * (long long)c;
    // first argmax (np.argmax): a thread meets its classes in increasing order, strict > keeps the first
    double bv = -CUDART_INF;
    int bi = 0x7fffffff;
    if (live)
      for (int j = t; j < c; j += GROUP) {
        const double v = (double)x[j];
        if (v > bv) { bv = v; bi = j; }
      }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (GROUP > 32) {
      if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
      __syncthreads();
      bv = s_v[0]; bi = s_i[0];
      for (int w = 1; w < 8; ++w)
        if (s_v[w] > bv || (s_v[w] == bv && s_i[w] < bi)) { bv = s_v[w]; bi = s_i[w]; }
      __syncthreads();
    }
    if (bi == 0x7fffffff) bi = 0;                   // row of NaN / -inf only (np.argmax gives 0)
    // sum of the OTHER classes (the reference zeroes the predicted entry before summing)
    double rest = 0.0;
    if (live)
      for (int j = t; j < c; j += GROUP)
        if (j != bi) rest += (double)x[j];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) rest += __shfl_xor_sync(0xffffffffu, rest, off);
    if (GROUP > 32) {
      if (lane == 0) s_sum[warp] = rest;
      __syncthreads();
      rest = 0.0;
      for (int w = 0; w < 8; ++w) rest += s_sum[w];
      __syncthreads();
    }
    if (!live) continue;
    const double tr = pdf_true[row], fa = pdf_false[row];
    const double cal = tr / fmax(tr + fa * ratio, 1e-10);
    const double scale = (1.0 - cal) / rest;
    double* o = out + row * (long long)c;
    for (int j = t; j < c; j += GROUP) o[j] = (j == bi) ? cal : (double)x[j] * scale;
    if (t == 0) {
      if (conf_cal_out) conf_cal_out[row] = cal;
      if (pred_out) pred_out[row] = bi;
    }
  }
}

template <typename T>
static int launch_rows(const T* probs, int64_t n, int c, const double* pt, const double* pf, double ratio, double* out,
                       double* cal, int* pred, cudaStream_t stream) {
  const long long cap = (long long)num_sms() * 8;
  if (c <= 2048) {
    const long long grid = std::min<long long>((n + 7) / 8, cap);
    density_ratio_rows_kernel<T, 32><<<(int)grid, 256, 0, stream>>>(probs, n, c, pt, pf, ratio, out, cal, pred);
  } else {
    const long long grid = std::min<long long>(n, cap);
    density_ratio_rows_kernel<T, 256><<<(int)grid, 256, 0, stream>>>(probs, n, c, pt, pf, ratio, out, cal, pred);
  }
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

}  // namespace ccal

using namespace ccal;

extern "C" int ccal_kde2_pdf(const double* data_x, const double* data_y, int64_t m, const double* query_x,
                                      const double* query_y, int64_t n, double bw_x, double bw_y, double* pdf_out,
                             ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE(data_x && data_y && query_x && query_y && pdf_out, "ccal_kde2_pdf: null pointer");
  CCAL_REQUIRE(m >= 1, "ccal_kde2_pdf: needs at least one data point (got %lld)", (long long)m);
  CCAL_REQUIRE(n >= 0, "ccal_kde2_pdf: negative query count");
  CCAL_REQUIRE(bw_x > 0.0 && bw_y > 0.0, "ccal_kde2_pdf: bandwidths must be positive (got %g, %g)", bw_x, bw_y);
  if (n == 0) return CCAL_OK;
  const long long qtiles = (n + kKdeThreads - 1) / kKdeThreads;
  // enough (query tile x data split) CTAs to fill the device a few times over, at least one tile of data per split
  long long splits = std::max<long long>(1, std::min<long long>((4LL * num_sms() + qtiles - 1) / qtiles,
                                                               (m + kKdeTile - 1) / kKdeTile));
  splits = std::min<long long>(splits, 65535);
  AsyncWorkspace ws;
  CCAL_CUDA_OK(ws.alloc((size_t)2 * splits * n * sizeof(double), stream));
  double* part_m = reinterpret_cast<double*>(ws.ptr);
  double* part_s = part_m + splits * n;
  dim3 grid((unsigned)qtiles, (unsigned)splits);
  kde2_partial_kernel<<<grid, kKdeThreads, 0, stream>>>(data_x, data_y, m, query_x, query_y, n, 1.0 / (2.0 * bw_x * bw_x),
                                                       1.0 / (2.0 * bw_y * bw_y), part_m, part_s);
  note_launch();
  const double norm = 1.0 / ((double)m * 2.0 * 3.14159265358979323846 * bw_x * bw_y);
  kde2_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(part_m, part_s, (int)splits, n, norm, pdf_out);
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

extern "C" int ccal_density_ratio_apply(const float* probs_f32, const double* probs_f64, int64_t n, int c,
                                                 const double* pdf_true, const double* pdf_false, double false_true_ratio,
                                                 double* probs_out, double* conf_cal_out, int32_t* pred_out,
                                        ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE((probs_f32 != nullptr) != (probs_f64 != nullptr),
               "ccal_density_ratio_apply: exactly one of probs_f32 / probs_f64 must be given");
  CCAL_REQUIRE(pdf_true && pdf_false && probs_out, "ccal_density_ratio_apply: null pointer");
  CCAL_REQUIRE(n >= 0 && c >= 1, "ccal_density_ratio_apply: bad shape n=%lld c=%d", (long long)n, c);
  if (n == 0) return CCAL_OK;
  if (probs_f32) return launch_rows<float>(probs_f32, n, c, pdf_true, pdf_false, false_true_ratio, probs_out, conf_cal_out, pred_out, stream);
  return launch_rows<double>(probs_f64, n, c, pdf_true, pdf_false, false_true_ratio, probs_out, conf_cal_out, pred_out, stream);
}
