// K4: drop-ins that work on MATERIALISED logits [n, c] (fp32, row-major).
//   dac_predict_logits : logits[i,:] *= class_conf[argmax_j logits[i,j]]     (in place)
//   logits_confidence  : pred_i, conf_i = max softmax(class_conf[pred_i] * logits[i,:])
// HBM-bound: each row is read from HBM once (the second sweep over the row hits L1/L2) and,
// for the in-place variant, written once.  Algorithmic bytes per image: 8*c (in place) or
// 4*c (+8 out) for the confidence variant.
// One warp per row while c <= 2048, one 256-thread CTA per row above that.
#include "ccal_common.cuh"

#include <math_constants.h>

namespace ccal {

struct MaxIdx {
  float v;
  int i;
};

__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) {
  // larger value wins; on equal values the lower index wins (numpy / first-max semantics)
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

__device__ __forceinline__ MaxIdx warp_argmax(MaxIdx m) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    MaxIdx o;
    o.v = __shfl_xor_sync(0xffffffffu, m.v, off);
    o.i = __shfl_xor_sync(0xffffffffu, m.i, off);
    m = better(m, o);
  }
  return m;
}

__device__ __forceinline__ float warp_sum(float s) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  return s;
}

enum RowOp { kScaleInPlace = 0, kConfidence = 1, kSoftmaxInPlace = 2, kArgmaxOnly = 3 };

// GROUP = threads cooperating on one row (32 = a warp, 256 = the CTA).  OP selects the operation.
template <int GROUP, int OP>
__global__ void __launch_bounds__(256)
logits_rows_kernel(float* __restrict__ logits, const float* __restrict__ class_conf, long long n, int c,
                   int* __restrict__ pred_out, float* __restrict__ conf_out) {
  __shared__ MaxIdx s_red[8];
  __shared__ float s_sum[8];
  constexpr int kGroupsPerCta = 256 / GROUP;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int t = threadIdx.x % GROUP;          // index inside the row group
  const long long group0 = (long long)blockIdx.x * kGroupsPerCta + threadIdx.x / GROUP;
  const long long group_stride = (long long)gridDim.x * kGroupsPerCta;
  // uniform trip count per CTA so that the __syncthreads below are safe
  const long long rows_per_sweep = group_stride;
  const long long sweeps = (n + rows_per_sweep - 1) / rows_per_sweep;
  for (long long sweep = 0; sweep < sweeps; ++sweep) {
    const long long row = group0 + sweep * group_stride;
    const bool live = row < n;
    float* x = logits + (live ? row : 0) * (long long)c;
    MaxIdx m{-CUDART_INF_F, 0x7fffffff};
    if (live)
      for (int j = t; j < c; j += GROUP) m = better(m, MaxIdx{x[j], j});
    m = warp_argmax(m);
    if (GROUP > 32) {
      if (lane == 0) s_red[warp] = m;
      __syncthreads();
      m = s_red[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) m = better(m, s_red[w]);
      __syncthreads();
    }
    const int pred = m.i;
    const float cc = (live && class_conf) ? class_conf[pred] : 1.0f;
    if (OP == kArgmaxOnly) {
      if (live && t == 0) {
        if (pred_out) pred_out[row] = pred;
        if (conf_out) conf_out[row] = m.v;
      }
    } else if (OP == kScaleInPlace) {
      if (live) {
        // fp32 multiply, exactly what `logits *= class_confidences[pred][:, None]` does
        for (int j = t; j < c; j += GROUP) x[j] = __fmul_rn(x[j], cc);
        if (t == 0 && pred_out) pred_out[row] = pred;
      }
    } else {
      const float mcc = __fmul_rn(m.v, cc);
      float s = 0.f;
      if (live)
        for (int j = t; j < c; j += GROUP) s += expf(__fsub_rn(__fmul_rn(x[j], cc), mcc));
      s = warp_sum(s);
      if (GROUP > 32) {
        if (lane == 0) s_sum[warp] = s;
        __syncthreads();
        s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_sum[w];
        __syncthreads();
      }
      if (OP == kSoftmaxInPlace && live) {
        // probabilities in place: exp(cc*x - cc*max) / sum  (scipy softmax of the DAC-scaled row)
        for (int j = t; j < c; j += GROUP) x[j] = __fdiv_rn(expf(__fsub_rn(__fmul_rn(x[j], cc), mcc)), s);
      }
      if (live && t == 0) {
        if (pred_out) pred_out[row] = pred;
        if (conf_out) conf_out[row] = 1.0f / s;
      }
    }
  }
}

template <int OP>
static int launch_rows(float* logits, const float* class_conf, int64_t n, int c, int* pred_out, float* conf_out,
                       cudaStream_t stream) {
  const int sms = num_sms();
  if (c <= 2048) {
    long long want = (n + 7) / 8;
    int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
    logits_rows_kernel<32, OP><<<grid, 256, 0, stream>>>(logits, class_conf, (long long)n, c, pred_out, conf_out);
  } else {
    long long want = n;
    int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
    logits_rows_kernel<256, OP><<<grid, 256, 0, stream>>>(logits, class_conf, (long long)n, c, pred_out, conf_out);
  }
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

}  // namespace ccal

using namespace ccal;

extern "C" int ccal_dac_predict_logits(float* logits, const float* class_conf, int64_t n, int c,
                                       int32_t* pred_out, ccal_stream_t stream) {
  CCAL_REQUIRE(n >= 0 && c >= 1, "ccal_dac_predict_logits: bad shape n=%lld c=%d", (long long)n, c);
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(logits && class_conf, "ccal_dac_predict_logits: NULL input");
  return launch_rows<kScaleInPlace>(logits, class_conf, n, c, pred_out, nullptr, (cudaStream_t)stream);
}

extern "C" int ccal_logits_confidence(const float* logits, const float* class_conf, int64_t n, int c,
                                      int32_t* pred_out, float* conf_out, ccal_stream_t stream) {
  CCAL_REQUIRE(n >= 0 && c >= 1, "ccal_logits_confidence: bad shape n=%lld c=%d", (long long)n, c);
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(logits, "ccal_logits_confidence: NULL input");
  return launch_rows<kConfidence>(const_cast<float*>(logits), class_conf, n, c, pred_out, conf_out, (cudaStream_t)stream);
}

extern "C" int ccal_dac_softmax_logits(float* logits, const float* class_conf, int64_t n, int c,
                                       int32_t* pred_out, float* conf_out, ccal_stream_t stream) {
  CCAL_REQUIRE(n >= 0 && c >= 1, "ccal_dac_softmax_logits: bad shape n=%lld c=%d", (long long)n, c);
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(logits, "ccal_dac_softmax_logits: NULL input");
  return launch_rows<kSoftmaxInPlace>(logits, class_conf, n, c, pred_out, conf_out, (cudaStream_t)stream);
}

extern "C" int ccal_row_argmax(const float* values, int64_t n, int c, int32_t* pred_out, float* max_out,
                               ccal_stream_t stream) {
  CCAL_REQUIRE(n >= 0 && c >= 1, "ccal_row_argmax: bad shape n=%lld c=%d", (long long)n, c);
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(values, "ccal_row_argmax: NULL input");
  return launch_rows<kArgmaxOnly>(const_cast<float*>(values), nullptr, n, c, pred_out, max_out, (cudaStream_t)stream);
}
