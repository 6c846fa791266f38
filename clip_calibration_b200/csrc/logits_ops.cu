// K4: drop-ins that work on MATERIALISED logits [n, c] (fp32, row-major).
//   dac_predict_logits  : logits[i,:] *= class_conf[argmax_j logits[i,j]]            (in place)
//   logits_confidence   : pred_i, conf_i = max softmax(class_conf[pred_i] * logits[i,:])
//   dac_softmax_logits  : logits[i,:] <- softmax(class_conf[pred_i] * logits[i,:])     (in place)
//   row_argmax          : first argmax + row maximum
// HBM-bound: each row is read from HBM once (the later sweeps over the row hit L1/L2) and, for
// the in-place variants, written once.  Algorithmic bytes per image: 8*c in place, 4*c (+8 out)
// for the read-only variants.
// GROUP threads cooperate on one row: 1 (c <= 16), 8 (c <= 128), 32 (c <= 2048), 256 = the CTA.
// 128-bit loads/stores whenever rows are 16-byte aligned (c % 4 == 0).
#include "ccal_common.cuh"

#include <math_constants.h>

namespace ccal {

__device__ __forceinline__ float exp2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct MaxIdx {
  float v;
  int i;
};

__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) {
  // larger value wins; on equal values the lower index wins (numpy / first-max semantics)
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

template <int GROUP>
__device__ __forceinline__ MaxIdx group_argmax(MaxIdx m) {
#pragma unroll
  for (int off = (GROUP > 32 ? 32 : GROUP) / 2; off > 0; off >>= 1) {
    MaxIdx o;
    o.v = __shfl_xor_sync(0xffffffffu, m.v, off);
    o.i = __shfl_xor_sync(0xffffffffu, m.i, off);
    m = better(m, o);
  }
  return m;
}

template <int GROUP>
__device__ __forceinline__ float group_sum(float s) {
#pragma unroll
  for (int off = (GROUP > 32 ? 32 : GROUP) / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  return s;
}

enum RowOp { kScaleInPlace = 0, kConfidence = 1, kSoftmaxInPlace = 2, kArgmaxOnly = 3 };

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
  // uniform trip count per CTA so that the collectives / __syncthreads below are safe
  const long long sweeps = (n + group_stride - 1) / group_stride;
  const bool vec = (c % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0) && GROUP > 1;
  const int c4 = c >> 2;
  for (long long sweep = 0; sweep < sweeps; ++sweep) {
    const long long row = group0 + sweep * group_stride;
    const bool live = row < n;
    float* x = logits + (live ? row : 0) * (long long)c;
    float4* x4 = reinterpret_cast<float4*>(x);

    // ---- sweep 1: first argmax (a thread visits its classes in increasing index order: strict > suffices locally)
    MaxIdx m{-CUDART_INF_F, 0x7fffffff};
    if (live) {
      if (vec) {
#pragma unroll 4
        for (int j = t; j < c4; j += GROUP) {
          const float4 v = x4[j];
          if (v.x > m.v) { m.v = v.x; m.i = 4 * j; }
          if (v.y > m.v) { m.v = v.y; m.i = 4 * j + 1; }
          if (v.z > m.v) { m.v = v.z; m.i = 4 * j + 2; }
          if (v.w > m.v) { m.v = v.w; m.i = 4 * j + 3; }
        }
      } else {
#pragma unroll 4
        for (int j = t; j < c; j += GROUP) {
          const float v = x[j];
          if (v > m.v) { m.v = v; m.i = j; }
        }
      }
    }
    m = group_argmax<GROUP>(m);
    if (GROUP > 32) {
      if (lane == 0) s_red[warp] = m;
      __syncthreads();
      m = s_red[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) m = better(m, s_red[w]);
      __syncthreads();
    }
    const int pred = (m.i == 0x7fffffff) ? 0 : m.i;   // row of NaN / -inf only: no entry ever compared greater (numpy gives 0)
    const float cc = (live && class_conf) ? class_conf[pred] : 1.0f;

    if (OP == kArgmaxOnly) {
      if (live && t == 0) {
        if (pred_out) pred_out[row] = pred;
        if (conf_out) conf_out[row] = m.v;
      }
      continue;
    }
    if (OP == kScaleInPlace) {
      if (live) {
        // fp32 multiply, exactly what `logits *= class_confidences[pred][:, None]` does
        if (vec) {
#pragma unroll 4
          for (int j = t; j < c4; j += GROUP) {
            float4 v = x4[j];
            v.x = __fmul_rn(v.x, cc); v.y = __fmul_rn(v.y, cc); v.z = __fmul_rn(v.z, cc); v.w = __fmul_rn(v.w, cc);
            x4[j] = v;
          }
        } else {
          for (int j = t; j < c; j += GROUP) x[j] = __fmul_rn(x[j], cc);
        }
        if (t == 0 && pred_out) pred_out[row] = pred;
      }
      continue;
    }

    // ---- sweep 2: sum of exp of the DAC-scaled, max-shifted row (scipy softmax arithmetic)
    const float mcc = __fmul_rn(m.v, cc);
    // probabilities that are RETURNED use the reference's fp32 arithmetic and the accurate expf; the
    // confidence-only variant uses one FMA + ex2 per class (relative error ~2e-6)
    const float cc2 = cc * 1.4426950408889634f, mcc2 = mcc * 1.4426950408889634f;
    auto ex = [&](float v) {
      if (OP == kConfidence) return exp2f_approx(fmaf(v, cc2, -mcc2));
      return expf(__fsub_rn(__fmul_rn(v, cc), mcc));
    };
    float s = 0.f;
    if (live) {
      if (vec) {
#pragma unroll 4
        for (int j = t; j < c4; j += GROUP) {
          const float4 v = x4[j];
          s += (ex(v.x) + ex(v.y)) + (ex(v.z) + ex(v.w));
        }
      } else {
        for (int j = t; j < c; j += GROUP) s += ex(x[j]);
      }
    }
    s = group_sum<GROUP>(s);
    if (GROUP > 32) {
      if (lane == 0) s_sum[warp] = s;
      __syncthreads();
      s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += s_sum[w];
      __syncthreads();
    }
    if (OP == kSoftmaxInPlace && live) {
      if (vec) {
#pragma unroll 4
        for (int j = t; j < c4; j += GROUP) {
          float4 v = x4[j];
          v.x = __fdiv_rn(ex(v.x), s); v.y = __fdiv_rn(ex(v.y), s); v.z = __fdiv_rn(ex(v.z), s); v.w = __fdiv_rn(ex(v.w), s);
          x4[j] = v;
        }
      } else {
        for (int j = t; j < c; j += GROUP) x[j] = __fdiv_rn(ex(x[j]), s);
      }
    }
    if (live && t == 0) {
      if (pred_out) pred_out[row] = pred;
      if (conf_out) conf_out[row] = 1.0f / s;
    }
  }
}


// Register-resident variant for 16-byte-aligned rows of up to 2048 classes: one warp per row, the whole
// row (NV4 float4 per lane) is loaded ONCE with all loads in flight, and every later sweep (max, sum of
// exp, scaling, write-back) runs from registers.  HBM traffic = the algorithmic bytes.
template <int NV4, int OP>
__global__ void __launch_bounds__(256)
logits_rows_reg_kernel(float* __restrict__ logits, const float* __restrict__ class_conf, long long n, int c,
                       int* __restrict__ pred_out, float* __restrict__ conf_out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long warp_stride = (long long)gridDim.x * 8;
  const int c4 = c >> 2;
  for (long long row = warp0; row < n; row += warp_stride) {
    float4* x4 = reinterpret_cast<float4*>(logits + row * (long long)c);
    float4 v[NV4];
#pragma unroll
    for (int u = 0; u < NV4; ++u) {
      const int j = lane + 32 * u;
      v[u] = (j < c4) ? __ldcs(x4 + j) : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    }
    // a lane visits its classes in increasing index order, so a strict > keeps the first maximum; the full
    // (value, index) tie rule is only needed across lanes
    MaxIdx m{-CUDART_INF_F, 0x7fffffff};
#pragma unroll
    for (int u = 0; u < NV4; ++u) {
      const int j = 4 * (lane + 32 * u);
      if (v[u].x > m.v) { m.v = v[u].x; m.i = j; }
      if (v[u].y > m.v) { m.v = v[u].y; m.i = j + 1; }
      if (v[u].z > m.v) { m.v = v[u].z; m.i = j + 2; }
      if (v[u].w > m.v) { m.v = v[u].w; m.i = j + 3; }
    }
    m = group_argmax<32>(m);
    const int pred = (m.i == 0x7fffffff) ? 0 : m.i;   // row of NaN / -inf only: no entry ever compared greater (numpy gives 0)
    const float cc = class_conf ? __ldg(class_conf + pred) : 1.0f;
    if (OP == kArgmaxOnly) {
      if (lane == 0) {
        if (pred_out) pred_out[row] = pred;
        if (conf_out) conf_out[row] = m.v;
      }
      continue;
    }
    if (OP == kScaleInPlace) {
#pragma unroll
      for (int u = 0; u < NV4; ++u) {
        const int j = lane + 32 * u;
        if (j < c4) {
          float4 o = v[u];
          o.x = __fmul_rn(o.x, cc); o.y = __fmul_rn(o.y, cc); o.z = __fmul_rn(o.z, cc); o.w = __fmul_rn(o.w, cc);
          __stcs(x4 + j, o);
        }
      }
      if (lane == 0 && pred_out) pred_out[row] = pred;
      continue;
    }
    const float mcc = __fmul_rn(m.v, cc);
    const float cc2 = cc * 1.4426950408889634f, mcc2 = mcc * 1.4426950408889634f;
    auto ex = [&](float t) {
      // returned probabilities: the reference's fp32 arithmetic (scale, shift, accurate exp); confidence only:
      // one FMA + ex2 per class (relative error ~2e-6, far inside the 1e-4 budget)
      if (OP == kConfidence) return exp2f_approx(fmaf(t, cc2, -mcc2));
      return expf(__fsub_rn(__fmul_rn(t, cc), mcc));
    };
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < NV4; ++u) {                 // padded lanes hold -inf -> exp = 0
      v[u].x = ex(v[u].x); v[u].y = ex(v[u].y); v[u].z = ex(v[u].z); v[u].w = ex(v[u].w);
      s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    }
    s = group_sum<32>(s);
    if (OP == kSoftmaxInPlace) {
#pragma unroll
      for (int u = 0; u < NV4; ++u) {
        const int j = lane + 32 * u;
        if (j < c4) {
          float4 o = v[u];
          o.x = __fdiv_rn(o.x, s); o.y = __fdiv_rn(o.y, s); o.z = __fdiv_rn(o.z, s); o.w = __fdiv_rn(o.w, s);
          __stcs(x4 + j, o);
        }
      }
    }
    if (lane == 0) {
      if (pred_out) pred_out[row] = pred;
      if (conf_out) conf_out[row] = 1.0f / s;
    }
  }
}

template <int NV4, int OP>
static int launch_reg(float* logits, const float* class_conf, int64_t n, int c, int* pred_out, float* conf_out,
                      cudaStream_t stream) {
  const long long want = (n + 7) / 8;
  const long long cap = (long long)num_sms() * 8;
  const int grid = (int)(want < cap ? want : cap);
  logits_rows_reg_kernel<NV4, OP><<<grid, 256, 0, stream>>>(logits, class_conf, (long long)n, c, pred_out, conf_out);
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

template <int GROUP, int OP>
static int launch_group(float* logits, const float* class_conf, int64_t n, int c, int* pred_out, float* conf_out,
                        cudaStream_t stream) {
  const long long rows_per_cta = 256 / GROUP;
  const long long want = (n + rows_per_cta - 1) / rows_per_cta;
  const long long cap = (long long)num_sms() * 8;
  const int grid = (int)(want < cap ? want : cap);
  logits_rows_kernel<GROUP, OP><<<grid, 256, 0, stream>>>(logits, class_conf, (long long)n, c, pred_out, conf_out);
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

template <int OP>
static int launch_rows(float* logits, const float* class_conf, int64_t n, int c, int* pred_out, float* conf_out,
                       cudaStream_t stream) {
  if (c <= 16) return launch_group<1, OP>(logits, class_conf, n, c, pred_out, conf_out, stream);
  if (c <= 128) return launch_group<8, OP>(logits, class_conf, n, c, pred_out, conf_out, stream);
  if (c <= 2048 && c % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) {
    const int c4 = c / 4;
    if (c4 <= 64) return launch_reg<2, OP>(logits, class_conf, n, c, pred_out, conf_out, stream);
    if (c4 <= 128) return launch_reg<4, OP>(logits, class_conf, n, c, pred_out, conf_out, stream);
    if (c4 <= 256) return launch_reg<8, OP>(logits, class_conf, n, c, pred_out, conf_out, stream);
    return launch_reg<16, OP>(logits, class_conf, n, c, pred_out, conf_out, stream);
  }
  if (c <= 2048) return launch_group<32, OP>(logits, class_conf, n, c, pred_out, conf_out, stream);
  return launch_group<256, OP>(logits, class_conf, n, c, pred_out, conf_out, stream);
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
