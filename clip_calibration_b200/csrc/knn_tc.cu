// K1 (tensor-core path): k nearest reference rows by exact fp32 Euclidean distance, found with a
// tcgen05 GEMM as a FILTER and verified exactly.
//
//   1. split_rows_kernel : x -> hi = bf16(x), lo = bf16(x - hi), 0.5*|x|^2 (fp32), plus one extra feature
//                          column (1 for queries, -|r|^2/2 for references) that folds the norm term into the GEMM.
//   2. knn_tc_kernel     : score = q.r - |r|^2/2 (= -d^2/2 + const) from three bf16 MMAs (hi.hi + hi.lo + lo.hi,
//                          fp32 accumulate in TMEM; the dropped lo.lo term is < 2^-18 |q||r|); the epilogue
//                          keeps per query the KP best candidates (branch-free prefilter, sorted insert on hits).
//   3. knn_verify_kernel : recomputes ||r - q||_2 for the KP candidates in fp32 with the
//                          reference's own formula, sorts them (distance, index) and PROVES that no
//                          discarded reference row can beat the k-th: every discarded row has
//                          approximate d^2 >= d2_cut, so exact d^2 >= d2_cut - eps.  Queries where
//                          the proof fails (ties at the cut, pathological data) are appended to a
//                          list and
//   4. knn_l2_kernel (list mode, csrc/knn_dac.cu) redoes them with the exhaustive exact scan.
//   The result is therefore always the exact-arithmetic answer; the GEMM only prunes.
//
// Pipeline = the fused scoring kernel's (warp 0 TMA producer, warp 1 MMA issuer, warps 2-5
// epilogue, two TMEM accumulator stages), both operands streamed: stage = {Qhi, Qlo, Rhi, Rlo}.
#include "ccal_common.cuh"
#include "sm100_ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdlib.h>

namespace ccal {

constexpr int kTcBlockM = 128, kTcBlockN = 256, kTcBlockK = 64, kTcUmmaK = 16;
constexpr int kTcQBytes = kTcBlockM * kTcBlockK * 2;     // 16 KB
constexpr int kTcRBytes = kTcBlockN * kTcBlockK * 2;     // 32 KB
constexpr int kTcStages = 3;                              // barrier slots (pairs use 3 stages, single CTAs 2)
constexpr int kTcThreads = 192;
constexpr int kTcCtlBytes = 1024;
constexpr int kTcScratchBytes = 32 * 128 * 4;            // epilogue staging: 32 scores of each of the 128 epilogue threads

struct __align__(16) KnnCtl {
  uint64_t full[kTcStages];
  uint64_t empty[kTcStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// ---------------------------------------------------------------------------------------- 1
// Operand modes.  Features that are exactly representable in a 16-bit format - the reference's default precision is
// fp16 (train.py:152), so its cached text / image features are; so are bf16 checkpoints - need no hi/lo split: one
// MMA per K step gives exact products.  exact16_scan_kernel decides once per call, on the device:
//   mode 1: every element of both matrices is a bf16 value   -> operands = bf16(x), one MMA
//   mode 2: every element is an fp16 value                    -> operands = fp16(x), one MMA (kind::f16, fp16 inputs)
//   mode 0: anything else -> x = hi + lo in bf16, three MMAs (hi.hi + hi.lo + lo.hi)
// flags[0] != 0: some element is not a bf16 value; flags[1] != 0: some element is not an fp16 value of magnitude <= 4.
__device__ __forceinline__ int operand_mode(const unsigned int* __restrict__ flags) {
  return flags[0] == 0u ? 1 : (flags[1] == 0u ? 2 : 0);
}

__global__ void __launch_bounds__(256)
exact16_scan_kernel(const float* __restrict__ x, long long n_elems, unsigned int* __restrict__ flags) {
  unsigned int not_bf16 = 0u, not_f16 = 0u;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems / 4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      not_bf16 |= (__bfloat162float(__float2bfloat16_rn(f[u])) != f[u]) ? 1u : 0u;
      // fp16 mode also needs the folded -|r|^2/2 (<= d * max|x|^2 / 2) far inside fp16's range: |x| <= 4, d <= 1024
      not_f16 |= (__half2float(__float2half_rn(f[u])) != f[u] || !(fabsf(f[u]) <= 4.0f)) ? 1u : 0u;
    }
  }
  not_bf16 = __any_sync(0xffffffffu, not_bf16 != 0u);
  not_f16 = __any_sync(0xffffffffu, not_f16 != 0u);
  if ((threadIdx.x & 31) == 0) {
    if (not_bf16) atomicOr(flags, 1u);
    if (not_f16) atomicOr(flags + 1, 1u);
  }
}

// Output rows are d + 64 wide: the extra 64-feature block folds the -|r|^2/2 term of the score into the GEMM itself.
// Queries carry 1.0 in its first three columns, reference rows a three-term 16-bit expansion of -|r|^2/2 (each term
// takes the next 8 or 11 mantissa bits, so the sum is the fp32 value exactly); the accumulator is then directly
// q.r - |r|^2/2 and the filter's epilogue is a bare compare per element.  In mode 0 the lo matrix's extra block is zero.
template <typename T16> __device__ __forceinline__ T16 to16(float v);
template <> __device__ __forceinline__ __nv_bfloat16 to16<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half to16<__half>(float v) { return __float2half_rn(v); }
template <typename T16> __device__ __forceinline__ float from16(T16 v);
template <> __device__ __forceinline__ float from16<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float from16<__half>(__half v) { return __half2float(v); }

template <typename T16, bool kWithLo>
__device__ __forceinline__ void split_row(const float* __restrict__ x, long long row, int d, T16* __restrict__ hi,
                                          T16* __restrict__ lo, float* __restrict__ half_norm2,
                                          unsigned int* __restrict__ max_norm2_bits, int is_reference, int lane) {
  const int dp = d + kTcBlockK;
  const float4* src = reinterpret_cast<const float4*>(x + row * d);
  float acc = 0.f;
  for (int j = lane; j < d / 4; j += 32) {
    const float4 v = src[j];
    const float f[4] = {v.x, v.y, v.z, v.w};
    T16 h[4], l[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      h[u] = to16<T16>(f[u]);
      l[u] = to16<T16>(f[u] - from16<T16>(h[u]));
      acc = fmaf(f[u], f[u], acc);
    }
    *reinterpret_cast<uint2*>(hi + row * dp + 4 * j) = *reinterpret_cast<uint2*>(h);
    if (kWithLo) *reinterpret_cast<uint2*>(lo + row * dp + 4 * j) = *reinterpret_cast<uint2*>(l);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  // extra block: columns d .. d+2 carry the fold term, the rest is zero
  float e[3];
  if (is_reference) {
    float rest = -0.5f * acc;
#pragma unroll
    for (int u = 0; u < 3; ++u) { e[u] = from16<T16>(to16<T16>(rest)); rest -= e[u]; }
  } else {
    e[0] = e[1] = e[2] = 1.0f;
  }
  for (int j = lane; j < kTcBlockK; j += 32) {
    hi[row * dp + d + j] = to16<T16>(j < 3 ? e[j] : 0.f);
    if (kWithLo) lo[row * dp + d + j] = to16<T16>(0.f);
  }
  if (lane == 0) {
    half_norm2[row] = 0.5f * acc;
    if (max_norm2_bits) atomicMax(max_norm2_bits, __float_as_uint(acc));   // non-negative floats order like uints
  }
}

__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ x, long long rows, int d, __nv_bfloat16* __restrict__ hi,
                  __nv_bfloat16* __restrict__ lo, float* __restrict__ half_norm2, unsigned int* __restrict__ max_norm2_bits,
                  int is_reference, const unsigned int* __restrict__ flags) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int mode = operand_mode(flags);                 // uniform over the grid
  if (mode == 0) split_row<__nv_bfloat16, true>(x, row, d, hi, lo, half_norm2, max_norm2_bits, is_reference, lane);
  else if (mode == 1) split_row<__nv_bfloat16, false>(x, row, d, hi, lo, half_norm2, max_norm2_bits, is_reference, lane);
  else split_row<__half, false>(x, row, d, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), half_norm2,
                                max_norm2_bits, is_reference, lane);
}

// ---------------------------------------------------------------------------------------- 2
struct KnnTcParams {
  long long nq, nr;
  int kblocks, n_col_tiles, n_row_tiles;
  uint32_t idesc;
  const float* r_half_norm2;     // [nr]
  const unsigned int* flags;     // operand mode (exact16_scan_kernel)
  int n_slots;                   // candidate lists per query: one per work unit that touches the query's row tile
  int* cand_idx;                 // [nq, n_slots, KP]  (-1 = empty; zero-filled slots are never written)
  float* cand_score;             // [nq, n_slots, KP]  GEMM score q.r - |r|^2/2 of each candidate, descending per slot
};

// Work decomposition (stream-K style): the (row tile, reference tile) pairs are numbered t = tile * NT + nt and cut
// into n_units equal contiguous ranges, one per CTA (pair) - every SM gets the same number of 256 x 256 x D tiles
// whatever the shape (86 row tiles on 74 pairs used to leave the second wave 16 % full).  A range may end inside
// a row tile; each unit that touches a row tile leaves its own candidate list for those queries (slot = unit - first
// unit touching the tile) and knn_verify_kernel merges the slots by score.
__device__ __forceinline__ long long unit_begin(long long u, long long total, long long n_units) { return u * total / n_units; }

// kCtas = 2: CTA pairs (cta_group::2): each CTA owns 128 query rows and loads only half (128 rows) of every
// reference tile; stage = {Qhi, Qlo, Rhi/2, Rlo/2} = 64 KB, 3 stages.  kCtas = 1: 96 KB stages, 2 stages.
template <int KP, int kCtas>
__global__ void __launch_bounds__(kTcThreads, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap map_qhi, const __grid_constant__ CUtensorMap map_qlo,
              const __grid_constant__ CUtensorMap map_rhi, const __grid_constant__ CUtensorMap map_rlo,
              const __grid_constant__ KnnTcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  KnnCtl* ctl = reinterpret_cast<KnnCtl*>(smem_dyn);
  const uint32_t op_base = (ptx::smem_u32(smem_dyn) + kTcCtlBytes + 1023u) & ~1023u;
  unsigned char* op_ptr = smem_dyn + (op_base - ptx::smem_u32(smem_dyn));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kRBytes = kTcRBytes / kCtas;                 // reference bytes this CTA loads per matrix per stage
  constexpr int kRRows = kTcBlockN / kCtas;
  constexpr int kStageBytes = 2 * kTcQBytes + 2 * kRBytes;
  constexpr int kStages = kCtas == 2 ? 3 : 2;
  constexpr int kTileRows = kTcBlockM * kCtas;
  const uint32_t rank = (kCtas == 2) ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int unit = (kCtas == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_units = (int)gridDim.x / kCtas;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_qhi); ptx::prefetch_tensormap(&map_qlo);
    ptx::prefetch_tensormap(&map_rhi); ptx::prefetch_tensormap(&map_rlo);
    for (int i = 0; i < kTcStages; ++i) { ptx::mbar_init(&ctl->full[i], 1); ptx::mbar_init(&ctl->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&ctl->tmem_full[i], 1); ptx::mbar_init(&ctl->tmem_empty[i], 4 * kCtas); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    if (kCtas == 2) { ptx::tmem_alloc_2sm(&ctl->tmem_base, 512); ptx::tmem_relinquish_2sm(); }
    else { ptx::tmem_alloc(&ctl->tmem_base, 512); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  if (kCtas == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const int NT = p.n_col_tiles, KB = p.kblocks;
  const long long total = (long long)p.n_row_tiles * NT;
  const long long t_lo = unit_begin(unit, total, n_units), t_hi = unit_begin(unit + 1, total, n_units);
  const int mode = operand_mode(p.flags);
  const bool single = mode != 0;                 // 16-bit-exact operands: hi tiles only, one MMA per K step

  if (warp == 0) {
    // TMA producer: whole warp loops, one elected lane issues (warp-uniform control flow)
    uint32_t stage = 0, phase = 0;
    for (long long t = t_lo; t < t_hi; ++t) {
      {
        const int tile = (int)(t / NT), nt = (int)(t % NT);
        const int row0 = tile * kTileRows + (int)rank * kTcBlockM;
        const int col0 = nt * kTcBlockN + (int)rank * kRRows;
        for (int kb = 0; kb < KB; ++kb) {
          unsigned char* sp = op_ptr + (size_t)stage * kStageBytes;
          ptx::mbar_wait(&ctl->empty[stage], phase ^ 1u);
          if (ptx::elect_one()) {
            if (leader) ptx::mbar_arrive_expect_tx(&ctl->full[stage], (single ? kStageBytes / 2 : kStageBytes) * kCtas);
            if (kCtas == 2) {
              ptx::tma_load_2d_2sm(sp, &map_qhi, &ctl->full[stage], kb * kTcBlockK, row0, ptx::kEvictNormal);
              if (!single) ptx::tma_load_2d_2sm(sp + kTcQBytes, &map_qlo, &ctl->full[stage], kb * kTcBlockK, row0, ptx::kEvictNormal);
              ptx::tma_load_2d_2sm(sp + 2 * kTcQBytes, &map_rhi, &ctl->full[stage], kb * kTcBlockK, col0, ptx::kEvictLast);
              if (!single) ptx::tma_load_2d_2sm(sp + 2 * kTcQBytes + kRBytes, &map_rlo, &ctl->full[stage], kb * kTcBlockK, col0, ptx::kEvictLast);
            } else {
              ptx::tma_load_2d(sp, &map_qhi, &ctl->full[stage], kb * kTcBlockK, row0, ptx::kEvictNormal);
              if (!single) ptx::tma_load_2d(sp + kTcQBytes, &map_qlo, &ctl->full[stage], kb * kTcBlockK, row0, ptx::kEvictNormal);
              ptx::tma_load_2d(sp + 2 * kTcQBytes, &map_rhi, &ctl->full[stage], kb * kTcBlockK, col0, ptx::kEvictLast);
              if (!single) ptx::tma_load_2d(sp + 2 * kTcQBytes + kRBytes, &map_rlo, &ctl->full[stage], kb * kTcBlockK, col0, ptx::kEvictLast);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer (pair leader): whole warp waits, one elected lane issues 12 MMAs + commit per 64-feature block
    if (leader) {
      uint32_t stage = 0, phase = 0, acc_it = 0;
      // instruction descriptor: A / B format bits 7-9 / 10-12 (1 = bf16, 0 = fp16) follow the operand mode
      const uint32_t fmt = (mode == 2) ? 0u : 1u;
      const uint32_t idesc = p.idesc | (fmt << 7) | (fmt << 10);
      for (long long t = t_lo; t < t_hi; ++t, ++acc_it) {
        {
          const uint32_t as = acc_it & 1u, aph = (acc_it >> 1) & 1u;
          ptx::mbar_wait(&ctl->tmem_empty[as], aph ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * kTcBlockN;
          for (int kb = 0; kb < KB; ++kb) {
            const uint32_t sp = op_base + stage * kStageBytes;
            ptx::mbar_wait(&ctl->full[stage], phase);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const uint64_t qhi = ptx::make_kmajor_sw128_desc(sp), qlo = ptx::make_kmajor_sw128_desc(sp + kTcQBytes);
              const uint64_t rhi = ptx::make_kmajor_sw128_desc(sp + 2 * kTcQBytes);
              const uint64_t rlo = ptx::make_kmajor_sw128_desc(sp + 2 * kTcQBytes + kRBytes);
              auto mma = [&](uint64_t ad, uint64_t bd, uint32_t acc) {
                if (kCtas == 2) ptx::umma_f16_2sm(d_tmem, ad, bd, idesc, acc); else ptx::umma_f16(d_tmem, ad, bd, idesc, acc);
              };
              if (single) {
#pragma unroll
                for (int k = 0; k < kTcBlockK / kTcUmmaK; ++k) {
                  const uint64_t o = (uint64_t)(k * 2);
                  mma(qhi + o, rhi + o, (uint32_t)((kb | k) != 0));
                }
              } else {
#pragma unroll
                for (int k = 0; k < kTcBlockK / kTcUmmaK; ++k) {
                  const uint64_t o = (uint64_t)(k * 2);
                  mma(qhi + o, rhi + o, (uint32_t)((kb | k) != 0));
                  mma(qhi + o, rlo + o, 1u);
                  mma(qlo + o, rhi + o, 1u);
                }
              }
              if (kCtas == 2) ptx::umma_commit_2sm(&ctl->empty[stage]); else ptx::umma_commit(&ctl->empty[stage]);
              if (kb == KB - 1) {
                if (kCtas == 2) ptx::umma_commit_2sm(&ctl->tmem_full[as]); else ptx::umma_commit(&ctl->tmem_full[as]);
              }
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    float* sc = reinterpret_cast<float*>(op_ptr + (size_t)kStages * kStageBytes);   // [32 columns][128 epilogue threads]
    const int et = quarter * 32 + lane;
    uint32_t acc_it = 0;
    for (long long t = t_lo; t < t_hi;) {
      // one segment = this unit's part [nt0, nt1) of one row tile
      const int tile = (int)(t / NT), nt0 = (int)(t % NT);
      const int nt1 = (int)min((long long)NT, (long long)nt0 + (t_hi - t));
      t += nt1 - nt0;
      const long long row = (long long)tile * kTileRows + (long long)rank * kTcBlockM + quarter * 32 + lane;
      float ts[KP];
      int ti[KP];
#pragma unroll
      for (int s = 0; s < KP; ++s) { ts[s] = -CUDART_INF_F; ti[s] = -1; }
      for (int nt = nt0; nt < nt1; ++nt, ++acc_it) {
        const uint32_t as = acc_it & 1u;
        ptx::mbar_wait(&ctl->tmem_full[as], (acc_it >> 1) & 1u);
        ptx::tc_fence_after();
        const long long valid = min((long long)kTcBlockN, p.nr - (long long)nt * kTcBlockN);
        for (int ch = 0; ch * 32 < valid; ++ch) {
          uint32_t raw[32];
          ptx::tmem_ld_32x32(tmem_base + lane_sel + as * kTcBlockN + ch * 32, raw);
          ptx::tmem_ld_wait(raw);
          const int col0 = nt * kTcBlockN + ch * 32;
          const int nv = (int)min(32ll, p.nr - col0);              // valid columns in this chunk (>= 1)
          // branch-free prefilter: which of the 32 scores beat the current worst kept candidate?
          const float worst = ts[KP - 1];
          uint32_t hits = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) hits |= (uint32_t)(__uint_as_float(raw[j]) > worst) << j;
          if (nv < 32) hits &= (1u << nv) - 1u;
          if (hits) {                                              // rare once the lists have warmed up
            // Stage this lane's 32 scores in shared memory and walk only ITS OWN hit bits (dynamic index): a warp
            // whose lanes hit at different columns used to step through all 32 columns with the insertion predicated
            // per lane - with few reference rows (lists never warm) that loop, not the GEMM, set the kernel's time.
#pragma unroll
            for (int j = 0; j < 32; ++j) sc[j * 128 + et] = __uint_as_float(raw[j]);
            while (hits) {
              const int j = __ffs(hits) - 1;
              hits &= hits - 1u;
              const float t = sc[j * 128 + et];
              if (t > ts[KP - 1]) {
                // sorted insert, descending; equal scores keep the earlier (lower) index ahead
                float ct = t;
                int ci = col0 + j;
                bool ins = false;
#pragma unroll
                for (int s = 0; s < KP; ++s) {
                  const bool c = ins || (ct > ts[s]);
                  const float ot = ts[s];
                  const int oi = ti[s];
                  ts[s] = c ? ct : ot; ti[s] = c ? ci : oi;
                  ct = c ? ot : ct; ci = c ? oi : ci;
                  ins = c;
                }
              }
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kCtas == 2) ptx::mbar_arrive_cluster(&ctl->tmem_empty[as], 0u); else ptx::mbar_arrive(&ctl->tmem_empty[as]);
        }
      }
      if (row < p.nq) {
        // first unit whose range reaches this row tile: the largest u with unit_begin(u) <= tile * NT
        const long long first = (((long long)tile * NT + 1) * n_units - 1) / total;
        const long long o = (row * p.n_slots + (unit - first)) * KP;
#pragma unroll
        for (int s = 0; s < KP; ++s) { p.cand_idx[o + s] = ti[s]; p.cand_score[o + s] = ts[s]; }
      }
    }
  }
  __syncwarp();
  ptx::tc_fence_before();
  if (kCtas == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) {
    if (kCtas == 2) ptx::tmem_dealloc_2sm(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------- 3
// One warp per query.  eps bounds the error of the GEMM-derived d^2.
constexpr int kMaxSlots = 4;                     // candidate lists per query (work units touching one row tile)

template <int KP>
__global__ void __launch_bounds__(256)
knn_verify_kernel(const float* __restrict__ ref, const float* __restrict__ query, long long nr, long long nq, int d,
                  int k, int drop_first, int n_slots, const int* __restrict__ cand_idx, const float* __restrict__ cand_score,
                  const float* __restrict__ q_half_norm2, const unsigned int* __restrict__ r_max_norm2_bits,
                  float* __restrict__ dist_out, int* __restrict__ idx_out, int* __restrict__ redo_list,
                  int* __restrict__ redo_count) {
  __shared__ int s_idx[8][KP];
  __shared__ float s_score[8][KP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * 8 + warp;
  if (q >= nq) return;
  const float4* qv = reinterpret_cast<const float4*>(query + q * d);
  // Merge the slots: the KP best GEMM scores over all of them, ties to the lower index.  Every row a unit discarded
  // scores at most that unit's KP-th entry, which is at most the merged KP-th score - the cut of the proof below.
  constexpr int kPerLane = (kMaxSlots * KP + 31) / 32;
  const int m_total = n_slots * KP;
  float es[kPerLane];
  int ei[kPerLane];
#pragma unroll
  for (int u = 0; u < kPerLane; ++u) {
    const int e = lane + 32 * u;
    ei[u] = e < m_total ? cand_idx[q * m_total + e] : -1;
    es[u] = (e < m_total && ei[u] >= 0) ? cand_score[q * m_total + e] : -CUDART_INF_F;
  }
  if (lane < KP) { s_idx[warp][lane] = -1; s_score[warp][lane] = -CUDART_INF_F; }
  __syncwarp();
  int rk[kPerLane];
#pragma unroll
  for (int u = 0; u < kPerLane; ++u) rk[u] = 0;
#pragma unroll
  for (int u2 = 0; u2 < kPerLane; ++u2) {
    if (u2 * 32 < m_total) {                                   // warp-uniform
      for (int l = 0; l < 32; ++l) {
        const float os = __shfl_sync(0xffffffffu, es[u2], l);
        const int oi = __shfl_sync(0xffffffffu, ei[u2], l);
#pragma unroll
        for (int u = 0; u < kPerLane; ++u)
          rk[u] += (oi >= 0 && (os > es[u] || (os == es[u] && oi < ei[u]))) ? 1 : 0;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kPerLane; ++u)
    if (ei[u] >= 0 && rk[u] < KP) { s_idx[warp][rk[u]] = ei[u]; s_score[warp][rk[u]] = es[u]; }
  __syncwarp();
  int my_idx = lane < KP ? s_idx[warp][lane] : -1;
  float cut = (s_idx[warp][KP - 1] >= 0) ? s_score[warp][KP - 1] : -CUDART_INF_F;
  const int want = k + (drop_first ? 1 : 0);                // list length the caller needs
  const int have = (int)min((long long)want, nr);
  const float qn2 = 2.0f * q_half_norm2[q];
  // |S_gemm - S| <= ~1.2e-5 |q||r| (dropped lo.lo + bf16 rounding of lo + fp32 accumulation);
  // d^2 carries twice that; |q||r| <= (|q|^2 + max|r|^2)/2; 4x safety factor
  const float eps = 5e-5f * (qn2 + __uint_as_float(*r_max_norm2_bits));
  // Candidates ranked beyond `want` whose GEMM score lies more than 4 eps (in d^2 units) below that of rank want-1
  // cannot reach the first `want` places: they are not fetched at all, and the best of them becomes the cut of the
  // proof (which then holds by construction).  Typically 5 of 8 candidate rows are read instead of 8 - this kernel
  // is bound by exactly that L2 traffic.
  if (nr > KP && want < KP) {
    const float s_want = s_score[warp][want - 1];
    const float s_mine = lane < KP ? s_score[warp][lane] : -CUDART_INF_F;
    const bool keep = lane < want || (my_idx >= 0 && 2.0f * (s_want - s_mine) < 4.0f * eps);
    const unsigned kept = __ballot_sync(0xffffffffu, lane < KP && my_idx >= 0 && keep);
    const int n_keep = __popc(kept);                         // ranks are score-sorted: the kept ones are ranks [0, n_keep)
    if (n_keep < KP && s_idx[warp][n_keep] >= 0) {
      cut = s_score[warp][n_keep];
      if (lane >= n_keep) my_idx = -1;
    }
  }
  float my_d2 = CUDART_INF_F;
  for (int c = 0; c < KP; ++c) {
    const int r = __shfl_sync(0xffffffffu, my_idx, c);
    if (r < 0) continue;                                   // warp-uniform
    const float4* rv = reinterpret_cast<const float4*>(ref + (long long)r * d);
    float acc = 0.f;
    for (int j = lane; j < d / 4; j += 32) {
      const float4 a = qv[j], b = rv[j];
      float df = b.x - a.x; acc = fmaf(df, df, acc);
      df = b.y - a.y; acc = fmaf(df, df, acc);
      df = b.z - a.z; acc = fmaf(df, df, acc);
      df = b.w - a.w; acc = fmaf(df, df, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == c) my_d2 = acc;
  }
  // rank of my candidate among all candidates by (distance, index)
  const float my_d = sqrtf(my_d2);
  int rank = 0;
  for (int c = 0; c < KP; ++c) {
    const float od = __shfl_sync(0xffffffffu, my_d, c);
    const int oi = __shfl_sync(0xffffffffu, my_idx, c);
    if (oi >= 0 && (od < my_d || (od == my_d && oi < my_idx))) ++rank;
  }
  // proof that nothing outside the candidate list belongs to the first `have` entries
  bool proven = true;
  if (nr > KP) {
    const float d2_cut = qn2 - 2.0f * cut;                  // approximate d^2 of the best discarded row (lower bound)
    // d2 of the candidate ranked have-1
    const unsigned who = __ballot_sync(0xffffffffu, my_idx >= 0 && rank == have - 1);
    const float kth_d2 = __shfl_sync(0xffffffffu, my_d2, who ? __ffs(who) - 1 : 0);
    proven = who != 0 && (kth_d2 < d2_cut - eps);
  }
  if (!proven) {
    if (lane == 0) redo_list[atomicAdd(redo_count, 1)] = (int)q;
    return;
  }
  const int slot = rank - (drop_first ? 1 : 0);
  if (my_idx >= 0 && slot >= 0 && slot < k && rank < have) {
    if (dist_out) dist_out[q * k + slot] = my_d;
    if (idx_out) idx_out[q * k + slot] = my_idx;
  }
  // unused tail (k larger than the number of reference rows)
  const int filled = have - (drop_first ? 1 : 0);
  if (lane >= filled && lane < k) {
    if (dist_out) dist_out[q * k + lane] = CUDART_INF_F;
    if (idx_out) idx_out[q * k + lane] = -1;
  }
}

// defined in knn_dac.cu: exhaustive exact scan, optionally restricted to a device-side query list
int launch_knn_exact(const float* ref, const float* query, int64_t nr, int64_t nq, int d, int k, int drop_first,
                     float* dist_out, int32_t* idx_out, const int* qlist, const int* qcount, cudaStream_t stream);

// Number of work units (CTAs or CTA pairs) and of candidate-list slots per query for the stream-K decomposition:
// equal contiguous ranges of (row tile, reference tile) pairs, at least ceil(NT / 3) pairs each so that no row tile is
// touched by more than kMaxSlots units.
struct TcPlan { int units, slots; };
static TcPlan plan_units(int n_row_tiles, int n_col_tiles, int max_units) {
  const long long total = (long long)n_row_tiles * n_col_tiles;
  const long long min_len = (n_col_tiles + 2) / 3;
  long long units = total / min_len;
  if (units > max_units) units = max_units;
  if (units < 1) units = 1;
  const long long len = total / units;                                 // shortest range
  long long slots = (n_col_tiles - 1 + len - 1) / len + 1;
  if (slots > n_col_tiles) slots = n_col_tiles;
  if (slots > units) slots = units;
  if (slots > kMaxSlots) slots = kMaxSlots;                           // (cannot happen: len >= ceil(NT / 3))
  return TcPlan{(int)units, (int)slots};
}

template <int KP>
static int run_tc(const float* ref, const float* query, int64_t nr, int64_t nq, int d, int k, int drop_first,
                  float* dist_out, int32_t* idx_out, const __nv_bfloat16* rhi, const __nv_bfloat16* rlo,
                  const float* rhn, const unsigned int* rmax, const unsigned int* flags, __nv_bfloat16* qhi, __nv_bfloat16* qlo, float* qhn, int* cand, float* cand_score,
                  int* redo_list, int* redo_count, cudaStream_t stream) {
  const int dp = d + kTcBlockK;                      // operand rows carry one extra 64-feature block (see split_rows_kernel)
  split_rows_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, stream>>>(query, nq, d, qhi, qlo, qhn, nullptr, 0, flags);
  note_launch();
  trace_mark(stream, 20);
  // CTA pairs when every pair gets at least one 256-row tile on most SMs; single CTAs for small query sets
  const int sms = num_sms();
  const int ctas = (nq >= (int64_t)kTcBlockM * 2 * (sms / 4)) ? 2 : 1;
  CUtensorMap mqh, mql, mrh, mrl;
  int rc;
  if ((rc = make_map(&mqh, qhi, nq, dp, kTcBlockM, CCAL_BF16))) return rc;
  if ((rc = make_map(&mql, qlo, nq, dp, kTcBlockM, CCAL_BF16))) return rc;
  if ((rc = make_map(&mrh, rhi, nr, dp, kTcBlockN / ctas, CCAL_BF16))) return rc;
  if ((rc = make_map(&mrl, rlo, nr, dp, kTcBlockN / ctas, CCAL_BF16))) return rc;
  KnnTcParams p{};
  p.nq = nq; p.nr = nr;
  p.kblocks = dp / kTcBlockK;
  p.n_col_tiles = (int)((nr + kTcBlockN - 1) / kTcBlockN);
  const int tile_rows = kTcBlockM * ctas;
  p.n_row_tiles = (int)((nq + tile_rows - 1) / tile_rows);
  p.idesc = (1u << 4) | ((uint32_t)(kTcBlockN >> 3) << 17) | ((uint32_t)(tile_rows >> 4) << 24);   // operand formats: in the kernel
  p.flags = flags;
  p.r_half_norm2 = rhn; p.cand_idx = cand; p.cand_score = cand_score;
  const size_t smem = kTcCtlBytes + 1024 + (ctas == 2 ? (size_t)3 * (2 * kTcQBytes + kTcRBytes) : (size_t)2 * (2 * kTcQBytes + 2 * kTcRBytes)) +
                      kTcScratchBytes;
  const TcPlan plan = plan_units(p.n_row_tiles, p.n_col_tiles, sms / ctas);
  p.n_slots = plan.slots;
  const int grid = plan.units * ctas;
  CCAL_CUDA_OK(cudaMemsetAsync(cand, 0xff, (size_t)nq * plan.slots * KP * sizeof(int), stream));   // -1 = empty slot
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  if (ctas == 2) {
    auto kern = knn_tc_kernel<KP, 2>;
    CCAL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfg.numAttrs = 1;
    CCAL_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, mqh, mql, mrh, mrl, p));
  } else {
    auto kern = knn_tc_kernel<KP, 1>;
    CCAL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfg.numAttrs = 0;
    CCAL_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, mqh, mql, mrh, mrl, p));
  }
  note_launch();
  trace_mark(stream, 21);
  CCAL_CUDA_OK(cudaGetLastError());
  knn_verify_kernel<KP><<<(unsigned)((nq + 7) / 8), 256, 0, stream>>>(ref, query, nr, nq, d, k, drop_first, plan.slots, cand, cand_score, qhn,
                                                                      rmax, dist_out, idx_out, redo_list, redo_count);
  note_launch();
  trace_mark(stream, 22);
  CCAL_CUDA_OK(cudaGetLastError());
  return launch_knn_exact(ref, query, nr, nq, d, k, drop_first, dist_out, idx_out, redo_list, redo_count, stream);
}

// Tensor-core kNN over query chunks; transient stream-ordered workspace only.
int knn_l2_tensor(const float* ref, const float* query, int64_t nr, int64_t nq, int d, int k, int drop_first,
                  float* dist_out, int32_t* idx_out, cudaStream_t stream) {
  const int need = k + (drop_first ? 1 : 0);
  // candidate-list length: a few more than needed; rows whose proof fails are redone exactly, so a short list only
  // costs time when neighbours are unusually dense (measured: 0-3 unproven rows per 20k-100k queries at KP = 8)
  int KP = need <= 5 ? 8 : (need <= 11 ? 16 : 24);
  if (const char* e = getenv("CCAL_KNN_KP")) {            // development aid: candidate-list length override
    const int v = atoi(e);
    if ((v == 8 || v == 16 || v == 24) && v > need) KP = v;
  }
  const int64_t chunk = nq < 262144 ? nq : 262144;
  const size_t bf = sizeof(__nv_bfloat16);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t dp = (size_t)d + kTcBlockK;
  const size_t o_rhi = take((size_t)nr * dp * bf), o_rlo = take((size_t)nr * dp * bf), o_rhn = take((size_t)nr * 4);
  const size_t o_qhi = take((size_t)chunk * dp * bf), o_qlo = take((size_t)chunk * dp * bf), o_qhn = take((size_t)chunk * 4);
  const size_t o_cand = take((size_t)chunk * kMaxSlots * KP * 4), o_cut = take((size_t)chunk * kMaxSlots * KP * 4);
  const size_t o_list = take((size_t)chunk * 4), o_cnt = take(256), o_rmax = take(256);
  AsyncWorkspace workspace;
  CCAL_CUDA_OK(workspace.alloc(off, stream));
  unsigned char* ws = workspace.ptr;
  __nv_bfloat16* rhi = (__nv_bfloat16*)(ws + o_rhi);
  __nv_bfloat16* rlo = (__nv_bfloat16*)(ws + o_rlo);
  float* rhn = (float*)(ws + o_rhn);
  unsigned int* rmax = (unsigned int*)(ws + o_rmax);
  unsigned int* flags = rmax + 4;
  trace_mark(stream, 10);                          // the workspace allocation is ordered before this mark
  cudaMemsetAsync(rmax, 0, 8 * sizeof(unsigned int), stream);
  trace_mark(stream, 11);
  // operand mode: are both matrices exactly representable in bf16 / fp16?  (one read pass, HBM-bound)
  {
    const int sms = num_sms();
    long long g = ((long long)nr * d / 4 + 255) / 256;
    exact16_scan_kernel<<<(int)(g < 8 * sms ? (g < 1 ? 1 : g) : 8 * sms), 256, 0, stream>>>(ref, (long long)nr * d, flags);
    g = ((long long)nq * d / 4 + 255) / 256;
    exact16_scan_kernel<<<(int)(g < 8 * sms ? (g < 1 ? 1 : g) : 8 * sms), 256, 0, stream>>>(query, (long long)nq * d, flags);
    note_launch(2);
  }
  trace_mark(stream, 12);
  split_rows_kernel<<<(unsigned)((nr + 7) / 8), 256, 0, stream>>>(ref, nr, d, rhi, rlo, rhn, rmax, 1, flags);
  note_launch();
  trace_mark(stream, 13);
  int rc = CCAL_OK;
  for (int64_t q0 = 0; q0 < nq && rc == CCAL_OK; q0 += chunk) {
    const int64_t m = (nq - q0) < chunk ? (nq - q0) : chunk;
    int* cnt = (int*)(ws + o_cnt);
    cudaMemsetAsync(cnt, 0, sizeof(int), stream);
    float* dptr = dist_out ? dist_out + q0 * k : nullptr;
    int32_t* iptr = idx_out ? idx_out + q0 * k : nullptr;
#define CCAL_RUN_TC(KPV)                                                                                       \
  rc = run_tc<KPV>(ref, query + q0 * d, nr, m, d, k, drop_first, dptr, iptr, rhi, rlo, rhn, rmax, flags,        \
                   (__nv_bfloat16*)(ws + o_qhi), (__nv_bfloat16*)(ws + o_qlo), (float*)(ws + o_qhn),            \
                   (int*)(ws + o_cand), (float*)(ws + o_cut), (int*)(ws + o_list), cnt, stream)
    if (KP == 8) CCAL_RUN_TC(8); else if (KP == 16) CCAL_RUN_TC(16); else CCAL_RUN_TC(24);
#undef CCAL_RUN_TC
    if (getenv("CCAL_KNN_DEBUG")) {          // development aid: how many rows needed the exhaustive redo
      int h = -1;
      cudaMemcpyAsync(&h, cnt, sizeof(int), cudaMemcpyDeviceToHost, stream);
      cudaStreamSynchronize(stream);
      fprintf(stderr, "ccal knn_l2_tensor: nr=%lld nq=%lld d=%d k=%d KP=%d unproven rows=%d\n", (long long)nr, (long long)m, d, k, KP, h);
    }
  }
  return rc;
}

}  // namespace ccal
