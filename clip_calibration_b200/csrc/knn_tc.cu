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
#include <math_constants.h>
#include <stdlib.h>

namespace ccal {

constexpr int kTcBlockM = 128, kTcBlockN = 256, kTcBlockK = 64, kTcUmmaK = 16;
constexpr int kTcQBytes = kTcBlockM * kTcBlockK * 2;     // 16 KB
constexpr int kTcRBytes = kTcBlockN * kTcBlockK * 2;     // 32 KB
constexpr int kTcStages = 3;                              // barrier slots (pairs use 3 stages, single CTAs 2)
constexpr int kTcThreads = 192;
constexpr int kTcCtlBytes = 1024;

struct __align__(16) KnnCtl {
  uint64_t full[kTcStages];
  uint64_t empty[kTcStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// ---------------------------------------------------------------------------------------- 1
// Output rows are d + 64 wide: the extra 64-feature block holds ONE non-zero column that folds the -|r|^2/2
// term of the score into the GEMM itself: queries get 1.0 there, reference rows get -|r|^2/2 (as a hi/lo pair),
// so the accumulator is directly  q.r - |r|^2/2  and the filter's epilogue is a bare compare per element.
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ x, long long rows, int d, __nv_bfloat16* __restrict__ hi,
                  __nv_bfloat16* __restrict__ lo, float* __restrict__ half_norm2, unsigned int* __restrict__ max_norm2_bits,
                  int is_reference) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int dp = d + kTcBlockK;
  const float4* src = reinterpret_cast<const float4*>(x + row * d);
  float acc = 0.f;
  for (int j = lane; j < d / 4; j += 32) {
    const float4 v = src[j];
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      h[u] = __float2bfloat16_rn(f[u]);
      l[u] = __float2bfloat16_rn(f[u] - __bfloat162float(h[u]));
      acc = fmaf(f[u], f[u], acc);
    }
    *reinterpret_cast<uint2*>(hi + row * dp + 4 * j) = *reinterpret_cast<uint2*>(h);
    *reinterpret_cast<uint2*>(lo + row * dp + 4 * j) = *reinterpret_cast<uint2*>(l);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  // extra block: column d carries the fold term, columns d+1 .. d+63 are zero
  const float extra = is_reference ? -0.5f * acc : 1.0f;
  const __nv_bfloat16 eh = __float2bfloat16_rn(extra);
  const __nv_bfloat16 el = __float2bfloat16_rn(extra - __bfloat162float(eh));
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  for (int j = lane; j < kTcBlockK; j += 32) {
    hi[row * dp + d + j] = (j == 0) ? eh : zero;
    lo[row * dp + d + j] = (j == 0) ? el : zero;
  }
  if (lane == 0) {
    half_norm2[row] = 0.5f * acc;
    if (max_norm2_bits) atomicMax(max_norm2_bits, __float_as_uint(acc));   // non-negative floats order like uints
  }
}

// ---------------------------------------------------------------------------------------- 2
struct KnnTcParams {
  long long nq, nr;
  int kblocks, n_col_tiles, n_row_tiles;
  uint32_t idesc;
  const float* r_half_norm2;     // [nr]
  int* cand_idx;                 // [nq, KP]
  float* cand_cut;               // [nq] score of the worst kept candidate (-inf if the list is not full)
};

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

  if (warp == 0) {
    // TMA producer: whole warp loops, one elected lane issues (warp-uniform control flow)
    uint32_t stage = 0, phase = 0;
    for (int tile = unit; tile < p.n_row_tiles; tile += n_units) {
      const int row0 = tile * kTileRows + (int)rank * kTcBlockM;
      for (int nt = 0; nt < NT; ++nt) {
        const int col0 = nt * kTcBlockN + (int)rank * kRRows;
        for (int kb = 0; kb < KB; ++kb) {
          unsigned char* sp = op_ptr + (size_t)stage * kStageBytes;
          ptx::mbar_wait(&ctl->empty[stage], phase ^ 1u);
          if (ptx::elect_one()) {
            if (leader) ptx::mbar_arrive_expect_tx(&ctl->full[stage], kStageBytes * kCtas);
            if (kCtas == 2) {
              ptx::tma_load_2d_2sm(sp, &map_qhi, &ctl->full[stage], kb * kTcBlockK, row0, ptx::kEvictNormal);
              ptx::tma_load_2d_2sm(sp + kTcQBytes, &map_qlo, &ctl->full[stage], kb * kTcBlockK, row0, ptx::kEvictNormal);
              ptx::tma_load_2d_2sm(sp + 2 * kTcQBytes, &map_rhi, &ctl->full[stage], kb * kTcBlockK, col0, ptx::kEvictLast);
              ptx::tma_load_2d_2sm(sp + 2 * kTcQBytes + kRBytes, &map_rlo, &ctl->full[stage], kb * kTcBlockK, col0, ptx::kEvictLast);
            } else {
              ptx::tma_load_2d(sp, &map_qhi, &ctl->full[stage], kb * kTcBlockK, row0, ptx::kEvictNormal);
              ptx::tma_load_2d(sp + kTcQBytes, &map_qlo, &ctl->full[stage], kb * kTcBlockK, row0, ptx::kEvictNormal);
              ptx::tma_load_2d(sp + 2 * kTcQBytes, &map_rhi, &ctl->full[stage], kb * kTcBlockK, col0, ptx::kEvictLast);
              ptx::tma_load_2d(sp + 2 * kTcQBytes + kRBytes, &map_rlo, &ctl->full[stage], kb * kTcBlockK, col0, ptx::kEvictLast);
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
      for (int tile = unit; tile < p.n_row_tiles; tile += n_units)
        for (int nt = 0; nt < NT; ++nt, ++acc_it) {
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
                if (kCtas == 2) ptx::umma_f16_2sm(d_tmem, ad, bd, p.idesc, acc); else ptx::umma_f16(d_tmem, ad, bd, p.idesc, acc);
              };
#pragma unroll
              for (int k = 0; k < kTcBlockK / kTcUmmaK; ++k) {
                const uint64_t o = (uint64_t)(k * 2);
                mma(qhi + o, rhi + o, (uint32_t)((kb | k) != 0));
                mma(qhi + o, rlo + o, 1u);
                mma(qlo + o, rhi + o, 1u);
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
  } else {
    const int quarter = warp & 3;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    uint32_t acc_it = 0;
    for (int tile = unit; tile < p.n_row_tiles; tile += n_units) {
      const long long row = (long long)tile * kTileRows + (long long)rank * kTcBlockM + quarter * 32 + lane;
      float ts[KP];
      int ti[KP];
#pragma unroll
      for (int s = 0; s < KP; ++s) { ts[s] = -CUDART_INF_F; ti[s] = -1; }
      for (int nt = 0; nt < NT; ++nt, ++acc_it) {
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
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float t = __uint_as_float(raw[j]);
              if (((hits >> j) & 1u) && t > ts[KP - 1]) {
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
#pragma unroll
        for (int s = 0; s < KP; ++s) p.cand_idx[row * KP + s] = ti[s];
        p.cand_cut[row] = (ti[KP - 1] >= 0) ? ts[KP - 1] : -CUDART_INF_F;
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
template <int KP>
__global__ void __launch_bounds__(256)
knn_verify_kernel(const float* __restrict__ ref, const float* __restrict__ query, long long nr, long long nq, int d,
                  int k, int drop_first, const int* __restrict__ cand_idx, const float* __restrict__ cand_cut,
                  const float* __restrict__ q_half_norm2, const unsigned int* __restrict__ r_max_norm2_bits,
                  float* __restrict__ dist_out, int* __restrict__ idx_out, int* __restrict__ redo_list,
                  int* __restrict__ redo_count) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * 8 + warp;
  if (q >= nq) return;
  const float4* qv = reinterpret_cast<const float4*>(query + q * d);
  const int my_idx = lane < KP ? cand_idx[q * KP + lane] : -1;
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
  const int want = k + (drop_first ? 1 : 0);                // list length the caller needs
  const int have = (int)min((long long)want, nr);
  // proof that nothing outside the candidate list belongs to the first `have` entries
  const float cut = cand_cut[q];
  bool proven = true;
  if (nr > KP) {
    const float qn2 = 2.0f * q_half_norm2[q];
    const float d2_cut = qn2 - 2.0f * cut;                  // approximate d^2 of the best discarded row (lower bound)
    // |S_gemm - S| <= ~1.2e-5 |q||r| (dropped lo.lo + bf16 rounding of lo + fp32 accumulation);
    // d^2 carries twice that; |q||r| <= (|q|^2 + max|r|^2)/2; 4x safety factor
    const float eps = 5e-5f * (qn2 + __uint_as_float(*r_max_norm2_bits));
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

template <int KP>
static int run_tc(const float* ref, const float* query, int64_t nr, int64_t nq, int d, int k, int drop_first,
                  float* dist_out, int32_t* idx_out, const __nv_bfloat16* rhi, const __nv_bfloat16* rlo,
                  const float* rhn, const unsigned int* rmax, __nv_bfloat16* qhi, __nv_bfloat16* qlo, float* qhn, int* cand, float* cut,
                  int* redo_list, int* redo_count, cudaStream_t stream) {
  const int dp = d + kTcBlockK;                      // operand rows carry one extra 64-feature block (see split_rows_kernel)
  split_rows_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, stream>>>(query, nq, d, qhi, qlo, qhn, nullptr, 0);
  note_launch();
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
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcBlockN >> 3) << 17) | ((uint32_t)(tile_rows >> 4) << 24);
  p.r_half_norm2 = rhn; p.cand_idx = cand; p.cand_cut = cut;
  const size_t smem = kTcCtlBytes + 1024 + (ctas == 2 ? (size_t)3 * (2 * kTcQBytes + kTcRBytes) : (size_t)2 * (2 * kTcQBytes + 2 * kTcRBytes));
  const int units = sms / ctas;
  const int grid = (p.n_row_tiles < units ? p.n_row_tiles : units) * ctas;
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
  CCAL_CUDA_OK(cudaGetLastError());
  knn_verify_kernel<KP><<<(unsigned)((nq + 7) / 8), 256, 0, stream>>>(ref, query, nr, nq, d, k, drop_first, cand, cut, qhn,
                                                                      rmax, dist_out, idx_out, redo_list, redo_count);
  note_launch();
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
  const size_t o_cand = take((size_t)chunk * KP * 4), o_cut = take((size_t)chunk * 4);
  const size_t o_list = take((size_t)chunk * 4), o_cnt = take(256), o_rmax = take(256);
  AsyncWorkspace workspace;
  CCAL_CUDA_OK(workspace.alloc(off, stream));
  unsigned char* ws = workspace.ptr;
  __nv_bfloat16* rhi = (__nv_bfloat16*)(ws + o_rhi);
  __nv_bfloat16* rlo = (__nv_bfloat16*)(ws + o_rlo);
  float* rhn = (float*)(ws + o_rhn);
  unsigned int* rmax = (unsigned int*)(ws + o_rmax);
  cudaMemsetAsync(rmax, 0, sizeof(unsigned int), stream);
  split_rows_kernel<<<(unsigned)((nr + 7) / 8), 256, 0, stream>>>(ref, nr, d, rhi, rlo, rhn, rmax, 1);
  note_launch();
  int rc = CCAL_OK;
  for (int64_t q0 = 0; q0 < nq && rc == CCAL_OK; q0 += chunk) {
    const int64_t m = (nq - q0) < chunk ? (nq - q0) : chunk;
    int* cnt = (int*)(ws + o_cnt);
    cudaMemsetAsync(cnt, 0, sizeof(int), stream);
    float* dptr = dist_out ? dist_out + q0 * k : nullptr;
    int32_t* iptr = idx_out ? idx_out + q0 * k : nullptr;
#define CCAL_RUN_TC(KPV)                                                                                       \
  rc = run_tc<KPV>(ref, query + q0 * d, nr, m, d, k, drop_first, dptr, iptr, rhi, rlo, rhn, rmax,               \
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
