// K2 / K5: fused scoring on tcgen05 + TMEM + TMA (sm_100a).
//
//   logits = s * img @ txt^T  is NEVER written to HBM.  For each image tile (256 rows per CTA pair, 128 per CTA):
//     pass 1: stream all text tiles (256 classes x 64 features per step), accumulate 128 x 256 fp32 tiles per CTA
//             in TMEM, epilogue keeps the running row max / first argmax;
//     pass 2: stream the text tiles again, epilogue accumulates sum_j exp(cc[pred]*(l_j - l_max))
//             -> confidence = 1 / sum, then bins (confidence, correct) into a shared-memory histogram
//             (warp-aggregated atomics).
//   The image tile (128 x D per CTA) stays RESIDENT in shared memory for both passes when D <= 640, so HBM reads
//   every image feature exactly once; the text matrix is re-streamed from L2.
//
// Variants of the one templated kernel (all parity-tested, all ~99.9 % tensor-pipe active under ncu):
//   kCtas      1 = single CTAs (cta_group::1), 2 = CTA pairs (cta_group::2, each CTA loads half of every text tile)
//   kResident  image tile resident (D <= 640) or both operands streamed through the ring (D up to 1024)
//   kMode      0 = DAC scoring (ccal_score_fused), 1 = temperature-scaling loss/gradient (ccal_ts_loss_grad:
//              pass 2 also accumulates sum exp*z and picks the label logit)
//   kSplit     fp32 features as fp16 hi/lo pairs, 3 MMAs per K step (fp32-grade logits)
//   column-split work units (runtime): few rows x many classes -> every row tile is cut into class ranges and the
//              kernel runs once per pass with two tiny combine kernels between (small-batch latency mode).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (pair leader), warps 2..5 =
// epilogue (each owns 32 TMEM lanes = 32 image rows).  Producer / issuer loops are warp-uniform with one
// elect.sync lane issuing.  Pipelines: A slabs (full/empty per 64-feature slab), B ring (full/empty per stage),
// two TMEM accumulator stages (full/empty) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "ccal_common.cuh"
#include "sm100_ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <math_constants.h>
#include <stdlib.h>

namespace ccal {

constexpr int kBlockM = 128;
constexpr int kBlockN = 256;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kASlabBytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kBTileBytes = kBlockN * kBlockK * 2;   // 32 KB
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;                       // 2 accumulator stages x 256 fp32 columns
constexpr int kMaxKBlocks = 16;                      // D <= 1024
constexpr int kMaxStages = 8;
constexpr int kCtlBytes = 2048;
constexpr int kSmemLimit = 232448;                   // 227 KB opt-in maximum per CTA
constexpr float kLog2e = 1.4426950408889634f;

struct __align__(16) ScoreCtl {                      // head of dynamic shared memory
  uint64_t a_full[kMaxKBlocks];
  uint64_t a_empty[kMaxKBlocks];
  uint64_t b_full[kMaxStages];
  uint64_t b_empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad_;
  float thr[CCAL_MAX_THRESHOLDS + 1];
  BinCell cells[CCAL_MAX_THRESHOLDS + 1];
};
static_assert(sizeof(ScoreCtl) <= kCtlBytes, "control block too large");

struct ThrBlock { float t[CCAL_MAX_THRESHOLDS]; };

struct ScoreParams {
  long long n;
  int c, d;
  int kblocks, n_col_tiles, n_row_tiles, stages;
  uint32_t idesc;
  float scale;                         // logit_scale (already exponentiated)
  const float* class_conf;
  int* pred_out;
  float* conf_out;
  float* rowmax_out;
  const long long* labels;
  int n_thr;
  unsigned long long* table;
  float* row_ws;                       // kMode 1: [2*n] per-row loss / gradient terms
  const int* split_exps;               // kSplit: {e_img, e_txt}: operands were pre-scaled by 2^e before the fp16 split
  // Column-split mode (few image rows, many classes): work unit = (row tile, column range).  Launched once
  // with pass_lo = pass_hi = 0 (partial max / argmax per unit -> part_max / part_arg), then, after a tiny
  // combine, once with pass_lo = pass_hi = 1 (partial sum-exp at the known row max / pred -> part_sum).
  int n_splits, pass_lo, pass_hi;
  float* part_max;                     // [n, n_splits] raw dot-product maxima
  int* part_arg;                       // [n, n_splits]
  float* part_sum;                     // [n, n_col_tiles] per-tile partial sums (canonical order)
  const float* row_max_in;             // [n] raw dot-product row max (pass 2 of the split mode)
  const int* row_pred_in;              // [n]
  // FP8-guess pipeline (launch_guess_verify): kernel A (kFp8) multiplies its row maxima by row_scale_inv (undoing the
  // per-row / per-matrix quantisation scales); kernel B (kMode 2) verifies the guess and appends the rows whose
  // multiplier turned out different to the redo list; the redo kernel (kGather) reads its rows through that list.
  const float* row_scale_inv;          // [n]
  int* redo_count;                     // [1]
  int* redo_rows;                      // [n]
  const double* log_scale_dev;         // kMode 1: log of the logit scale read from device memory (device-side SGD loop)
  const void* gather_src;              // kGather: the image matrix itself (rows are fetched by index, not by TMA)
  int split_cap;                       // kGather: column-split only when the redo list holds at most this many rows
};


// ---- epilogue building blocks: one 32-column chunk of one TMEM lane ----------------------
// kMasked = the chunk holds fewer than 32 valid classes (last tile of a ragged vocabulary).
template <bool kMasked>
__device__ __forceinline__ void max_chunk(const uint32_t (&raw)[32], int nv, int col0, float& m, int& arg) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = (!kMasked || j < nv) ? __uint_as_float(raw[j]) : -CUDART_INF_F;
  float cm0 = fmaxf(v[0], v[1]), cm1 = fmaxf(v[2], v[3]);
#pragma unroll
  for (int j = 4; j < 32; j += 2) { cm0 = fmaxf(cm0, v[j]); cm1 = fmaxf(cm1, v[j + 1]); }
  const float cm = fmaxf(cm0, cm1);
  if (cm > m) {                                  // strict: earlier columns win ties (first max)
    int first = 31;
#pragma unroll
    for (int j = 30; j >= 0; --j) first = (v[j] == cm) ? j : first;
    m = cm;
    arg = col0 + first;
  }
}

// Returns the chunk's partial sum of 2^(a2 x - b2) (the caller adds it to the tile's running sum).
template <bool kMasked, int kMode>
__device__ __forceinline__ float exp_chunk(const uint32_t (&raw)[32], int nv, float a2, float b2, float& wsum) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, w0 = 0.f, w1 = 0.f;
  if constexpr (!kMasked && kMode == 0) {
    // full chunk, confidence only: the scale-and-shift and the four partial sums run as packed fp32 pairs
    // (FFMA2 / FADD2) - the same IEEE operations in the same order as the scalar code below, two per instruction
    const unsigned long long aa = ptx::pack2(a2, a2), bb = ptx::pack2(-b2, -b2);
    unsigned long long s01 = 0, s23 = 0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float x0, x1, x2, x3;
      ptx::unpack2(ptx::fma2(ptx::pack2(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1])), aa, bb), x0, x1);
      ptx::unpack2(ptx::fma2(ptx::pack2(__uint_as_float(raw[j + 2]), __uint_as_float(raw[j + 3])), aa, bb), x2, x3);
      const unsigned long long e01 = ptx::pack2(ptx::ex2_approx(x0), ptx::ex2_approx(x1));
      const unsigned long long e23 = ptx::pack2(ptx::ex2_approx(x2), ptx::ex2_approx(x3));
      s01 = j == 0 ? e01 : ptx::add2(s01, e01);          // 0 + e == e exactly (e >= +0)
      s23 = j == 0 ? e23 : ptx::add2(s23, e23);
    }
    ptx::unpack2(s01, s0, s1);
    ptx::unpack2(s23, s2, s3);
    return (s0 + s1) + (s2 + s3);
  }
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float e[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      e[u] = ptx::ex2_approx(fmaf(__uint_as_float(raw[j + u]), a2, -b2));
      if (kMasked && j + u >= nv) e[u] = 0.f;
    }
    s0 += e[0]; s1 += e[1]; s2 += e[2]; s3 += e[3];
    if (kMode == 1) {
      w0 = fmaf(e[0], __uint_as_float(raw[j + 0]), w0); w1 = fmaf(e[1], __uint_as_float(raw[j + 1]), w1);
      w0 = fmaf(e[2], __uint_as_float(raw[j + 2]), w0); w1 = fmaf(e[3], __uint_as_float(raw[j + 3]), w1);
    }
  }
  // chunk partial first, then into the running total: two-level summation keeps the
  // rounding error of a 49k-term sum at the ~1e-6 level
  if (kMode == 1) wsum += w0 + w1;
  return (s0 + s1) + (s2 + s3);
}

// Verify pass: a chunk can hold the row maximum only if one of its terms 2^(a2 (x - m)) reaches 1 (m = the guessed
// class's exact logit, itself one of the row's logits) - so only chunks whose partial sum is at least this close
// to 1 are searched for the exact maximum / first argmax.  Skipped chunks hold logits strictly below m (by more
// than 0.01 / a2, far beyond any rounding), and the guessed class's own chunk always qualifies.
constexpr float kMaxSearchThreshold = 0.99f;

// kCtas = 1: one CTA per 128-row image tile (tcgen05 cta_group::1).
// kCtas = 2: a CTA PAIR (2-CTA cluster, cta_group::2) per 256-row tile: each CTA keeps its own 128 image
//            rows resident and loads only HALF of every text tile (128 classes); the pair's leader issues
//            256x256x16 MMAs that read both halves.  Halves the L2->SMEM text traffic and the SMEM->tensor
//            B traffic per flop - on a power-capped B200 that is what buys throughput.
// kSplit (fp32 features, streaming only): every operand arrives as an fp16 pair x*2^e = hi + lo and each K step
//            issues hi.hi + hi.lo + lo.hi (the dropped lo.lo term is < 2^-22 |a||b|): fp32-grade logits from
//            16-bit tensor-core operands at 3x the MMA work.
// In-kernel timing of the scoring launches (bench.py's roofline): every CTA stamps %globaltimer and clock64 when it
// starts and when it leaves; per kernel kind the library keeps the number of launches, the summed launch spans
// (first CTA in -> last CTA out), and the CTAs' summed busy nanoseconds and SM cycles, whose ratio is the mean SM
// clock the kernel actually ran at (nvidia-smi samples every 100 ms and cannot see inside a launch).
// Launches of one kind are assumed not to overlap each other (they are issued on one stream).
struct KernelTrace {
  unsigned long long start_ns, end_ns, done, launches, span_ns, busy_ns, cycles, pad_;
};
constexpr int kTraceKinds = 5;          // 0 guess (fp8 pass 1), 1 verify (bf16 pass 2 + exact max), 2 redo, 3 two-pass, 4 temperature scaling
__device__ KernelTrace g_trace[kTraceKinds] = {{~0ull, 0, 0, 0, 0, 0, 0, 0}, {~0ull, 0, 0, 0, 0, 0, 0, 0}, {~0ull, 0, 0, 0, 0, 0, 0, 0},
                                               {~0ull, 0, 0, 0, 0, 0, 0, 0}, {~0ull, 0, 0, 0, 0, 0, 0, 0}};

__device__ __forceinline__ void trace_exit(int kind, unsigned long long t0, long long c0) {
  KernelTrace& tr = g_trace[kind];
  const unsigned long long t1 = ptx::globaltimer_ns();
  const long long c1 = clock64();
  atomicMin(&tr.start_ns, t0);
  atomicMax(&tr.end_ns, t1);
  atomicAdd(&tr.busy_ns, t1 - t0);
  atomicAdd(&tr.cycles, (unsigned long long)(c1 - c0));
  __threadfence();
  if (atomicAdd(&tr.done, 1ull) + 1ull == (unsigned long long)gridDim.x) {      // last CTA of this launch
    __threadfence();
    const unsigned long long e = atomicExch(&tr.end_ns, 0ull), b = atomicExch(&tr.start_ns, ~0ull);
    atomicAdd(&tr.span_ns, e - b);
    atomicAdd(&tr.launches, 1ull);
    atomicExch(&tr.done, 0ull);
  }
}

// Class ranges per row tile of the redo (gather) kernel, derived on the device from the length of the redo list:
// the S that minimises a two-term cost model (below), every range at least 4 text tiles long.  The finish
// kernel evaluates the same function, so both agree without a host round trip.
__host__ __device__ inline int gather_splits(int n_listed, int tile_rows, int n_units, int n_col_tiles, int split_cap) {
  const int n_row_tiles = (n_listed + tile_rows - 1) / tile_rows;
  int S = 1;
  if (n_row_tiles > 0 && n_listed <= split_cap) {
    // time of the launch in text-tile units: waves x (text tiles per unit + the unit's fixed cost).  The fixed cost is
    // the by-index gather of the unit's image rows by one warp - measured at ~100 us, the MMA time of ~32 text tiles
    // at d = 512 - so cutting tiles into class ranges only pays when the row tiles leave most SMs without work
    // (62 row tiles on 74 pairs: S = 1 takes 0.63 ms where the wave-efficiency rule's S = 7 took 1.27 ms).
    constexpr int kUnitCost = 32;
    const int s_max = n_col_tiles / 4 < 32 ? (n_col_tiles / 4 < 1 ? 1 : n_col_tiles / 4) : 32;
    long long best = -1;
    for (int s = 1; s <= s_max; ++s) {
      const long long units = (long long)n_row_tiles * s, waves = (units + n_units - 1) / n_units;
      const long long cost = waves * ((n_col_tiles + s - 1) / s + kUnitCost);
      if (best < 0 || cost < best) { best = cost; S = s; }
    }
  }
  return S;
}

// kFp8 (pass 1 only): e4m3 operands, kind::f8f6f4 - the same bytes per stage carry 128 features instead of 64.
// kMode 2 (pass 2 only, "verify"): pass-2 sums at the multiplier of a GUESSED class while the exact bf16 row
//            maximum / first argmax are tracked alongside; rows whose exact argmax has a different multiplier go to
//            the redo list, everything else is final.
// kGather (pass 2 only, resident): work = the redo list; the image rows are gathered by index into the swizzled
//            slab layout by the producer warp (generic stores + proxy fence) instead of TMA, and the work
//            decomposition (row tiles x class ranges) is derived on the device from the list length.
template <int kCtas, bool kResident, int kMode, bool kSplit, bool kFp8 = false, bool kGather = false>
__global__ void __launch_bounds__(kThreads, 1)
score_fused_kernel(const __grid_constant__ CUtensorMap map_img, const __grid_constant__ CUtensorMap map_txt,
                   const __grid_constant__ CUtensorMap map_img_lo, const __grid_constant__ CUtensorMap map_txt_lo,
                   const __grid_constant__ ScoreParams p, const __grid_constant__ ThrBlock thr) {
  static_assert(!(kSplit && kResident), "the split-precision variant streams both operands");
  static_assert(!kFp8 || (kMode == 0 && !kSplit && kResident), "the FP8 guess pass is a resident pass-1-only variant");
  static_assert(!kGather || (kMode == 0 && !kSplit && kResident && !kFp8), "the gather variant is resident, pass 2 only");
  constexpr int kSlabElems = kFp8 ? 128 : kBlockK;       // features per 128-byte slab row
  extern __shared__ unsigned char smem_dyn[];
  ScoreCtl* ctl = reinterpret_cast<ScoreCtl*>(smem_dyn);
  // operand area starts at the next 1024-byte boundary (128B-swizzle atoms are 1024 B)
  const uint32_t ctl_end = ptx::smem_u32(smem_dyn) + kCtlBytes;
  const uint32_t op_base = (ctl_end + 1023u) & ~1023u;
  unsigned char* op_ptr = smem_dyn + (op_base - ptx::smem_u32(smem_dyn));
  constexpr int kBBytes = kBTileBytes / kCtas;           // text bytes this CTA loads per stage
  constexpr int kBRows = kBlockN / kCtas;                // classes this CTA loads per text tile
  constexpr int kTileRows = kBlockM * kCtas;             // image rows per (pair) tile
  // resident: [A slab 0..kblocks) then ring of B tiles; streaming: ring of {A slab, B tile}
  constexpr int kParts = kSplit ? 2 : 1;                 // hi (+ lo) copies of each operand per stage
  const uint32_t stage_bytes = kResident ? kBBytes : kParts * (kASlabBytes + kBBytes);
  unsigned char* ring_ptr = op_ptr + (kResident ? p.kblocks * kASlabBytes : 0);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (kCtas == 2) ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int unit = (kCtas == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_units = (int)gridDim.x / kCtas;

  // ------------------------------------------------------------------ one-time setup
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_img);
    ptx::prefetch_tensormap(&map_txt);
    // gather variant: one arrival per CTA of the pair (each producer warp announces its own half of the slab)
    for (int i = 0; i < kMaxKBlocks; ++i) { ptx::mbar_init(&ctl->a_full[i], kGather ? kCtas : 1); ptx::mbar_init(&ctl->a_empty[i], 1); }
    for (int i = 0; i < kMaxStages; ++i) { ptx::mbar_init(&ctl->b_full[i], 1); ptx::mbar_init(&ctl->b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&ctl->tmem_full[i], 1); ptx::mbar_init(&ctl->tmem_empty[i], 4 * kCtas); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    if (kCtas == 2) { ptx::tmem_alloc_2sm(&ctl->tmem_base, kTmemCols); ptx::tmem_relinquish_2sm(); }
    else { ptx::tmem_alloc(&ctl->tmem_base, kTmemCols); ptx::tmem_relinquish(); }
  }
  for (int i = threadIdx.x; i <= CCAL_MAX_THRESHOLDS; i += kThreads) {
    ctl->cells[i] = BinCell{0u, 0u, 0ull};
    ctl->thr[i] = (i < p.n_thr) ? thr.t[i] : CUDART_INF_F;
  }
  ptx::tc_fence_before();
  if (kCtas == 2) ptx::cluster_sync_all(); else __syncthreads();   // peer barriers initialised before any remote signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  unsigned long long trace_t0 = 0;
  long long trace_c0 = 0;
  if (threadIdx.x == 0) { trace_t0 = ptx::globaltimer_ns(); trace_c0 = clock64(); }

  const int NT = p.n_col_tiles;
  const int KB = p.kblocks;
  // work units = (row tile, class range).  Normally fixed by the host; the gather variant derives them from the
  // length of the redo list, which only exists on the device: few listed rows -> cut every row tile into S class
  // ranges so that the units fill the SMs (S is the smallest count within 2 % of the best wave efficiency).
  int S = p.n_splits;                             // class ranges per row tile (1 = the normal mode)
  int n_row_tiles = p.n_row_tiles;
  int n_listed = 0;
  if (kGather) {
    n_listed = *reinterpret_cast<const volatile int*>(p.redo_count);
    n_row_tiles = (n_listed + kTileRows - 1) / kTileRows;
    S = gather_splits(n_listed, kTileRows, n_units, NT, p.split_cap);
  }
  const int n_work = n_row_tiles * S;

  if (warp == 0) {
    // ================================================================ TMA producer (whole warp loops, one
    // elected lane issues; keeping the control flow warp-uniform keeps the issue path short)
    uint32_t stage = 0, phase = 0, ti = 0;
    for (int u = unit; u < n_work; u += n_units, ++ti) {
      const int tile = u / S, nt_begin = (u % S) * NT / S, nt_end = ((u % S) + 1) * NT / S;
      const int row0 = tile * kTileRows + (int)rank * kBlockM;
      for (int pass = p.pass_lo; pass <= p.pass_hi; ++pass) {
        for (int nt = nt_begin; nt < nt_end; ++nt) {
          const int col0 = nt * kBlockN + (int)rank * kBRows;
          const bool load_a = kResident && pass == p.pass_lo && nt == nt_begin;
          for (int kb = 0; kb < KB; ++kb) {
            unsigned char* sp = ring_ptr + (size_t)stage * stage_bytes;
            // slab kb of the previous row tile must have been consumed by its last MMA
            if (load_a) ptx::mbar_wait(&ctl->a_empty[kb], (ti & 1u) ^ 1u);
            if (kGather && load_a) {
              // 128 listed rows x 128 bytes of slab kb: lane = row (4 rounds of 32 rows), 8 x 16-byte chunks each,
              // stored where TMA's 128-byte swizzle would have put them (chunk index XOR row % 8)
              unsigned char* slab = op_ptr + (size_t)kb * kASlabBytes;
              const unsigned char* src = reinterpret_cast<const unsigned char*>(p.gather_src);
              if (kb == 0) {
                // pull the rows of this CTA's NEXT unit towards L2 while the current one is being scored
                const int un = u + n_units;
                if (un < n_work) {
                  const int next_row0 = (un / S) * kTileRows + (int)rank * kBlockM;
                  for (int r = lane; r < kBlockM; r += 32) {
                    const int ns = next_row0 + r;
                    if (ns < n_listed) {
                      const unsigned char* gp = src + (size_t)p.redo_rows[ns] * (size_t)p.d * 2;
                      for (int off = 0; off < p.d * 2; off += 128) ptx::prefetch_l2(gp + off);
                    }
                  }
                }
              }
              uint4 v[kBlockM / 32][8];                     // the whole slab's loads in flight before the first store
#pragma unroll
              for (int rr = 0; rr < kBlockM / 32; ++rr) {
                const int slot = row0 + rr * 32 + lane;
                if (slot < n_listed) {
                  const uint4* g = reinterpret_cast<const uint4*>(src + (size_t)p.redo_rows[slot] * (size_t)p.d * 2 + (size_t)kb * 128);
#pragma unroll
                  for (int q = 0; q < 8; ++q) v[rr][q] = __ldg(g + q);
                } else {
#pragma unroll
                  for (int q = 0; q < 8; ++q) v[rr][q] = make_uint4(0u, 0u, 0u, 0u);
                }
              }
#pragma unroll
              for (int rr = 0; rr < kBlockM / 32; ++rr) {
                const int r = rr * 32 + lane;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  *reinterpret_cast<uint4*>(slab + r * 128 + ((q ^ (r & 7)) << 4)) = v[rr][q];
              }
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_release_cluster(&ctl->a_full[kb], 0u);
            }
            ptx::mbar_wait(&ctl->b_empty[stage], phase ^ 1u);
            if (ptx::elect_one()) {
              if (kResident) {
                if (load_a && !kGather) {
                  if (leader) ptx::mbar_arrive_expect_tx(&ctl->a_full[kb], kASlabBytes * kCtas);
                  if (kCtas == 2) ptx::tma_load_2d_2sm(op_ptr + (size_t)kb * kASlabBytes, &map_img, &ctl->a_full[kb], kb * kSlabElems, row0, ptx::kEvictFirst);
                  else ptx::tma_load_2d(op_ptr + (size_t)kb * kASlabBytes, &map_img, &ctl->a_full[kb], kb * kSlabElems, row0, ptx::kEvictFirst);
                }
                if (leader) ptx::mbar_arrive_expect_tx(&ctl->b_full[stage], kBBytes * kCtas);
                if (kCtas == 2) ptx::tma_load_2d_2sm(sp, &map_txt, &ctl->b_full[stage], kb * kSlabElems, col0, ptx::kEvictLast);
                else ptx::tma_load_2d(sp, &map_txt, &ctl->b_full[stage], kb * kSlabElems, col0, ptx::kEvictLast);
              } else {
                // stage layout: [A hi][A lo?][B hi][B lo?]
                unsigned char* sb = sp + kParts * kASlabBytes;
                if (leader) ptx::mbar_arrive_expect_tx(&ctl->b_full[stage], kParts * (kASlabBytes + kBBytes) * kCtas);
                if (kCtas == 2) {
                  ptx::tma_load_2d_2sm(sp, &map_img, &ctl->b_full[stage], kb * kBlockK, row0, ptx::kEvictNormal);
                  ptx::tma_load_2d_2sm(sb, &map_txt, &ctl->b_full[stage], kb * kBlockK, col0, ptx::kEvictLast);
                  if (kSplit) {
                    ptx::tma_load_2d_2sm(sp + kASlabBytes, &map_img_lo, &ctl->b_full[stage], kb * kBlockK, row0, ptx::kEvictNormal);
                    ptx::tma_load_2d_2sm(sb + kBBytes, &map_txt_lo, &ctl->b_full[stage], kb * kBlockK, col0, ptx::kEvictLast);
                  }
                } else {
                  ptx::tma_load_2d(sp, &map_img, &ctl->b_full[stage], kb * kBlockK, row0, ptx::kEvictNormal);
                  ptx::tma_load_2d(sb, &map_txt, &ctl->b_full[stage], kb * kBlockK, col0, ptx::kEvictLast);
                  if (kSplit) {
                    ptx::tma_load_2d(sp + kASlabBytes, &map_img_lo, &ctl->b_full[stage], kb * kBlockK, row0, ptx::kEvictNormal);
                    ptx::tma_load_2d(sb + kBBytes, &map_txt_lo, &ctl->b_full[stage], kb * kBlockK, col0, ptx::kEvictLast);
                  }
                }
              }
            }
            __syncwarp();
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (pair leader only; whole warp
    // waits, one elected lane issues: ~4 MMAs + 1-2 commits per 64-feature block must fit in 512 cycles)
    if (leader) {
      uint32_t stage = 0, phase = 0, acc_it = 0, ti = 0;
      const uint32_t ring_base = op_base + (kResident ? (uint32_t)KB * kASlabBytes : 0u);
      for (int u = unit; u < n_work; u += n_units, ++ti) {
        const int nt_begin = (u % S) * NT / S, nt_end = ((u % S) + 1) * NT / S;
        for (int pass = p.pass_lo; pass <= p.pass_hi; ++pass) {
          for (int nt = nt_begin; nt < nt_end; ++nt, ++acc_it) {
            const uint32_t as = acc_it & 1u;
            const uint32_t aph = (acc_it >> 1) & 1u;
            ptx::mbar_wait_short(&ctl->tmem_empty[as], aph ^ 1u);      // epilogue(s) have drained this stage
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * kBlockN;
            const bool first_use_of_a = kResident && pass == p.pass_lo && nt == nt_begin;
            const bool last_use_of_a = kResident && pass == p.pass_hi && nt == nt_end - 1;
            for (int kb = 0; kb < KB; ++kb) {
              const uint32_t sp = ring_base + stage * stage_bytes;
              if (first_use_of_a) {
                // gathered slabs were written with generic stores by BOTH CTAs' producer warps: acquire at cluster scope
                if (kGather) ptx::mbar_wait_cluster(&ctl->a_full[kb], ti & 1u); else ptx::mbar_wait(&ctl->a_full[kb], ti & 1u);
              }
              ptx::mbar_wait(&ctl->b_full[stage], phase);
              ptx::tc_fence_after();
              if (ptx::elect_one()) {
                const uint32_t a_addr = kResident ? op_base + (uint32_t)kb * kASlabBytes : sp;
                const uint32_t b_addr = kResident ? sp : sp + kParts * kASlabBytes;
                const uint64_t a_desc = ptx::make_kmajor_sw128_desc(a_addr);
                const uint64_t b_desc = ptx::make_kmajor_sw128_desc(b_addr);
                const uint64_t a_lo = kSplit ? ptx::make_kmajor_sw128_desc(a_addr + kASlabBytes) : 0;
                const uint64_t b_lo = kSplit ? ptx::make_kmajor_sw128_desc(b_addr + kBBytes) : 0;
                auto mma = [&](uint64_t ad, uint64_t bd, uint32_t acc) {
                  if (kFp8) { if (kCtas == 2) ptx::umma_f8_2sm(d_tmem, ad, bd, p.idesc, acc); else ptx::umma_f8(d_tmem, ad, bd, p.idesc, acc); }
                  else if (kCtas == 2) ptx::umma_f16_2sm(d_tmem, ad, bd, p.idesc, acc);
                  else ptx::umma_f16(d_tmem, ad, bd, p.idesc, acc);
                };
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                  // +32 bytes per 16-element K step inside the 128 B swizzle span (encoded >> 4)
                  const uint64_t o = (uint64_t)(k * 2);
                  mma(a_desc + o, b_desc + o, (uint32_t)((kb | k) != 0));
                  if (kSplit) { mma(a_desc + o, b_lo + o, 1u); mma(a_lo + o, b_desc + o, 1u); }
                }
                // ring slot (in both CTAs) reusable when these MMAs retire
                if (kCtas == 2) ptx::umma_commit_2sm(&ctl->b_empty[stage]); else ptx::umma_commit(&ctl->b_empty[stage]);
                if (last_use_of_a) {
                  if (kCtas == 2) ptx::umma_commit_2sm(&ctl->a_empty[kb]); else ptx::umma_commit(&ctl->a_empty[kb]);
                }
                // accumulator tile complete (each CTA's epilogue reads its own 128 rows)
                if (kb == KB - 1) {
                  if (kCtas == 2) ptx::umma_commit_2sm(&ctl->tmem_full[as]); else ptx::umma_commit(&ctl->tmem_full[as]);
                }
              }
              __syncwarp();
              if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else {
    // ================================================================ epilogue (warps 2..5)
    const int quarter = warp & 3;                                  // TMEM lanes [32q, 32q+32)
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    uint32_t acc_it = 0;
    // split operands were pre-scaled by powers of two: fold 2^-(e_img + e_txt) into the logit scale (exact)
    const float scale_in = (kMode == 1 && p.log_scale_dev != nullptr) ? expf((float)*p.log_scale_dev) : p.scale;
    const float scale = kSplit ? scale_in * exp2f(-(float)(p.split_exps[0] + p.split_exps[1])) : scale_in;
    auto release_acc = [&](uint32_t as) {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCtas == 2) ptx::mbar_arrive_cluster(&ctl->tmem_empty[as], 0u);   // the leader's MMA warp waits for both CTAs
        else ptx::mbar_arrive(&ctl->tmem_empty[as]);
      }
    };
    for (int u = unit; u < n_work; u += n_units) {
      const int tile = u / S, split = u % S, nt_begin = split * NT / S, nt_end = (split + 1) * NT / S;
      // slot = position of this thread's row in the work list; row = the image it stands for (the same thing
      // unless the rows come from the redo list)
      const long long slot = (long long)tile * kTileRows + (long long)rank * kBlockM + quarter * 32 + lane;
      const bool row_ok = kGather ? slot < (long long)n_listed : slot < p.n;
      const long long row = kGather ? (row_ok ? (long long)p.redo_rows[slot] : 0ll) : slot;
      // ---------------- pass 1: running max / first argmax of the raw dot products
      float m = -CUDART_INF_F;
      int arg = 0;
      if (kMode != 2 && !kGather && p.pass_lo == 0)
      for (int nt = nt_begin; nt < nt_end; ++nt, ++acc_it) {
        const uint32_t as = acc_it & 1u;
        ptx::mbar_wait(&ctl->tmem_full[as], (acc_it >> 1) & 1u);
        ptx::tc_fence_after();
        const int valid = min(kBlockN, p.c - nt * kBlockN);
        if (valid == kBlockN) {
          // full tile (all but the last one): straight-line code, no per-chunk loop or mask bookkeeping - every
          // instruction the epilogue does not issue is power the tensor pipe keeps
#pragma unroll
          for (int ch = 0; ch < kBlockN / 32; ++ch) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tmem_base + lane_sel + as * kBlockN + ch * 32, raw);
            ptx::tmem_ld_wait(raw);
            max_chunk<false>(raw, 32, nt * kBlockN + ch * 32, m, arg);
          }
        } else {
          for (int ch = 0; ch * 32 < valid; ++ch) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tmem_base + lane_sel + as * kBlockN + ch * 32, raw);
            ptx::tmem_ld_wait(raw);
            const int nv = valid - ch * 32;
            if (nv >= 32) max_chunk<false>(raw, 32, nt * kBlockN + ch * 32, m, arg);
            else max_chunk<true>(raw, nv, nt * kBlockN + ch * 32, m, arg);
          }
        }
        release_acc(as);
      }
      if (kFp8 || p.pass_hi == 0) {               // pass 1 only: partial (max, argmax) per class range
        if (row_ok) {
          // the FP8 guess pass hands its maxima back in the units of the un-quantised dot product
          p.part_max[row * S + split] = (kFp8 && p.row_scale_inv != nullptr) ? m * p.row_scale_inv[row] : m;
          p.part_arg[row * S + split] = arg;
        }
        continue;
      }
      if constexpr (!kFp8) {
      if ((kMode == 2 || kGather || p.pass_lo == 1) && row_ok) { m = p.row_max_in[row]; arg = p.row_pred_in[row]; }
      // ---------------- pass 2: sum of exp at the predicted class's multiplier
      float cc = 1.0f;
      if (kMode != 1 && p.class_conf != nullptr) cc = __ldg(p.class_conf + arg);
      const float a2 = cc * scale * kLog2e;
      const float b2 = m * a2;
      float sum = 0.f, wsum = 0.f, zy = 0.f;
      // kMode 2: (m, arg) above are the GUESS (approximate maximum, likely argmax); the exact ones are tracked here
      float xm = -CUDART_INF_F;
      int xarg = 0;
      const int label = (kMode == 1 && row_ok) ? (int)p.labels[row] : -1;
      for (int nt = nt_begin; nt < nt_end; ++nt, ++acc_it) {
        const uint32_t as = acc_it & 1u;
        ptx::mbar_wait(&ctl->tmem_full[as], (acc_it >> 1) & 1u);
        ptx::tc_fence_after();
        const int valid = min(kBlockN, p.c - nt * kBlockN);
        // canonical summation order: 32-class partials -> one partial per 256-class tile -> row total over the
        // tiles in class order.  The column-split mode stores the per-tile partials and its finish kernel adds
        // them in the same order, so confidences are bit-identical however the work was cut.
        float tile_sum = 0.f;
        if (kMode != 1 && valid == kBlockN) {
#pragma unroll
          for (int ch = 0; ch < kBlockN / 32; ++ch) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tmem_base + lane_sel + as * kBlockN + ch * 32, raw);
            ptx::tmem_ld_wait(raw);
            const float part = exp_chunk<false, 0>(raw, 32, a2, b2, wsum);
            tile_sum += part;
            if (kMode == 2 && !(part < kMaxSearchThreshold)) max_chunk<false>(raw, 32, nt * kBlockN + ch * 32, xm, xarg);
          }
        } else
        for (int ch = 0; ch * 32 < valid; ++ch) {
          uint32_t raw[32];
          ptx::tmem_ld_32x32(tmem_base + lane_sel + as * kBlockN + ch * 32, raw);
          ptx::tmem_ld_wait(raw);
          const int nv = valid - ch * 32;
          const float part = (nv >= 32) ? exp_chunk<false, kMode == 1 ? 1 : 0>(raw, 32, a2, b2, wsum)
                                        : exp_chunk<true, kMode == 1 ? 1 : 0>(raw, nv, a2, b2, wsum);
          tile_sum += part;
          if (kMode == 2 && !(part < kMaxSearchThreshold)) {
            if (nv >= 32) max_chunk<false>(raw, 32, nt * kBlockN + ch * 32, xm, xarg);
            else max_chunk<true>(raw, nv, nt * kBlockN + ch * 32, xm, xarg);
          }
          if (kMode == 1) {
            const int rel = label - (nt * kBlockN + ch * 32);
            if (rel >= 0 && rel < 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) zy = (j == rel) ? __uint_as_float(raw[j]) : zy;
            }
          }
        }
        sum += tile_sum;
        if (S > 1 && row_ok) p.part_sum[slot * NT + nt] = tile_sum;
        release_acc(as);
      }
      if (S > 1) continue;                        // column-split mode, second launch: per-tile partials are out
      // ---------------- per-row results
      if (kMode == 2) {
        // The sum was taken at the guessed class's multiplier, shifted by that class's EXACT bf16 dot product m
        // (guess_logit_kernel).  When the guess is right, m is the row maximum and every term is the very fma / ex2
        // the two-pass kernel evaluates, so 1 / sum is its confidence bit for bit.  The row is final when the
        // tracked exact maximum equals m and the first argmax carries the same multiplier (it normally IS the
        // guessed class; an equal-valued earlier class with an equal multiplier changes nothing); otherwise it is
        // listed for the redo kernel, which repeats pass 2 from the exact (xm, xarg).
        float cc_exact = 1.0f;
        if (p.class_conf != nullptr && row_ok) cc_exact = __ldg(p.class_conf + xarg);
        const bool accept = row_ok && xm == m && cc_exact == cc;
        const float conf = 1.0f / sum;
        if (row_ok) {
          if (p.pred_out) p.pred_out[row] = xarg;
          if (accept) {
            if (p.conf_out) p.conf_out[row] = conf;
            if (p.rowmax_out) p.rowmax_out[row] = xm * scale;
          } else {
            p.part_max[row] = xm;
            p.part_arg[row] = xarg;
            p.redo_rows[atomicAdd(p.redo_count, 1)] = (int)row;
          }
        }
        if (p.table != nullptr) {
          const bool correct = accept && ((long long)xarg == p.labels[row]);
          warp_bin_add(ctl->cells, bin_of(conf, ctl->thr, p.n_thr), correct, conf_to_fx(conf), accept);
        }
      } else if (kMode == 0) {
        const float conf = 1.0f / sum;
        if (row_ok) {
          if (p.pred_out) p.pred_out[row] = arg;
          if (p.conf_out) p.conf_out[row] = conf;
          if (p.rowmax_out) p.rowmax_out[row] = m * scale;
        }
        if (p.table != nullptr) {
          const bool correct = row_ok && ((long long)arg == p.labels[row]);
          warp_bin_add(ctl->cells, bin_of(conf, ctl->thr, p.n_thr), correct, conf_to_fx(conf), row_ok);
        }
      } else if (row_ok) {
        // loss_i = logsumexp_j(s z_j) - s z_y ;  d loss_i / dt = s (sum_j p_j z_j - z_y), s = exp(t)
        p.row_ws[row] = logf(sum) + scale * (m - zy);
        p.row_ws[p.n + row] = scale * (wsum / sum - zy);
      }
      }
    }
  }

  // ------------------------------------------------------------------ teardown
  __syncwarp();                                                    // single-lane roles rejoin their warp
  ptx::tc_fence_before();
  if (kCtas == 2) ptx::cluster_sync_all(); else __syncthreads();   // the peer may still read my SMEM / TMEM until here
  ptx::tc_fence_after();
  if (warp == 1) {
    if (kCtas == 2) ptx::tmem_dealloc_2sm(tmem_base, kTmemCols); else ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
  if (threadIdx.x == 0) trace_exit(kFp8 ? 0 : kMode == 2 ? 1 : kGather ? 2 : kMode == 0 ? 3 : 4, trace_t0, trace_c0);
  if (kMode != 1 && p.table != nullptr) {
    for (int i = threadIdx.x; i <= p.n_thr; i += kThreads) {
      const BinCell cell = ctl->cells[i];
      if (cell.count) {
        atomicAdd(&p.table[3 * i + 0], (unsigned long long)cell.count);
        atomicAdd(&p.table[3 * i + 1], (unsigned long long)cell.correct);
        atomicAdd(&p.table[3 * i + 2], cell.sum_fx);
      }
    }
  }
}

// deterministic fixed-order reduction of the per-row loss / gradient terms (kMode 1)
__global__ void __launch_bounds__(1024)
ts_reduce_kernel(const float* __restrict__ row_ws, long long n, double* __restrict__ out2) {
  __shared__ double s_loss[32], s_grad[32];
  double loss = 0.0, grad = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    loss += (double)row_ws[i];
    grad += (double)row_ws[n + i];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, off);
    grad += __shfl_xor_sync(0xffffffffu, grad, off);
  }
  if ((threadIdx.x & 31) == 0) { s_loss[threadIdx.x >> 5] = loss; s_grad[threadIdx.x >> 5] = grad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0, g = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { l += s_loss[w]; g += s_grad[w]; }
    out2[0] = l / (double)n;
    out2[1] = g / (double)n;
  }
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// [rows, d] row-major matrix, box = {one 128-byte slab row, box_rows}, 128-byte swizzle, zero OOB fill.
// 16-bit dtypes: 64 features per slab row; kE4M3 (internal, the FP8 guess operands): 128.
constexpr int kE4M3 = 3;
int make_map(CUtensorMap* map, const void* base, long long rows, int d, int box_rows, int dtype) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(CCAL_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  const int esize = dtype == kE4M3 ? 1 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d * esize};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == kE4M3 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                               : dtype == CCAL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CCAL_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return CCAL_OK;
}

// ---- column-split mode glue (few rows, many classes) -----------------------------------------------
// combine the per-range (max, first argmax): ranges are in increasing class order, so on equal maxima the
// lower range wins (first-max semantics)
__global__ void __launch_bounds__(256)
split_combine_max_kernel(const float* __restrict__ part_max, const int* __restrict__ part_arg, long long n, int S,
                         float* __restrict__ row_max, int* __restrict__ row_pred) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  float m = part_max[row * S];
  int a = part_arg[row * S];
  for (int s = 1; s < S; ++s) {
    const float v = part_max[row * S + s];
    if (v > m) { m = v; a = part_arg[row * S + s]; }
  }
  row_max[row] = m;
  row_pred[row] = a;
}

// add the per-tile partial sums in class order (the unsplit kernel's order), emit (pred, conf, rowmax) and bin
__global__ void __launch_bounds__(256)
split_finish_kernel(const float* __restrict__ part_sum, const float* __restrict__ row_max, const int* __restrict__ row_pred,
                    long long n, int n_tiles, float scale, const int* __restrict__ split_exps, int* __restrict__ pred_out,
                    float* __restrict__ conf_out, float* __restrict__ rowmax_out, const long long* __restrict__ labels,
                    const __grid_constant__ ThrBlock thr, int n_thr, unsigned long long* __restrict__ table,
                    const int* __restrict__ redo_rows = nullptr, const int* __restrict__ redo_count = nullptr,
                    int tile_rows = 0, int n_units = 0, int split_cap = 0) {
  __shared__ BinCell cells[CCAL_MAX_THRESHOLDS + 1];
  __shared__ float s_thr[CCAL_MAX_THRESHOLDS + 1];
  if (redo_rows != nullptr) {
    // redo list: rows are named by the list, and there is only something to finish when the gather kernel
    // decided (by the same function) to cut its row tiles into class ranges
    n = *redo_count;
    if (gather_splits((int)n, tile_rows, n_units, n_tiles, split_cap) == 1) return;
  }
  for (int i = threadIdx.x; i <= CCAL_MAX_THRESHOLDS; i += blockDim.x) {
    cells[i] = BinCell{0u, 0u, 0ull};
    s_thr[i] = i < n_thr ? thr.t[i] : CUDART_INF_F;
  }
  __syncthreads();
  if (split_exps) scale *= exp2f(-(float)(split_exps[0] + split_exps[1]));
  const long long n_round = ((n + 31) / 32) * 32;
  for (long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x; slot < n_round;
       slot += (long long)gridDim.x * blockDim.x) {
    const bool ok = slot < n;
    const long long row = (redo_rows != nullptr && ok) ? (long long)redo_rows[slot] : slot;
    float conf = 1.f;
    int pred = 0;
    if (ok) {
      float sum = 0.f;
      for (int t = 0; t < n_tiles; ++t) sum += part_sum[slot * n_tiles + t];     // same order as the unsplit kernel
      conf = 1.0f / sum;
      pred = row_pred[row];
      if (pred_out) pred_out[row] = pred;
      if (conf_out) conf_out[row] = conf;
      if (rowmax_out) rowmax_out[row] = row_max[row] * scale;
    }
    if (table != nullptr) {
      const bool correct = ok && ((long long)pred == labels[row]);
      warp_bin_add(cells, bin_of(conf, s_thr, n_thr), correct, conf_to_fx(conf), ok);
    }
  }
  __syncthreads();
  if (table != nullptr)
    for (int i = threadIdx.x; i <= n_thr; i += blockDim.x) {
      const BinCell cell = cells[i];
      if (cell.count) {
        atomicAdd(&table[3 * i + 0], (unsigned long long)cell.count);
        atomicAdd(&table[3 * i + 1], (unsigned long long)cell.correct);
        atomicAdd(&table[3 * i + 2], cell.sum_fx);
      }
    }
}

template <int kCtas, bool kResident, int kMode, bool kSplit, bool kFp8 = false, bool kGather = false>
static int launch_variant(const CUtensorMap& mi, const CUtensorMap& mt, const CUtensorMap& mi_lo, const CUtensorMap& mt_lo,
                          const ScoreParams& p, const ThrBlock& thr, int grid, size_t smem, cudaStream_t stream) {
  auto kern = score_fused_kernel<kCtas, kResident, kMode, kSplit, kFp8, kGather>;
  CCAL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (kCtas == 2) ? 1 : 0;
  CCAL_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, mi, mt, mi_lo, mt_lo, p, thr));
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

// CTAs per image tile: 2 (CTA pair, cta_group::2) unless CCAL_SCORE_CTAS=1 is set or the shard is too small
// to give every pair a tile.
static int choose_ctas(int64_t n) {
  const char* e = getenv("CCAL_SCORE_CTAS");
  if (e && e[0] == '1') return 1;
  if (e && e[0] == '2') return 2;
  return n >= (int64_t)kBlockM * 2 * (num_sms() / 2) ? 2 : 1;
}

// ---- FP8 guess -> bf16 verify -> redo ------------------------------------------------------------------
// The two-pass algorithm executes 4*N*C*D tensor flops because pass 2 needs the multiplier of the FINAL argmax.
// Here pass 1 runs on e4m3 copies of the operands (kind::f8f6f4: half the tensor time of a bf16 pass) and only
// GUESSES the argmax and the row maximum.  The bf16 pass then sums exp at the guessed class's multiplier while
// tracking the exact maximum / first argmax of the same bf16 logits the two-pass kernel sees: a row is final when
// the exact argmax carries the multiplier the sum was taken at (nearly always: it is the guessed class); the
// others (a few per cent on low-margin data) are listed and pass 2 alone is repeated for them at the right
// multiplier.  Labels are therefore exactly those of the two-pass kernel and confidences agree to a few ulp
// (the accepted rows' sums are shifted by the guessed maximum instead of the exact one).
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(256)
absmax16_kernel(const T* __restrict__ x, long long n_elems, unsigned int* __restrict__ max_bits) {
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems / 8; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(to_f32<T>(e[j])));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m > 0.f && isfinite(m)) atomicMax(max_bits, __float_as_uint(m));
}

// power of two that brings a maximum magnitude into [128, 256) - comfortably inside e4m3's finite range (448)
__device__ __forceinline__ int e4m3_exponent(float amax) {
  return (amax > 0.f && isfinite(amax)) ? max(-100, min(100, 7 - ilogbf(amax))) : 0;
}

__device__ __forceinline__ unsigned int pack_e4m3x4(float a, float b, float c, float d) {
  const unsigned int lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const unsigned int hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}

// whole matrix, one scale (the text side: a per-class scale would change the argmax)
template <typename T>
__global__ void __launch_bounds__(256)
quant_matrix_e4m3_kernel(const T* __restrict__ x, long long n_elems, const unsigned int* __restrict__ max_bits,
                         unsigned char* __restrict__ q, int* __restrict__ exp_out) {
  const int e = e4m3_exponent(__uint_as_float(*max_bits));
  const float sc = exp2f((float)e);
  if (blockIdx.x == 0 && threadIdx.x == 0) *exp_out = e;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems / 8; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
    const T* el = reinterpret_cast<const T*>(&v);
    uint2 o;
    o.x = pack_e4m3x4(to_f32<T>(el[0]) * sc, to_f32<T>(el[1]) * sc, to_f32<T>(el[2]) * sc, to_f32<T>(el[3]) * sc);
    o.y = pack_e4m3x4(to_f32<T>(el[4]) * sc, to_f32<T>(el[5]) * sc, to_f32<T>(el[6]) * sc, to_f32<T>(el[7]) * sc);
    reinterpret_cast<uint2*>(q)[i] = o;
  }
}

// image side: one warp per row, one power-of-two scale per ROW (a positive row scale cannot change the row's
// argmax, and it keeps every row inside e4m3's 3-bit-mantissa sweet spot whatever its norm);
// inv_scale[row] = 2^-(e_row + e_txt) turns the FP8 row maximum back into the units of the original dot product
template <typename T>
__global__ void __launch_bounds__(256)
quant_rows_e4m3_kernel(const T* __restrict__ x, long long rows, int d, const int* __restrict__ txt_exp,
                       unsigned char* __restrict__ q, float* __restrict__ inv_scale) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int chunks = d / 8;                                  // 16-byte chunks per row (<= 128)
  for (long long row = warp0; row < rows; row += n_warps) {
    const uint4* src = reinterpret_cast<const uint4*>(x + row * d);
    uint4 v[4];
    float m = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = lane + 32 * j;
      v[j] = ch < chunks ? __ldcs(src + ch) : make_uint4(0u, 0u, 0u, 0u);
      const T* el = reinterpret_cast<const T*>(&v[j]);
#pragma unroll
      for (int u = 0; u < 8; ++u) m = fmaxf(m, fabsf(to_f32<T>(el[u])));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    const int e = e4m3_exponent(m);
    const float sc = exp2f((float)e);
    uint2* dst = reinterpret_cast<uint2*>(q + row * d);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = lane + 32 * j;
      if (ch < chunks) {
        const T* el = reinterpret_cast<const T*>(&v[j]);
        uint2 o;
        o.x = pack_e4m3x4(to_f32<T>(el[0]) * sc, to_f32<T>(el[1]) * sc, to_f32<T>(el[2]) * sc, to_f32<T>(el[3]) * sc);
        o.y = pack_e4m3x4(to_f32<T>(el[4]) * sc, to_f32<T>(el[5]) * sc, to_f32<T>(el[6]) * sc, to_f32<T>(el[7]) * sc);
        dst[ch] = o;
      }
    }
    if (lane == 0) inv_scale[row] = exp2f(-(float)(e + *txt_exp));
  }
}

// Exact bf16 dot product of every image with the text row of its GUESSED class, computed by the tensor core with the
// accumulation order of the scoring kernels (K ascending in steps of 16 into an fp32 TMEM accumulator), so that it
// is bit-identical to the logit the verify pass sees in that class's column.  One 128-row tile per CTA: A = the image
// tile (TMA), B = the 128 guessed text rows gathered into the same swizzled slab layout, D = 128 x 128 of which only
// the diagonal is read.  0.4 % of one pass's tensor work; two 64-feature slabs per round keep three CTAs per SM.
constexpr int kDiagThreads = 128;
constexpr int kDiagSlabs = 2;
__global__ void __launch_bounds__(kDiagThreads)
guess_logit_kernel(const __grid_constant__ CUtensorMap map_img, const unsigned char* __restrict__ txt,
                   const int* __restrict__ guess, long long n, int c, int d, int kblocks, uint32_t idesc,
                   float* __restrict__ out) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (ptx::smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* op = smem_dyn + (base - ptx::smem_u32(smem_dyn));      // [kDiagSlabs A slabs][kDiagSlabs B slabs]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * kBlockM;
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&map_img);
    ptx::mbar_init(&bar_full, 1);
    ptx::mbar_init(&bar_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 128); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const long long my_row = row0 + threadIdx.x;
  int g = (my_row < n) ? guess[my_row] : 0;
  g = min(max(g, 0), c - 1);
  const unsigned char* src = txt + (size_t)g * (size_t)d * 2;
  const int r = threadIdx.x;
  const int rounds = (kblocks + kDiagSlabs - 1) / kDiagSlabs;
  for (int rd = 0; rd < rounds; ++rd) {
    const int kb0 = rd * kDiagSlabs, nkb = min(kDiagSlabs, kblocks - kb0);
    if (rd > 0) ptx::mbar_wait(&bar_done, (uint32_t)(rd - 1) & 1u);     // the previous round's MMAs have read the slabs
    if (threadIdx.x == 0) {
      ptx::mbar_arrive_expect_tx(&bar_full, (uint32_t)nkb * kASlabBytes);
      for (int s = 0; s < nkb; ++s)
        ptx::tma_load_2d(op + (size_t)s * kASlabBytes, &map_img, &bar_full, (kb0 + s) * kBlockK, (int)row0, ptx::kEvictNormal);
    }
    for (int s = 0; s < nkb; ++s) {
      const uint4* gsrc = reinterpret_cast<const uint4*>(src + (size_t)(kb0 + s) * 128);
      unsigned char* slab = op + (size_t)(kDiagSlabs + s) * kASlabBytes;
      uint4 v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = __ldg(gsrc + q);
#pragma unroll
      for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(slab + r * 128 + ((q ^ (r & 7)) << 4)) = v[q];
    }
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      ptx::mbar_wait(&bar_full, (uint32_t)rd & 1u);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        for (int s = 0; s < nkb; ++s) {
          const uint64_t a_desc = ptx::make_kmajor_sw128_desc(base + (uint32_t)s * kASlabBytes);
          const uint64_t b_desc = ptx::make_kmajor_sw128_desc(base + (uint32_t)(kDiagSlabs + s) * kASlabBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            ptx::umma_f16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (uint32_t)((rd | s | k) != 0));
        }
        ptx::umma_commit(&bar_done);
      }
      __syncwarp();
    }
  }
  ptx::mbar_wait(&bar_done, (uint32_t)(rounds - 1) & 1u);
  ptx::tc_fence_after();
  {
    uint32_t raw[32];
    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(warp * 32), raw);
    ptx::tmem_ld_wait(raw);
    float x = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) x = (j == lane) ? __uint_as_float(raw[j]) : x;
    if (my_row < n) out[my_row] = x;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 128);
}

// The same computation as a PERSISTENT, software-pipelined kernel (what the pipeline launches): one CTA per SM walks
// 128-row tiles; a six-stage ring of {image slab by TMA, gathered text slab by cp.async} keeps ~100 KB of loads in
// flight per SM, so the kernel runs at memory speed instead of paying the load latency once per round per CTA
// (guess_logit_kernel above: 0.54 ms at 1M x 512, 1.4 ms at 1.75M x 768; this one: see profiles/).
//   warp 0      TMA producer of the image slabs
//   warp 1      TMEM allocator + MMA issuer (128 x 128 x 16, two 128-column accumulator stages)
//   warps 2-5   thread r gathers the text row of image row r's guessed class (cp.async, three slabs in flight, then a
//               writer-side proxy fence and one arrive per warp), and reads the diagonal of the PREVIOUS tile
// The arithmetic is that of guess_logit_kernel and of the scoring kernels: K ascending in steps of 16 into one fp32
// TMEM accumulator.
constexpr int kDiagPThreads = 192;
constexpr int kDiagPStages = 6;
constexpr int kDiagPLookahead = 3;
struct __align__(16) DiagCtl {
  uint64_t full[kDiagPStages];
  uint64_t empty[kDiagPStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kDiagPThreads, 1)
guess_logit_persistent_kernel(const __grid_constant__ CUtensorMap map_img, const unsigned char* __restrict__ txt,
                              const int* __restrict__ guess, long long n, int c, int d, int kblocks, uint32_t idesc,
                              float* __restrict__ out) {
  extern __shared__ unsigned char smem_dyn[];
  DiagCtl* ctl = reinterpret_cast<DiagCtl*>(smem_dyn);
  const uint32_t base = (ptx::smem_u32(smem_dyn) + 1024u + 1023u) & ~1023u;       // ring: [stage][A slab | B slab]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = (n + kBlockM - 1) / kBlockM;
  constexpr uint32_t kStageBytes = 2 * kASlabBytes;
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&map_img);
    for (int i = 0; i < kDiagPStages; ++i) { ptx::mbar_init(&ctl->full[i], 1 + 4); ptx::mbar_init(&ctl->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&ctl->acc_full[i], 1); ptx::mbar_init(&ctl->acc_empty[i], 4); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) { ptx::tmem_alloc(&ctl->tmem_base, 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    uint32_t stage = 0, phase = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < kblocks; ++kb) {
        ptx::mbar_wait(&ctl->empty[stage], phase ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&ctl->full[stage], kASlabBytes);
          ptx::tma_load_2d(smem_dyn + (base - ptx::smem_u32(smem_dyn)) + (size_t)stage * kStageBytes, &map_img, &ctl->full[stage],
                           kb * kBlockK, (int)(tile * kBlockM), ptx::kEvictFirst);
        }
        __syncwarp();
        if (++stage == kDiagPStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0, it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1u;
      ptx::mbar_wait_short(&ctl->acc_empty[as], ((it >> 1) & 1u) ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * 128u;
      for (int kb = 0; kb < kblocks; ++kb) {
        ptx::mbar_wait_short(&ctl->full[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t a_desc = ptx::make_kmajor_sw128_desc(base + stage * kStageBytes);
          const uint64_t b_desc = ptx::make_kmajor_sw128_desc(base + stage * kStageBytes + kASlabBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            ptx::umma_f16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
          ptx::umma_commit(&ctl->empty[stage]);
          if (kb == kblocks - 1) ptx::umma_commit(&ctl->acc_full[as]);
        }
        __syncwarp();
        if (++stage == kDiagPStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    const int r = (warp & 3) * 32 + lane;                       // TMEM lane = image row within the tile = B row
    uint32_t stage = 0, phase = 0;                              // ring position of the slab being ISSUED
    uint32_t a_stage = 0;                                       // ring position of the slab being ANNOUNCED
    int in_flight = 0;
    uint32_t it = 0;
    long long prev_row = -1;
    auto announce = [&]() {                                     // the oldest issued slab has landed: hand it to the MMA warp
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ctl->full[a_stage]);
      if (++a_stage == kDiagPStages) a_stage = 0;
      --in_flight;
    };
    auto epilogue = [&](uint32_t j, long long row) {            // diagonal of accumulator stage j & 1 -> out[row]
      const uint32_t as = j & 1u;
      ptx::mbar_wait(&ctl->acc_full[as], (j >> 1) & 1u);
      ptx::tc_fence_after();
      uint32_t raw[32];
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + as * 128u + (uint32_t)((warp & 3) * 32), raw);
      ptx::tmem_ld_wait(raw);
      float x = 0.f;
#pragma unroll
      for (int q = 0; q < 32; ++q) x = (q == lane) ? __uint_as_float(raw[q]) : x;
      if (row >= 0 && row < n) out[row] = x;
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ctl->acc_empty[as]);
    };
    long long tile = blockIdx.x;
    int g = 0;
    if (tile < n_tiles) { const long long row = tile * kBlockM + r; g = row < n ? __ldg(guess + row) : 0; }
    // Copy mapping: eight consecutive lanes fetch the eight 16-byte chunks of ONE text row's 128-byte slab line, so a
    // warp instruction touches 4 lines instead of 32 (lane = row would be one L1 wavefront per lane); eight
    // instructions cover the warp's 32 rows.  The row pointers travel by shuffle from the lane that owns the row.
    const int sub = lane >> 3, q = lane & 7;
    for (; tile < n_tiles; tile += gridDim.x, ++it) {
      const long long row = tile * kBlockM + r;
      const unsigned long long mine = reinterpret_cast<unsigned long long>(txt + (size_t)min(max(g, 0), c - 1) * (size_t)d * 2);
      const unsigned char* src[8];
      uint32_t dst[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + sub;                               // row within this warp's 32
        src[i] = reinterpret_cast<const unsigned char*>(__shfl_sync(0xffffffffu, mine, rr)) + q * 16;
        const int rt = (warp & 3) * 32 + rr;                      // row within the tile
        dst[i] = base + kASlabBytes + (uint32_t)rt * 128u + (uint32_t)((q ^ (rt & 7)) << 4);
      }
      const long long next = tile + gridDim.x;                  // next tile's guess: its latency hides under this tile
      int g_next = 0;
      if (next < n_tiles) { const long long nr = next * kBlockM + r; g_next = nr < n ? __ldg(guess + nr) : 0; }
      for (int kb = 0; kb < kblocks; ++kb) {
        ptx::mbar_wait(&ctl->empty[stage], phase ^ 1u);
#pragma unroll
        for (int i = 0; i < 8; ++i) ptx::cp_async_16(dst[i] + stage * kStageBytes, src[i] + (size_t)kb * 128);
        ptx::cp_async_commit();
        if (++stage == kDiagPStages) { stage = 0; phase ^= 1u; }
        if (++in_flight > kDiagPLookahead) { ptx::cp_async_wait<kDiagPLookahead>(); announce(); }
      }
      if (kblocks <= kDiagPLookahead) {        // narrow features: the previous tile's last slabs may still be unannounced,
        ptx::cp_async_wait<0>();                //   and its accumulator cannot complete before they are
        while (in_flight > 0) announce();
      }
      if (it > 0) epilogue(it - 1, prev_row);
      prev_row = row;
      g = g_next;
    }
    // drain: the last slabs, then the last tile's diagonal
    if (in_flight > 2) { ptx::cp_async_wait<2>(); announce(); }
    if (in_flight > 1) { ptx::cp_async_wait<1>(); announce(); }
    if (in_flight > 0) { ptx::cp_async_wait<0>(); announce(); }
    if (it > 0) epilogue(it - 1, prev_row);
  }
  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 256);
}

__device__ unsigned long long g_guess_stats[2];            // {rows scored through the FP8-guess pipeline, rows redone}
__global__ void guess_stats_kernel(long long n, const int* __restrict__ redo_count) {
  if (threadIdx.x == 0) { atomicAdd(&g_guess_stats[0], (unsigned long long)n); atomicAdd(&g_guess_stats[1], (unsigned long long)*redo_count); }
}

// shared-memory plan of one variant: operand bytes, ring depth
struct SmemPlan { int stages; size_t smem; bool resident; };
static SmemPlan plan_smem(int kblocks, int ctas, bool want_resident, int parts, int min_resident_stages = 2) {
  const int avail = kSmemLimit - kCtlBytes - 1024;          // after control block and alignment slack
  const int b_bytes = kBTileBytes / ctas;
  SmemPlan pl;
  // a resident image tile next to a ring of fewer than ~4 text stages starves the tensor pipe (measured at d = 768:
  // 2 stages -> 66 % of the pipe's rate); such shapes stream both operands instead
  pl.resident = want_resident && (kblocks * kASlabBytes + min_resident_stages * b_bytes) <= avail;
  int stages = pl.resident ? (avail - kblocks * kASlabBytes) / b_bytes : avail / (parts * (kASlabBytes + b_bytes));
  if (stages > kMaxStages) stages = kMaxStages;
  pl.stages = stages;
  pl.smem = kCtlBytes + 1024 + (pl.resident ? (size_t)kblocks * kASlabBytes + (size_t)stages * b_bytes
                                            : (size_t)stages * parts * (kASlabBytes + b_bytes));
  return pl;
}

static bool guess_pipeline_applies(int mode, int64_t n, int c, int d, int dtype, int pass_sel) {
  if (mode != 0 || pass_sel >= 0 || (dtype != CCAL_BF16 && dtype != CCAL_F16)) return false;
  if (d % 128 != 0 || d > 768 || c < 512) return false;                 // redo kernel keeps 128 x d resident next to 2 stages
  const char* e = getenv("CCAL_SCORE_FP8");
  if (e && e[0] == '0') return false;
  if (e && e[0] == '1') return true;                                    // tests: force it at any size
  return n >= (int64_t)kBlockM * 2 * (num_sms() / 2) && (double)n * (double)c >= 1.0e9;
}

// development aid (CCAL_SCORE_TIMELINE=1): CUDA events between the pipeline's kernels, printed after a stream sync
struct Timeline {
  bool on;
  cudaStream_t stream;
  int n = 0;
  cudaEvent_t ev[16];
  const char* name[16];
  Timeline(cudaStream_t s) : on(getenv("CCAL_SCORE_TIMELINE") != nullptr), stream(s) {}
  void mark(const char* what) {
    if (!on || n >= 16) return;
    cudaEventCreate(&ev[n]);
    cudaEventRecord(ev[n], stream);
    name[n++] = what;
  }
  void report() {
    if (!on) return;
    cudaStreamSynchronize(stream);
    fprintf(stderr, "ccal timeline:");
    for (int i = 1; i < n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      fprintf(stderr, " %s %.3f", name[i], ms);
    }
    float tot = 0.f;
    if (n > 1) cudaEventElapsedTime(&tot, ev[0], ev[n - 1]);
    fprintf(stderr, " | total %.3f ms\n", tot);
    for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
  }
};

static int launch_guess_verify(const void* img, const void* txt, int64_t n, int c, int d, int dtype, ScoreParams p,
                               const ThrBlock& thr, cudaStream_t stream) {
  constexpr int ctas = 2;
  Timeline tl(stream);
  tl.mark("start");
  const int units = num_sms() / ctas;
  const int tile_rows = kBlockM * ctas;
  const int n_col_tiles = (c + kBlockN - 1) / kBlockN;
  int split_cap = 32768;
  if (const char* e = getenv("CCAL_REDO_SPLIT_CAP")) split_cap = atoi(e);       // development aid
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_qi = take((size_t)n * d), o_qt = take((size_t)c * d), o_inv = take((size_t)n * 4);
  const size_t o_max = take((size_t)n * 4), o_guess = take((size_t)n * 4), o_rows = take((size_t)n * 4);
  const size_t o_small = take(256);
  const size_t o_part = take((size_t)(n < split_cap ? n : split_cap) * n_col_tiles * 4);
  AsyncWorkspace workspace;
  CCAL_CUDA_OK(workspace.alloc(off, stream));
  unsigned char* ws = workspace.ptr;
  unsigned int* maxbits = (unsigned int*)(ws + o_small);
  int* txt_exp = (int*)(ws + o_small + 16);
  int* redo_count = (int*)(ws + o_small + 32);
  float* inv_scale = (float*)(ws + o_inv);
  float* row_max = (float*)(ws + o_max);
  int* guess = (int*)(ws + o_guess);
  int* redo_rows = (int*)(ws + o_rows);
  CCAL_CUDA_OK(cudaMemsetAsync(ws + o_small, 0, 256, stream));

  // ---- e4m3 copies of both operands
  const long long t_elems = (long long)c * d;
  long long g = (t_elems / 8 + 255) / 256;
  const int g_txt = (int)(g < 4 * num_sms() ? (g < 1 ? 1 : g) : 4 * num_sms());
  g = (n * 32 + 255) / 256;
  const int g_img = (int)(g < 16 * num_sms() ? g : 16 * num_sms());
  if (dtype == CCAL_BF16) {
    absmax16_kernel<__nv_bfloat16><<<g_txt, 256, 0, stream>>>((const __nv_bfloat16*)txt, t_elems, maxbits);
    quant_matrix_e4m3_kernel<__nv_bfloat16><<<g_txt, 256, 0, stream>>>((const __nv_bfloat16*)txt, t_elems, maxbits, ws + o_qt, txt_exp);
    quant_rows_e4m3_kernel<__nv_bfloat16><<<g_img, 256, 0, stream>>>((const __nv_bfloat16*)img, n, d, txt_exp, ws + o_qi, inv_scale);
  } else {
    absmax16_kernel<__half><<<g_txt, 256, 0, stream>>>((const __half*)txt, t_elems, maxbits);
    quant_matrix_e4m3_kernel<__half><<<g_txt, 256, 0, stream>>>((const __half*)txt, t_elems, maxbits, ws + o_qt, txt_exp);
    quant_rows_e4m3_kernel<__half><<<g_img, 256, 0, stream>>>((const __half*)img, n, d, txt_exp, ws + o_qi, inv_scale);
  }
  note_launch(3);
  CCAL_CUDA_OK(cudaGetLastError());
  tl.mark("quant");

  int rc;
  p.n = n; p.c = c; p.d = d;
  p.n_col_tiles = n_col_tiles;
  p.n_row_tiles = (int)((n + tile_rows - 1) / tile_rows);
  p.n_splits = 1;
  const int grid_rows = (p.n_row_tiles < units ? p.n_row_tiles : units) * ctas;
  const uint32_t fmt = (dtype == CCAL_BF16) ? 1u : 0u;
  const uint32_t shape = ((uint32_t)(kBlockN >> 3) << 17) | ((uint32_t)(tile_rows >> 4) << 24);
  CUtensorMap map_i8, map_t8, map_img, map_txt;
  if ((rc = make_map(&map_i8, ws + o_qi, n, d, kBlockM, kE4M3))) return rc;
  if ((rc = make_map(&map_t8, ws + o_qt, c, d, kBlockN / ctas, kE4M3))) return rc;
  if ((rc = make_map(&map_img, img, n, d, kBlockM, dtype))) return rc;
  if ((rc = make_map(&map_txt, txt, c, d, kBlockN / ctas, dtype))) return rc;

  // ---- kernel A: FP8 pass 1 -> guessed argmax + approximate row maximum
  {
    ScoreParams a = p;
    a.kblocks = d / 128;
    a.idesc = (1u << 4) | shape;                                        // kind::f8f6f4: D = f32, A = B = e4m3, K-major
    const SmemPlan pl = plan_smem(a.kblocks, ctas, true, 1);
    a.stages = pl.stages;
    a.pass_lo = a.pass_hi = 0;
    a.part_max = row_max; a.part_arg = guess; a.row_scale_inv = inv_scale;
    a.pred_out = nullptr; a.conf_out = nullptr; a.rowmax_out = nullptr; a.table = nullptr;
    if ((rc = launch_variant<2, true, 0, false, true, false>(map_i8, map_t8, map_i8, map_t8, a, thr, grid_rows, pl.smem, stream))) return rc;
    tl.mark("guess");
  }
  // ---- exact bf16 logit of every row's guessed class (overwrites the FP8 estimate of the row maximum)
  {
    CUtensorMap map_img128;
    if ((rc = make_map(&map_img128, img, n, d, kBlockM, dtype))) return rc;
    const uint32_t idesc128 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const long long n_tiles = (n + kBlockM - 1) / kBlockM;
    if (getenv("CCAL_DIAG_SIMPLE")) {                     // development aid: the one-tile-per-CTA form
      const size_t smem = 1024 + (size_t)2 * kDiagSlabs * kASlabBytes;
      CCAL_CUDA_OK(cudaFuncSetAttribute(guess_logit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      guess_logit_kernel<<<(unsigned)n_tiles, kDiagThreads, smem, stream>>>(
          map_img128, (const unsigned char*)txt, guess, (long long)n, c, d, d / kBlockK, idesc128, row_max);
    } else {
      const size_t smem = 2048 + (size_t)kDiagPStages * 2 * kASlabBytes;
      CCAL_CUDA_OK(cudaFuncSetAttribute(guess_logit_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
      guess_logit_persistent_kernel<<<grid, kDiagPThreads, smem, stream>>>(
          map_img128, (const unsigned char*)txt, guess, (long long)n, c, d, d / kBlockK, idesc128, row_max);
    }
    note_launch();
    CCAL_CUDA_OK(cudaGetLastError());
    tl.mark("guess_logit");
  }
  // ---- kernel B: bf16 pass 2 at the guessed multiplier + exact max / argmax; mismatches -> redo list
  p.kblocks = d / kBlockK;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | shape;
  p.pass_lo = p.pass_hi = 1;
  p.row_max_in = row_max; p.row_pred_in = guess;
  p.part_max = row_max; p.part_arg = guess;
  p.redo_count = redo_count; p.redo_rows = redo_rows;
  {
    const SmemPlan pl = plan_smem(p.kblocks, ctas, true, 1, 4);
    ScoreParams b = p;
    b.stages = pl.stages;
    if (pl.resident) rc = launch_variant<2, true, 2, false>(map_img, map_txt, map_img, map_txt, b, thr, grid_rows, pl.smem, stream);
    else rc = launch_variant<2, false, 2, false>(map_img, map_txt, map_img, map_txt, b, thr, grid_rows, pl.smem, stream);
    if (rc) return rc;
    tl.mark("verify");
  }
  // ---- redo: pass 2 for the listed rows at the multiplier of their exact argmax
  {
    ScoreParams r = p;
    const SmemPlan pl = plan_smem(r.kblocks, ctas, true, 1);
    CCAL_REQUIRE(pl.resident, "internal: redo kernel needs the resident layout (d = %d)", d);
    r.stages = pl.stages;
    r.gather_src = img;
    r.split_cap = split_cap;
    r.part_sum = (float*)(ws + o_part);
    if ((rc = launch_variant<2, true, 0, false, false, true>(map_img, map_txt, map_img, map_txt, r, thr, units * ctas, pl.smem, stream))) return rc;
    tl.mark("redo");
    split_finish_kernel<<<num_sms(), 256, 0, stream>>>(r.part_sum, row_max, guess, 0, n_col_tiles, p.scale, nullptr, p.pred_out, p.conf_out,
                                                      p.rowmax_out, p.labels, thr, p.n_thr, p.table, redo_rows, redo_count,
                                                      tile_rows, units, split_cap);
    guess_stats_kernel<<<1, 32, 0, stream>>>((long long)n, redo_count);
    note_launch(2);
    CCAL_CUDA_OK(cudaGetLastError());
    tl.mark("finish");
  }
  tl.report();
  return CCAL_OK;
}

// One launch over 16-bit operands.  `img_lo` / `txt_lo` != NULL selects the split-precision variant.
// pass_sel: -1 = both passes in one launch; 0 = pass 1 only (row maximum of the raw dot products -> p.part_max,
// first argmax -> p.part_arg); 1 = pass 2 only (reads them back through p.row_max_in / p.row_pred_in).
static int launch_fused(int mode, const void* img, const void* txt, const void* img_lo, const void* txt_lo, int64_t n,
                        int c, int d, int dtype, ScoreParams p, const ThrBlock& thr, cudaStream_t stream,
                        int pass_sel = -1) {
  const bool split = img_lo != nullptr;
  if (!split && guess_pipeline_applies(mode, n, c, d, dtype, pass_sel)) return launch_guess_verify(img, txt, n, c, d, dtype, p, thr, stream);
  int ctas = choose_ctas(n);
  // Column-split mode: when the image rows cannot fill the SMs but the vocabulary is large, every row tile is
  // cut into S class ranges (S work units per tile) and the kernel runs once per pass with tiny combines between.
  const int sms_all = num_sms();
  const int tiles128 = (int)((n + kBlockM - 1) / kBlockM);
  const int col_tiles = (c + kBlockN - 1) / kBlockN;
  int S = 1;
  if (mode == 0 && pass_sel < 0 && col_tiles >= 4 && tiles128 * 2 <= sms_all && !getenv("CCAL_SCORE_NOSPLIT")) {
    S = sms_all / tiles128;
    if (S > col_tiles) S = col_tiles;
    if (S > 1) ctas = 1;
  }
  CUtensorMap map_img, map_txt, map_img_lo, map_txt_lo;
  int rc;
  if ((rc = make_map(&map_img, img, n, d, kBlockM, dtype))) return rc;
  if ((rc = make_map(&map_txt, txt, c, d, kBlockN / ctas, dtype))) return rc;
  if ((rc = make_map(&map_img_lo, split ? img_lo : img, n, d, kBlockM, dtype))) return rc;
  if ((rc = make_map(&map_txt_lo, split ? txt_lo : txt, c, d, kBlockN / ctas, dtype))) return rc;

  p.n = n; p.c = c; p.d = d;
  p.kblocks = d / kBlockK;
  p.n_col_tiles = (c + kBlockN - 1) / kBlockN;
  const int tile_rows = kBlockM * ctas;
  p.n_row_tiles = (int)((n + tile_rows - 1) / tile_rows);
  const uint32_t fmt = (dtype == CCAL_BF16) ? 1u : 0u;
  // tcgen05 instruction descriptor, kind::f16: D=f32, A/B=fmt, both K-major, N=256, M=128 per CTA
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kBlockN >> 3) << 17) | ((uint32_t)(tile_rows >> 4) << 24);

  const int avail = kSmemLimit - kCtlBytes - 1024;          // after control block and alignment slack
  const int b_bytes = kBTileBytes / ctas;
  const int parts = split ? 2 : 1;
  bool resident = !split && (p.kblocks * kASlabBytes + 2 * kBTileBytes) <= avail;
  if (const char* e = getenv("CCAL_SCORE_RESIDENT")) {      // development aid: force the resident layout when >= 2 stages fit
    if (e[0] == '1' && !split && (p.kblocks * kASlabBytes + 2 * b_bytes) <= avail) resident = true;
  }
  int stages = resident ? (avail - p.kblocks * kASlabBytes) / b_bytes : avail / (parts * (kASlabBytes + b_bytes));
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  const size_t smem = kCtlBytes + 1024 +
                      (resident ? (size_t)p.kblocks * kASlabBytes + (size_t)stages * b_bytes
                                : (size_t)stages * parts * (kASlabBytes + b_bytes));
  const int units = num_sms() / ctas;
  const int n_work = p.n_row_tiles * S;
  const int grid = (n_work < units ? n_work : units) * ctas;
  p.n_splits = S; p.pass_lo = 0; p.pass_hi = 1;
  if (pass_sel >= 0) p.pass_lo = p.pass_hi = pass_sel;
  auto launch = [&](const ScoreParams& q) -> int {
#define CCAL_LAUNCH(C, R, M, SP) \
  launch_variant<C, R, M, SP>(map_img, map_txt, map_img_lo, map_txt_lo, q, thr, grid, smem, stream)
    if (split) {
      if (ctas == 2) return mode == 0 ? CCAL_LAUNCH(2, false, 0, true) : CCAL_LAUNCH(2, false, 1, true);
      return mode == 0 ? CCAL_LAUNCH(1, false, 0, true) : CCAL_LAUNCH(1, false, 1, true);
    }
    if (ctas == 2) {
      if (mode == 0) return resident ? CCAL_LAUNCH(2, true, 0, false) : CCAL_LAUNCH(2, false, 0, false);
      return resident ? CCAL_LAUNCH(2, true, 1, false) : CCAL_LAUNCH(2, false, 1, false);
    }
    if (mode == 0) return resident ? CCAL_LAUNCH(1, true, 0, false) : CCAL_LAUNCH(1, false, 0, false);
    return resident ? CCAL_LAUNCH(1, true, 1, false) : CCAL_LAUNCH(1, false, 1, false);
#undef CCAL_LAUNCH
  };
  if (S == 1) return launch(p);

  // ---- column-split orchestration: pass 1 per range -> combine -> pass 2 per range -> finish
  AsyncWorkspace workspace;
  const size_t per = ((size_t)n * S * 4 + 255) & ~(size_t)255, per_row = ((size_t)n * 4 + 255) & ~(size_t)255;
  const size_t per_tiles = ((size_t)n * p.n_col_tiles * 4 + 255) & ~(size_t)255;
  CCAL_CUDA_OK(workspace.alloc(2 * per + per_tiles + 2 * per_row, stream));
  float* part_max = (float*)workspace.ptr;
  int* part_arg = (int*)(workspace.ptr + per);
  float* part_sum = (float*)(workspace.ptr + 2 * per);
  float* row_max = (float*)(workspace.ptr + 2 * per + per_tiles);
  int* row_pred = (int*)(workspace.ptr + 2 * per + per_tiles + per_row);
  ScoreParams q = p;
  q.pass_lo = q.pass_hi = 0;
  q.part_max = part_max; q.part_arg = part_arg;
  if ((rc = launch(q))) return rc;
  const int rows_grid = (int)((n + 255) / 256);
  split_combine_max_kernel<<<rows_grid, 256, 0, stream>>>(part_max, part_arg, (long long)n, S, row_max, row_pred);
  note_launch();
  q.pass_lo = q.pass_hi = 1;
  q.row_max_in = row_max; q.row_pred_in = row_pred; q.part_sum = part_sum;
  if ((rc = launch(q))) return rc;
  split_finish_kernel<<<rows_grid < 4 * sms_all ? rows_grid : 4 * sms_all, 256, 0, stream>>>(
      part_sum, row_max, row_pred, (long long)n, p.n_col_tiles, p.scale, p.split_exps, p.pred_out, p.conf_out, p.rowmax_out, p.labels, thr,
      p.n_thr, p.table);
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

// ---- fp32 features: x * 2^e = hi + lo in fp16, e chosen per matrix so that max|x| * 2^e is in [2^9, 2^10)
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, long long n_elems, unsigned int* __restrict__ max_bits) {
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(max_bits, __float_as_uint(m));   // non-negative floats order like uints
}

__global__ void choose_exponents_kernel(const unsigned int* __restrict__ max_bits, int* __restrict__ exps) {
  if (threadIdx.x < 2) {
    const float m = __uint_as_float(max_bits[threadIdx.x]);
    int e = 0;
    if (m > 0.f && isfinite(m)) e = 9 - ilogbf(m);
    exps[threadIdx.x] = max(-100, min(100, e));
  }
}

__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ x, long long n_elems, const int* __restrict__ exp_ptr, __half* __restrict__ hi,
                 __half* __restrict__ lo) {
  const float s = exp2f((float)*exp_ptr);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems / 4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float f[4] = {v.x * s, v.y * s, v.z * s, v.w * s};
    __half h[4], l[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { h[u] = __float2half_rn(f[u]); l[u] = __float2half_rn(f[u] - __half2float(h[u])); }
    reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
    reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
  }
}

static int grid_for_elems(long long n, int per_thread) {
  long long want = (n / per_thread + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

static int run_fused(int mode, const void* img, const void* txt, int64_t n, int c, int d, int dtype, ScoreParams p,
                     const ThrBlock& thr, cudaStream_t stream, int pass_sel = -1) {
  int rc = ccal_check_device();
  if (rc) return rc;
  CCAL_REQUIRE(n >= 1 && c >= 1, "fused scoring: bad shape n=%lld c=%d", (long long)n, c);
  CCAL_REQUIRE(d >= 64 && d % 64 == 0 && d <= 64 * kMaxKBlocks,
               "fused scoring: feature width must be a multiple of 64 in [64, %d] (got %d)", 64 * kMaxKBlocks, d);
  CCAL_REQUIRE(dtype == CCAL_BF16 || dtype == CCAL_F16 || dtype == CCAL_F32, "fused scoring: unknown operand dtype %d", dtype);
  CCAL_REQUIRE(img && txt, "fused scoring: NULL feature pointer");
  CCAL_REQUIRE(((uintptr_t)img % 16 == 0) && ((uintptr_t)txt % 16 == 0), "fused scoring: 16-byte alignment required");
  CCAL_REQUIRE(n <= 2147483647ll - 2 * kBlockM, "fused scoring: n must fit int32 row coordinates");
  if (dtype != CCAL_F32) return launch_fused(mode, img, txt, nullptr, nullptr, n, c, d, dtype, p, thr, stream, pass_sel);
  CCAL_REQUIRE(pass_sel < 0, "the two-launch form (ccal_score_pass1 / ccal_score_pass2) takes fp16 / bf16 operands only");

  // ---- fp32 features: split into fp16 pairs (transient stream-ordered workspace), score chunk by chunk
  const int64_t chunk = n < 262144 ? n : 262144;
  CCAL_REQUIRE(mode == 0 || n <= chunk, "ccal_ts_loss_grad with fp32 operands supports up to %lld rows per call", (long long)chunk);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_thi = take((size_t)c * d * 2), o_tlo = take((size_t)c * d * 2);
  const size_t o_ihi = take((size_t)chunk * d * 2), o_ilo = take((size_t)chunk * d * 2);
  const size_t o_max = take(8), o_exp = take(8);
  AsyncWorkspace workspace;
  CCAL_CUDA_OK(workspace.alloc(off, stream));
  unsigned char* ws = workspace.ptr;
  unsigned int* maxbits = (unsigned int*)(ws + o_max);
  int* exps = (int*)(ws + o_exp);
  const float* fimg = (const float*)img;
  const float* ftxt = (const float*)txt;
  CCAL_CUDA_OK(cudaMemsetAsync(maxbits, 0, 8, stream));
  absmax_kernel<<<grid_for_elems((long long)n * d, 4), 256, 0, stream>>>(fimg, (long long)n * d, maxbits);
  note_launch();
  absmax_kernel<<<grid_for_elems((long long)c * d, 4), 256, 0, stream>>>(ftxt, (long long)c * d, maxbits + 1);
  note_launch();
  choose_exponents_kernel<<<1, 32, 0, stream>>>(maxbits, exps);
  note_launch();
  split_f16_kernel<<<grid_for_elems((long long)c * d, 4), 256, 0, stream>>>(ftxt, (long long)c * d, exps + 1, (__half*)(ws + o_thi),
                                                                           (__half*)(ws + o_tlo));
  note_launch();
  p.split_exps = exps;
  rc = CCAL_OK;
  for (int64_t q0 = 0; q0 < n && rc == CCAL_OK; q0 += chunk) {
    const int64_t m = (n - q0) < chunk ? (n - q0) : chunk;
    split_f16_kernel<<<grid_for_elems((long long)m * d, 4), 256, 0, stream>>>(fimg + q0 * d, (long long)m * d, exps, (__half*)(ws + o_ihi),
                                                                             (__half*)(ws + o_ilo));
    note_launch();
    ScoreParams pc = p;
    if (pc.pred_out) pc.pred_out += q0;
    if (pc.conf_out) pc.conf_out += q0;
    if (pc.rowmax_out) pc.rowmax_out += q0;
    if (pc.labels) pc.labels += q0;
    rc = launch_fused(mode, ws + o_ihi, ws + o_thi, ws + o_ilo, ws + o_tlo, m, c, d, CCAL_F16, pc, thr, stream);
  }
  if (rc == CCAL_OK) CCAL_CUDA_OK(cudaGetLastError());
  return rc;
}

}  // namespace ccal

using namespace ccal;

extern "C" int ccal_score_fused(const void* img, const void* txt, const float* class_conf, float logit_scale,
                                int64_t n, int c, int d, int dtype, int32_t* pred_out, float* conf_out,
                                float* rowmax_out, const int64_t* labels, const double* thresholds_host, int n_thr,
                                unsigned long long* table, ccal_stream_t stream) {
  if (n == 0) return CCAL_OK;
  ScoreParams p{};
  ThrBlock thr{};
  p.scale = logit_scale;
  p.class_conf = class_conf;
  p.pred_out = pred_out;
  p.conf_out = conf_out;
  p.rowmax_out = rowmax_out;
  p.labels = reinterpret_cast<const long long*>(labels);
  p.table = table;
  p.n_thr = 0;
  if (table != nullptr) {
    CCAL_REQUIRE(labels != nullptr, "ccal_score_fused: labels are required when a bin table is requested");
    CCAL_REQUIRE(n_thr >= 0 && n_thr <= CCAL_MAX_THRESHOLDS, "ccal_score_fused: n_thr out of range");
    CCAL_REQUIRE(n_thr == 0 || thresholds_host != nullptr, "ccal_score_fused: thresholds NULL");
    CCAL_REQUIRE(n < (1ll << 32), "ccal_score_fused: n must be < 2^32 per call when binning");
    p.n_thr = n_thr;
    for (int i = 0; i < n_thr; ++i) thr.t[i] = ceil_to_f32(thresholds_host[i]);
  }
  CCAL_REQUIRE(logit_scale > 0.f, "ccal_score_fused: logit_scale must be positive");
  return run_fused(0, img, txt, n, c, d, dtype, p, thr, (cudaStream_t)stream);
}

extern "C" int ccal_score_guess_stats(unsigned long long* out2_host, int reset) {
  CCAL_REQUIRE(out2_host != nullptr, "ccal_score_guess_stats: NULL output");
  CCAL_CUDA_OK(cudaDeviceSynchronize());
  CCAL_CUDA_OK(cudaMemcpyFromSymbol(out2_host, g_guess_stats, 2 * sizeof(unsigned long long)));
  if (reset) {
    const unsigned long long zero[2] = {0ull, 0ull};
    CCAL_CUDA_OK(cudaMemcpyToSymbol(g_guess_stats, zero, sizeof(zero)));
  }
  return CCAL_OK;
}

extern "C" int ccal_score_trace(unsigned long long* out_host, int reset) {
  CCAL_REQUIRE(out_host != nullptr, "ccal_score_trace: NULL output");
  CCAL_CUDA_OK(cudaDeviceSynchronize());
  KernelTrace tr[kTraceKinds];
  CCAL_CUDA_OK(cudaMemcpyFromSymbol(tr, g_trace, sizeof(tr)));
  for (int k = 0; k < kTraceKinds; ++k) {
    out_host[4 * k + 0] = tr[k].launches;
    out_host[4 * k + 1] = tr[k].span_ns;
    out_host[4 * k + 2] = tr[k].busy_ns;
    out_host[4 * k + 3] = tr[k].cycles;
  }
  if (reset) {
    for (int k = 0; k < kTraceKinds; ++k) tr[k] = KernelTrace{~0ull, 0, 0, 0, 0, 0, 0, 0};
    CCAL_CUDA_OK(cudaMemcpyToSymbol(g_trace, tr, sizeof(tr)));
  }
  return CCAL_OK;
}

extern "C" int ccal_score_pass1(const void* img, const void* txt, int64_t n, int c, int d, int dtype,
                                float* rowdot_max_out, int32_t* pred_out, ccal_stream_t stream) {
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(rowdot_max_out && pred_out, "ccal_score_pass1: NULL output");
  ScoreParams p{};
  ThrBlock thr{};
  p.scale = 1.0f;
  p.part_max = rowdot_max_out;
  p.part_arg = pred_out;
  return run_fused(0, img, txt, n, c, d, dtype, p, thr, (cudaStream_t)stream, 0);
}

extern "C" int ccal_score_pass2(const void* img, const void* txt, const float* class_conf, float logit_scale,
                                int64_t n, int c, int d, int dtype, const float* rowdot_max_in, const int32_t* pred_in,
                                float* conf_out, float* rowmax_out, const int64_t* labels,
                                const double* thresholds_host, int n_thr, unsigned long long* table,
                                ccal_stream_t stream) {
  if (n == 0) return CCAL_OK;
  CCAL_REQUIRE(rowdot_max_in && pred_in, "ccal_score_pass2: the row maxima and labels of ccal_score_pass1 are required");
  ScoreParams p{};
  ThrBlock thr{};
  p.scale = logit_scale;
  p.class_conf = class_conf;
  p.conf_out = conf_out;
  p.rowmax_out = rowmax_out;
  p.row_max_in = rowdot_max_in;
  p.row_pred_in = pred_in;
  p.labels = reinterpret_cast<const long long*>(labels);
  p.table = table;
  p.n_thr = 0;
  if (table != nullptr) {
    CCAL_REQUIRE(labels != nullptr, "ccal_score_pass2: labels are required when a bin table is requested");
    CCAL_REQUIRE(n_thr >= 0 && n_thr <= CCAL_MAX_THRESHOLDS, "ccal_score_pass2: n_thr out of range");
    CCAL_REQUIRE(n_thr == 0 || thresholds_host != nullptr, "ccal_score_pass2: thresholds NULL");
    CCAL_REQUIRE(n < (1ll << 32), "ccal_score_pass2: n must be < 2^32 per call when binning");
    p.n_thr = n_thr;
    for (int i = 0; i < n_thr; ++i) thr.t[i] = ceil_to_f32(thresholds_host[i]);
  }
  CCAL_REQUIRE(logit_scale > 0.f, "ccal_score_pass2: logit_scale must be positive");
  return run_fused(0, img, txt, n, c, d, dtype, p, thr, (cudaStream_t)stream, 1);
}

static int ts_loss_grad_impl(const void* img, const void* txt, const int64_t* labels, float log_scale,
                             const double* log_scale_dev, int64_t n, int c, int d, int dtype, float* row_ws, double* out2,
                             cudaStream_t stream) {
  CCAL_REQUIRE(n >= 1, "ccal_ts_loss_grad: n must be >= 1");
  CCAL_REQUIRE(labels && row_ws && out2, "ccal_ts_loss_grad: NULL pointer");
  ScoreParams p{};
  ThrBlock thr{};
  p.scale = expf(log_scale);
  p.log_scale_dev = log_scale_dev;
  p.labels = reinterpret_cast<const long long*>(labels);
  p.row_ws = row_ws;
  int rc = run_fused(1, img, txt, n, c, d, dtype, p, thr, stream);
  if (rc) return rc;
  ts_reduce_kernel<<<1, 1024, 0, stream>>>(row_ws, (long long)n, out2);
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

extern "C" int ccal_ts_loss_grad(const void* img, const void* txt, const int64_t* labels, float log_scale,
                                 int64_t n, int c, int d, int dtype, float* row_ws, double* out2,
                                 ccal_stream_t stream) {
  return ts_loss_grad_impl(img, txt, labels, log_scale, nullptr, n, c, d, dtype, row_ws, out2, (cudaStream_t)stream);
}

extern "C" int ccal_ts_loss_grad_dev(const void* img, const void* txt, const int64_t* labels, const double* log_scale_dev,
                                     int64_t n, int c, int d, int dtype, float* row_ws, double* out2,
                                     ccal_stream_t stream) {
  CCAL_REQUIRE(log_scale_dev != nullptr, "ccal_ts_loss_grad_dev: NULL log-scale pointer");
  CCAL_REQUIRE(dtype != CCAL_F32, "ccal_ts_loss_grad_dev: fp16 / bf16 operands only");
  return ts_loss_grad_impl(img, txt, labels, 0.f, log_scale_dev, n, c, d, dtype, row_ws, out2, (cudaStream_t)stream);
}

// state = {t, velocity, sum of the batch losses, batches}; one thread
__global__ void sgd_scalar_step_kernel(double* __restrict__ state, const double* __restrict__ loss_grad, double lr,
                                       double momentum, double weight_decay) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double g = loss_grad[1] + weight_decay * state[0];
    const double v = momentum * state[1] + g;
    state[1] = v;
    state[0] -= lr * v;
    state[2] += loss_grad[0];
    state[3] += 1.0;
  }
}

extern "C" int ccal_sgd_scalar_step(double* state, const double* loss_grad, double lr, double momentum, double weight_decay,
                                    ccal_stream_t stream) {
  CCAL_REQUIRE(state && loss_grad, "ccal_sgd_scalar_step: NULL pointer");
  sgd_scalar_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(state, loss_grad, lr, momentum, weight_decay);
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}
