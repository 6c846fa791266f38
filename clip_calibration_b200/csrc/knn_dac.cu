// K1: exact fp32 k-nearest-rows by Euclidean distance (warp-level top-k) and the DAC map.
//
// knn_l2_kernel: CTA = 64 query rows, loops over ALL reference rows in tiles of 64, feature
// chunks of 16 staged (transposed) in shared memory; each thread owns a 4x4 block of squared
// distances accumulated as sum (q-r)^2 in fp32 - the reference's own formula
// (np.linalg.norm(base - cur[i], axis=1)), not 2-2cos, so the "< 0.05" base-class test and
// the neighbour order do not suffer cancellation.  After a tile, each warp merges the 64 new
// candidates of its 8 query rows into per-row sorted lists that live in registers, one list
// entry per lane: an insert is one ballot (rank) + one shuffle (shift).
// Ties: lower reference index first.
#include "ccal_common.cuh"
#include <mutex>

#include <cuda_fp16.h>
#include <math_constants.h>
#include <algorithm>

namespace ccal {

constexpr int kTile = 64;        // queries per CTA and references per tile
constexpr int kChunk = 32;       // feature chunk (multiple of 16)
constexpr int kPad = 68;         // padded row length of the transposed chunks
constexpr int kRowsPerWarp = 8;
constexpr int kMaxList = CCAL_MAX_K + 1;
constexpr int kRedoSmallRows = 96;

struct TopList {                 // lane l holds the l-th smallest (d, i) seen so far
  float d;
  int i;
};

__device__ __forceinline__ void list_insert(TopList& e, float d, int i, int cap, int lane) {
  // rank of the candidate = number of entries ordered before it
  const bool before = (e.d < d) || (e.d == d && e.i < i);
  const int pos = __popc(__ballot_sync(0xffffffffu, before && lane < cap));
  const float up_d = __shfl_up_sync(0xffffffffu, e.d, 1);
  const int up_i = __shfl_up_sync(0xffffffffu, e.i, 1);
  if (lane < cap) {
    if (lane == pos) { e.d = d; e.i = i; }
    else if (lane > pos) { e.d = up_d; e.i = up_i; }
  }
}

// List mode (qlist != NULL): only the *qcount query rows named in qlist are processed (the rows the
// tensor-core filter could not prove, csrc/knn_tc.cu); CTAs loop over 64-query tiles.
__global__ void __launch_bounds__(256)
knn_l2_kernel(const float* __restrict__ ref, const float* __restrict__ query, long long nr, long long nq_all,
              int d, int k, int drop_first, float* __restrict__ dist_out, int* __restrict__ idx_out,
              const int* __restrict__ qlist, const int* __restrict__ qcount) {
  __shared__ __align__(16) float Qs[kChunk][kPad];
  __shared__ __align__(16) float Rs[kChunk][kPad];
  __shared__ float Dt[kTile][kTile + 1];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;           // 16 x 16 threads, 4 x 4 outputs each
  const int ld_row = tid >> 2, ld_col = (tid & 3) * 4;
  const int cap = min((long long)(k + (drop_first ? 1 : 0)), nr);   // list length actually used
  const long long nq = qlist ? (long long)*qcount : nq_all;
  const long long q_first = qlist ? kRedoSmallRows : 0;      // list mode: the first rows belong to knn_redo_scan_kernel
  for (long long q0 = q_first + (long long)blockIdx.x * kTile; q0 < nq; q0 += (long long)gridDim.x * kTile) {
  // global row of this thread's staged query row (list mode gathers)
  const long long ld_q = (q0 + ld_row < nq) ? (qlist ? (long long)qlist[q0 + ld_row] : q0 + ld_row) : -1;

  TopList lists[kRowsPerWarp];
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) lists[r] = TopList{CUDART_INF_F, 0x7fffffff};

  for (long long r0 = 0; r0 < nr; r0 += kTile) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    // software pipeline: the next feature chunk is loaded into registers while the current one is consumed from
    // shared memory, so the global-load latency is paid once per tile instead of once per chunk
    float4 qv[kChunk / 16], rv[kChunk / 16];
    auto load_chunk = [&](int d0) {
#pragma unroll
      for (int h = 0; h < kChunk / 16; ++h) {
        qv[h] = make_float4(0.f, 0.f, 0.f, 0.f);
        rv[h] = qv[h];
        const int dc = d0 + 16 * h + ld_col;
        if (dc < d) {                                   // d % 4 == 0 is required by the host
          if (ld_q >= 0) qv[h] = *reinterpret_cast<const float4*>(query + ld_q * d + dc);
          if (r0 + ld_row < nr) rv[h] = *reinterpret_cast<const float4*>(ref + (r0 + ld_row) * d + dc);
        }
      }
    };
    load_chunk(0);
    for (int d0 = 0; d0 < d; d0 += kChunk) {
      __syncthreads();                                   // previous chunk fully consumed
#pragma unroll
      for (int h = 0; h < kChunk / 16; ++h) {
        const int c0 = 16 * h + ld_col;
        Qs[c0 + 0][ld_row] = qv[h].x; Qs[c0 + 1][ld_row] = qv[h].y;
        Qs[c0 + 2][ld_row] = qv[h].z; Qs[c0 + 3][ld_row] = qv[h].w;
        Rs[c0 + 0][ld_row] = rv[h].x; Rs[c0 + 1][ld_row] = rv[h].y;
        Rs[c0 + 2][ld_row] = rv[h].z; Rs[c0 + 3][ld_row] = rv[h].w;
      }
      __syncthreads();
      if (d0 + kChunk < d) load_chunk(d0 + kChunk);
#pragma unroll
      for (int kk = 0; kk < kChunk; ++kk) {
        const float4 q4 = *reinterpret_cast<const float4*>(&Qs[kk][ty * 4]);
        const float4 r4 = *reinterpret_cast<const float4*>(&Rs[kk][tx * 4]);
        const float qa[4] = {q4.x, q4.y, q4.z, q4.w};
        const float ra[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const float df = ra[b] - qa[a];
            acc[a][b] = fmaf(df, df, acc[a][b]);
          }
      }
    }
    // publish the 64 x 64 distances of this tile
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) Dt[ty * 4 + a][tx * 4 + b] = sqrtf(acc[a][b]);
    __syncthreads();

    // warp-level top-k merge: warp w owns query rows 8w .. 8w+7
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const int row = warp * kRowsPerWarp + r;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int col = lane + 32 * half;
        const float cand = (r0 + col < nr) ? Dt[row][col] : CUDART_INF_F;
        float kth = __shfl_sync(0xffffffffu, lists[r].d, cap - 1);
        unsigned pending = __ballot_sync(0xffffffffu, cand < kth);
        while (pending) {
          const int src = __ffs(pending) - 1;
          pending &= pending - 1;
          const float cd = __shfl_sync(0xffffffffu, cand, src);
          if (cd < kth) {                               // warp-uniform
            list_insert(lists[r], cd, (int)(r0 + 32 * half + src), cap, lane);
            kth = __shfl_sync(0xffffffffu, lists[r].d, cap - 1);
          }
        }
      }
    }
    // Dt is rewritten only after the next tile's chunk loop, which starts with __syncthreads
  }

  const int skip = drop_first ? 1 : 0;
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) {
    const long long ql = q0 + warp * kRowsPerWarp + r;
    // entry `lane` of the list goes to output slot lane - skip
    const int slot = lane - skip;
    if (ql < nq && slot >= 0 && slot < k) {
      const long long q = qlist ? (long long)qlist[ql] : ql;
      const bool have = lane < cap;
      if (dist_out) dist_out[q * k + slot] = have ? lists[r].d : CUDART_INF_F;
      if (idx_out) idx_out[q * k + slot] = have ? lists[r].i : -1;
    }
  }
  __syncthreads();                                   // Dt / Qs / Rs are reused by the next query tile
  }
}


// Redo path for a HANDFUL of query rows (the rows the tensor-core filter could not prove).  The reference rows are
// spread over ALL SMs (thread t of the grid owns reference rows t, t + T, ...), so one unproven row costs microseconds
// instead of a serial scan by one SM.  Every lane accumulates the exact fp32 distance to its own reference row over
// the features in ascending order - the same order as the tiled scan, so both paths return identical floats.  Each
// CTA writes its sorted partial list; knn_redo_merge_kernel combines the partials with the (distance, index) order.
constexpr int kRedoSmallMax = kRedoSmallRows;   // rows [0, kRedoSmallRows) of the list take this path, the rest the tiled scan
constexpr int kRedoThreads = 128;

// all lanes call; lanes offer (cd, ci) (ci < 0: nothing); the list keeps the `cap` smallest by (d, i)
__device__ __forceinline__ void list_offer(TopList& list, float cd, int ci, int cap, int lane) {
  float kth = __shfl_sync(0xffffffffu, list.d, cap - 1);
  int kth_i = __shfl_sync(0xffffffffu, list.i, cap - 1);
  unsigned pending = __ballot_sync(0xffffffffu, ci >= 0 && (cd < kth || (cd == kth && ci < kth_i)));
  while (pending) {
    const int src = __ffs(pending) - 1;
    pending &= pending - 1;
    const float sd = __shfl_sync(0xffffffffu, cd, src);
    const int si = __shfl_sync(0xffffffffu, ci, src);
    if (sd < kth || (sd == kth && si < kth_i)) {                      // warp-uniform
      list_insert(list, sd, si, cap, lane);
      kth = __shfl_sync(0xffffffffu, list.d, cap - 1);
      kth_i = __shfl_sync(0xffffffffu, list.i, cap - 1);
    }
  }
}

__global__ void __launch_bounds__(kRedoThreads)
knn_redo_scan_kernel(const float* __restrict__ ref, const float* __restrict__ query, long long nr, int d, int cap,
                     const int* __restrict__ qlist, const int* __restrict__ qcount,
                     float* __restrict__ part_d, int* __restrict__ part_i) {
  extern __shared__ __align__(16) float s_q[];          // the query row
  __shared__ float s_d[kRedoThreads / 32][kMaxList];
  __shared__ int s_i[kRedoThreads / 32][kMaxList];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int count = min(*qcount, kRedoSmallMax);
  const long long stride = (long long)gridDim.x * kRedoThreads;
  for (int e = blockIdx.y; e < count; e += gridDim.y) {        // listed rows in parallel over grid.y (usually 0-4 rows)
    const long long q = qlist[e];
    for (int j = threadIdx.x; j < d; j += kRedoThreads) s_q[j] = query[q * d + j];
    __syncthreads();
    TopList mine{CUDART_INF_F, 0x7fffffff};
    for (long long base = (long long)blockIdx.x * kRedoThreads + warp * 32; base < nr; base += stride) {
      const long long r = base + lane;
      float dist = CUDART_INF_F;
      if (r < nr) {
        const float4* rv = reinterpret_cast<const float4*>(ref + r * d);
        const float4* qv = reinterpret_cast<const float4*>(s_q);
        float acc = 0.f;
        const int d4 = d >> 2;
        for (int j0 = 0; j0 < d4; j0 += 16) {
          float4 b[16];                                 // sixteen loads issued before the first use
#pragma unroll
          for (int u = 0; u < 16; ++u) b[u] = (j0 + u < d4) ? __ldg(rv + j0 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            if (j0 + u < d4) {
              const float4 a = qv[j0 + u];
              float df = b[u].x - a.x; acc = fmaf(df, df, acc);
              df = b[u].y - a.y; acc = fmaf(df, df, acc);
              df = b[u].z - a.z; acc = fmaf(df, df, acc);
              df = b[u].w - a.w; acc = fmaf(df, df, acc);
            }
          }
        }
        dist = sqrtf(acc);
      }
      list_offer(mine, dist, r < nr ? (int)r : -1, cap, lane);
    }
    if (lane < cap) { s_d[warp][lane] = mine.d; s_i[warp][lane] = mine.i; }
    __syncthreads();
    if (warp == 0) {
      TopList all{CUDART_INF_F, 0x7fffffff};
      for (int w = 0; w < kRedoThreads / 32; ++w) {
        const float cd = lane < cap ? s_d[w][lane] : CUDART_INF_F;
        const int ci = lane < cap ? s_i[w][lane] : -1;
        list_offer(all, cd, ci == 0x7fffffff ? -1 : ci, cap, lane);
      }
      if (lane < cap) {
        const long long o = ((long long)e * gridDim.x + blockIdx.x) * cap + lane;
        part_d[o] = all.d;
        part_i[o] = all.i;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(32)
knn_redo_merge_kernel(const float* __restrict__ part_d, const int* __restrict__ part_i, int parts, int cap, int k,
                      int drop_first, float* __restrict__ dist_out, int* __restrict__ idx_out,
                      const int* __restrict__ qlist, const int* __restrict__ qcount) {
  const int lane = threadIdx.x;
  const int count = min(*qcount, kRedoSmallMax);
  for (int e = blockIdx.x; e < count; e += gridDim.x) {
    const long long q = qlist[e];
    TopList all{CUDART_INF_F, 0x7fffffff};
    const long long total = (long long)parts * cap;
    for (long long t0 = 0; t0 < total; t0 += 32) {
      const long long t = t0 + lane;
      const float cd = t < total ? part_d[(long long)e * total + t] : CUDART_INF_F;
      const int ci = t < total ? part_i[(long long)e * total + t] : -1;
      list_offer(all, cd, ci == 0x7fffffff ? -1 : ci, cap, lane);
    }
    const int slot = lane - (drop_first ? 1 : 0);
    if (slot >= 0 && slot < k) {
      const bool have = lane < cap;
      if (dist_out) dist_out[q * k + slot] = have ? all.d : CUDART_INF_F;
      if (idx_out) idx_out[q * k + slot] = have ? all.i : -1;
    }
  }
}

// numpy's float32 add.reduce over a short contiguous vector (pairwise_sum for n < 128)
__device__ __forceinline__ float numpy_sum_f32(const float* a, int n) {
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
  }
  float r[8];
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, a[i]);
  return res;
}

__device__ __forceinline__ float dac_map_value(const float* a, const float* b, int k, int kk) {
  const float kf = (float)k;                              // divides by k even when kk < k
  const float zs_score = expf(-__fdiv_rn(numpy_sum_f32(a, kk), kf));
  const float fs_score = expf(-__fdiv_rn(numpy_sum_f32(b, kk), kf));
  return ((double)b[0] < 0.05) ? 1.0f : __fdiv_rn(fs_score, zs_score);
}

// class_conf[i] = 1 if nearest tuned distance < 0.05 else exp(-sum(tuned)/k) / exp(-sum(zs)/k)
__global__ void dac_map_kernel(const float* __restrict__ dist_zs, const float* __restrict__ dist_tuned,
                               int c, int k, int kk, float* __restrict__ class_conf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  float a[CCAL_MAX_K], b[CCAL_MAX_K];
  for (int j = 0; j < kk; ++j) { a[j] = dist_zs[(long long)i * k + j]; b[j] = dist_tuned[(long long)i * k + j]; }
  class_conf[i] = dac_map_value(a, b, k, kk);
}

__device__ __forceinline__ unsigned long long pack2f(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2f(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2f(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2f(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// ---- small DAC fits in ONE launch -------------------------------------------------------------------------
// EuroSAT / SUN397 / ImageNet-sized vocabularies (C x B up to ~2M pairs) are latency-bound: the general path is a
// chain of ~19 launches (two kNN problems x {split, tensor-core filter, verify, redo, ...} + the map) of which each
// does microseconds of work.  Here one launch does everything: CTA = (problem, 64-query tile, 64-reference tile)
// computes its exact fp32 distance tile (the tiled scan's arithmetic: features in ascending order, fma of squared
// differences, so the floats are those of ccal_knn_l2_exhaustive) and leaves a sorted partial list per query row;
// the LAST CTA of a query tile (atomic ticket) merges the partial lists by (distance, index) and writes the
// neighbours; the last of the two problems then applies the DAC map to the tile's 64 classes.
// All lanes call.  Each lane holds n candidates (distance bits of non-negative floats, index; 0x7f800000 / 0x7fffffff
// = none).  On return (pd, pi) is the smallest candidate of the whole warp that comes strictly after the incoming
// (pd, pi) in (distance, index) order - calling it cap times walks the cap smallest in order.
__device__ __forceinline__ void warp_next_smallest(const unsigned int* cd, const int* ci, int n, unsigned int& pd, int& pi) {
  unsigned int bd = 0x7f800000u;
  int bi = 0x7fffffff;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (u < n) {
      const bool after = cd[u] > pd || (cd[u] == pd && ci[u] > pi);
      if (after && (cd[u] < bd || (cd[u] == bd && ci[u] < bi))) { bd = cd[u]; bi = ci[u]; }
    }
  }
  const unsigned int md = __reduce_min_sync(0xffffffffu, bd);
  const int mi = __reduce_min_sync(0xffffffffu, bd == md ? bi : 0x7fffffff);
  pd = md;
  pi = mi;
}

struct DacSmallParams {
  const float* ref[2];
  const float* qry[2];
  float* dist[2];
  int* idx[2];
  int b, c, d, k, cap, n_qt, n_rt;
  float* part_d;               // [2][c][n_rt][cap]
  int* part_i;
  unsigned int* tickets;       // [2 * n_qt] per (problem, query tile) + [n_qt] per query tile, zeroed by the host
  float* class_conf;
};

__global__ void __launch_bounds__(256)
dac_fit_small_kernel(const __grid_constant__ DacSmallParams P) {
  __shared__ __align__(16) float Qs[kChunk][kPad];
  __shared__ __align__(16) float Rs[kChunk][kPad];
  __shared__ float Dt[kTile][kTile + 1];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;
  const int ld_row = tid >> 2, ld_col = (tid & 3) * 4;
  const int per = P.n_qt * P.n_rt;
  const int prob = blockIdx.x / per, qt = (blockIdx.x % per) / P.n_rt, rt = blockIdx.x % P.n_rt;
  const float* __restrict__ ref = P.ref[prob];
  const float* __restrict__ query = P.qry[prob];
  const int d = P.d, cap = P.cap;
  const long long q0 = (long long)qt * kTile, r0 = (long long)rt * kTile;
  const long long ld_q = (q0 + ld_row < P.c) ? q0 + ld_row : -1;

  // squared distances accumulate as packed fp32 pairs (FADD2 / FFMA2: the same IEEE operations in the same order as
  // the scalar tiled scan, two per instruction - sm_100 issues packed pairs at twice the lane rate of scalar FFMA);
  // the query chunk is staged NEGATED so that r - q is one packed add
  unsigned long long acc2[4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a) { acc2[a][0] = 0ull; acc2[a][1] = 0ull; }
  float4 qv[kChunk / 16], rv[kChunk / 16];
  auto load_chunk = [&](int d0) {
#pragma unroll
    for (int h = 0; h < kChunk / 16; ++h) {
      qv[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      rv[h] = qv[h];
      const int dc = d0 + 16 * h + ld_col;
      if (dc < d) {
        if (ld_q >= 0) qv[h] = *reinterpret_cast<const float4*>(query + ld_q * d + dc);
        if (r0 + ld_row < P.b) rv[h] = *reinterpret_cast<const float4*>(ref + (r0 + ld_row) * d + dc);
      }
    }
  };
  load_chunk(0);
  for (int d0 = 0; d0 < d; d0 += kChunk) {
    __syncthreads();
#pragma unroll
    for (int h = 0; h < kChunk / 16; ++h) {
      const int c0 = 16 * h + ld_col;
      Qs[c0 + 0][ld_row] = -qv[h].x; Qs[c0 + 1][ld_row] = -qv[h].y;
      Qs[c0 + 2][ld_row] = -qv[h].z; Qs[c0 + 3][ld_row] = -qv[h].w;
      Rs[c0 + 0][ld_row] = rv[h].x; Rs[c0 + 1][ld_row] = rv[h].y;
      Rs[c0 + 2][ld_row] = rv[h].z; Rs[c0 + 3][ld_row] = rv[h].w;
    }
    __syncthreads();
    if (d0 + kChunk < d) load_chunk(d0 + kChunk);
#pragma unroll
    for (int kk = 0; kk < kChunk; ++kk) {
      const float4 q4 = *reinterpret_cast<const float4*>(&Qs[kk][ty * 4]);          // -q
      const float4 r4 = *reinterpret_cast<const float4*>(&Rs[kk][tx * 4]);
      const float nq[4] = {q4.x, q4.y, q4.z, q4.w};
      const unsigned long long r01 = pack2f(r4.x, r4.y), r23 = pack2f(r4.z, r4.w);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const unsigned long long nqa = pack2f(nq[a], nq[a]);
        const unsigned long long d01 = add2f(r01, nqa), d23 = add2f(r23, nqa);          // r + (-q) == r - q exactly
        acc2[a][0] = fma2f(d01, d01, acc2[a][0]);
        acc2[a][1] = fma2f(d23, d23, acc2[a][1]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float v0, v1, v2, v3;
    unpack2f(acc2[a][0], v0, v1);
    unpack2f(acc2[a][1], v2, v3);
    Dt[ty * 4 + a][tx * 4 + 0] = sqrtf(v0); Dt[ty * 4 + a][tx * 4 + 1] = sqrtf(v1);
    Dt[ty * 4 + a][tx * 4 + 2] = sqrtf(v2); Dt[ty * 4 + a][tx * 4 + 3] = sqrtf(v3);
  }
  __syncthreads();

  // sorted partial list (by distance, then index) of every query row over this reference tile: `cap` rounds of a
  // warp-wide arg-min over the lane-local candidates that come after the previous winner in (distance, index) order
  // (two REDUX per round instead of a serial insertion per candidate)
#pragma unroll 1
  for (int r = 0; r < kRowsPerWarp; ++r) {
    const int row = warp * kRowsPerWarp + r;
    const long long q = q0 + row;
    unsigned int cd[2];
    int ci[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int col = lane + 32 * half;
      const bool ok = r0 + col < P.b;
      cd[half] = ok ? __float_as_uint(Dt[row][col]) : 0x7f800000u;       // non-negative floats order like their bits
      ci[half] = ok ? (int)(r0 + col) : 0x7fffffff;
    }
    unsigned int pd = 0u;
    int pi = -1;
    unsigned int my_d = 0x7f800000u;
    int my_i = 0x7fffffff;
    for (int t = 0; t < cap; ++t) {
      warp_next_smallest(cd, ci, 2, pd, pi);
      if (lane == t) { my_d = pd; my_i = pi; }
    }
    if (q < P.c && lane < cap) {
      const long long o = (((long long)prob * P.c + q) * P.n_rt + rt) * cap + lane;
      P.part_d[o] = __uint_as_float(my_d);
      P.part_i[o] = my_i;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&P.tickets[prob * P.n_qt + qt], 1u) == (unsigned)(P.n_rt - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // merge the partial lists of this query tile (same arg-min rounds over the n_rt * cap candidates, re-read from L2
  // each round when a lane holds more than four of them)
  const int k = P.k;
#pragma unroll 1
  for (int r = 0; r < kRowsPerWarp; ++r) {
    const long long q = q0 + warp * kRowsPerWarp + r;
    if (q >= P.c) continue;                              // warp-uniform
    const int total = P.n_rt * cap;
    const long long o = ((long long)prob * P.c + q) * total;
    unsigned int pd = 0u;
    int pi = -1;
    unsigned int my_d = 0x7f800000u;
    int my_i = -1;
    if (total <= 128) {
      unsigned int cd[4];
      int ci[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = lane + 32 * u;
        cd[u] = t < total ? __float_as_uint(__ldcg(P.part_d + o + t)) : 0x7f800000u;
        ci[u] = t < total ? __ldcg(P.part_i + o + t) : 0x7fffffff;
      }
      for (int t = 0; t < cap; ++t) {
        warp_next_smallest(cd, ci, 4, pd, pi);
        if (lane == t) { my_d = pd; my_i = pi; }
      }
    } else {
      for (int t = 0; t < cap; ++t) {
        unsigned int bd = 0x7f800000u;
        int bi = 0x7fffffff;
        for (int u = lane; u < total; u += 32) {
          const unsigned int d1 = __float_as_uint(__ldcg(P.part_d + o + u));
          const int i1 = __ldcg(P.part_i + o + u);
          const bool after = d1 > pd || (d1 == pd && i1 > pi);
          if (after && (d1 < bd || (d1 == bd && i1 < bi))) { bd = d1; bi = i1; }
        }
        const unsigned int md = __reduce_min_sync(0xffffffffu, bd);
        const int mi = __reduce_min_sync(0xffffffffu, bd == md ? bi : 0x7fffffff);
        pd = md; pi = mi;
        if (lane == t) { my_d = pd; my_i = pi; }
      }
    }
    if (lane < k) {
      const bool have = lane < cap && my_i != 0x7fffffff;
      P.dist[prob][q * k + lane] = have ? __uint_as_float(my_d) : CUDART_INF_F;
      if (P.idx[prob]) P.idx[prob][q * k + lane] = have ? my_i : -1;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&P.tickets[2 * P.n_qt + qt], 1u) == 1u);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < kTile && q0 + tid < P.c) {
    const long long i = q0 + tid;
    float a[CCAL_MAX_K], bvals[CCAL_MAX_K];
    for (int j = 0; j < cap; ++j) { a[j] = __ldcg(P.dist[0] + i * k + j); bvals[j] = __ldcg(P.dist[1] + i * k + j); }
    P.class_conf[i] = dac_map_value(a, bvals, k, cap);
  }
}

// ---- DAC fit in the reference's float16 arithmetic ----------------------------------------------------------
// The reference's default precision is fp16 (train.py:152), its cached text features are float16 numpy arrays, and
// np.linalg.norm / np.sum / np.exp keep that dtype: every elementwise result is rounded to
// half, the squares are summed in float32 in numpy's pairwise order and rounded to half once per row.  The float32
// path above is closer to the exact answer, but differs from what the reference computes by up to 1e-3 relative in
// class_confidence; this kernel reproduces numpy's half arithmetic operation by operation (opt-in:
// DistanseAwareCalibration.fit(..., arithmetic="input")).  One CTA per class, one thread per base row.
__device__ __forceinline__ float half_sq_term(__half r, float q) {
  const float diff = __half2float(__float2half_rn(__fsub_rn(__half2float(r), q)));     // HALF_subtract
  return __half2float(__float2half_rn(__fmul_rn(diff, diff)));                         // HALF_multiply
}

// numpy's pairwise_sum over the n rounded squares of one row (float32 accumulation; numpy/core/src/umath/loops_utils.h)
__device__ float half_row_pairwise(const __half* __restrict__ r, const float* __restrict__ q, int n) {
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res = __fadd_rn(res, half_sq_term(r[i], q[i]));
    return res;
  }
  if (n <= 128) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = half_sq_term(r[j], q[j]);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = __fadd_rn(a[j], half_sq_term(r[i + j], q[i + j]));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])),
                          __fadd_rn(__fadd_rn(a[4], a[5]), __fadd_rn(a[6], a[7])));
    for (; i < n; ++i) res = __fadd_rn(res, half_sq_term(r[i], q[i]));
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __fadd_rn(half_row_pairwise(r, q, n2), half_row_pairwise(r + n2, q + n2, n - n2));
}

constexpr int kHalfFitMaxBase = 16384;

__global__ void __launch_bounds__(256)
dac_fit_half_kernel(const __half* __restrict__ base_zs, const __half* __restrict__ cur_zs,
                    const __half* __restrict__ base_tuned, const __half* __restrict__ cur_tuned, int b, int c, int d, int k,
                    float* __restrict__ class_conf, int* __restrict__ idx_zs, int* __restrict__ idx_tuned,
                    float* __restrict__ dist_zs, float* __restrict__ dist_tuned) {
  extern __shared__ __align__(16) unsigned char half_fit_smem[];
  float* s_q = reinterpret_cast<float*>(half_fit_smem);                     // [d] the class's feature row
  unsigned short* s_d = reinterpret_cast<unsigned short*>(s_q + d);         // [b] distances (half bits; non-negative)
  __shared__ unsigned int s_best[8];
  __shared__ int s_besti[8];
  __shared__ float s_top[2][CCAL_MAX_K];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kk = k < b ? k : b;
  for (int cls = blockIdx.x; cls < c; cls += gridDim.x) {
    for (int prob = 0; prob < 2; ++prob) {
      const __half* base = prob ? base_tuned : base_zs;
      const __half* cur = (prob ? cur_tuned : cur_zs) + (size_t)cls * d;
      __syncthreads();
      for (int j = tid; j < d; j += 256) s_q[j] = __half2float(cur[j]);
      __syncthreads();
      for (int r = tid; r < b; r += 256) {
        const float s = half_row_pairwise(base + (size_t)r * d, s_q, d);
        const __half s16 = __float2half_rn(s);                               // HALF_add reduce: one rounding per row
        const __half dist = __float2half_rn(sqrtf(__half2float(s16)));       // np.sqrt on float16
        s_d[r] = __half_as_ushort(dist);
      }
      __syncthreads();
      // k rounds of (distance, index) arg-min; non-negative halves order like their bit patterns
      for (int t = 0; t < kk; ++t) {
        unsigned int best = 0xffffffffu;
        int besti = 0x7fffffff;
        for (int r = tid; r < b; r += 256) {
          const unsigned int v = s_d[r];
          if (v < best) { best = v; besti = r; }                             // ascending r per thread: first index wins
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const unsigned int ov = __shfl_xor_sync(0xffffffffu, best, off);
          const int oi = __shfl_xor_sync(0xffffffffu, besti, off);
          if (ov < best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        if (lane == 0) { s_best[warp] = best; s_besti[warp] = besti; }
        __syncthreads();
        if (tid == 0) {
          for (int w = 1; w < 8; ++w)
            if (s_best[w] < best || (s_best[w] == best && s_besti[w] < besti)) { best = s_best[w]; besti = s_besti[w]; }
          const float dv = __half2float(__ushort_as_half((unsigned short)best));
          s_top[prob][t] = dv;
          float* dout = prob ? dist_tuned : dist_zs;
          int* iout = prob ? idx_tuned : idx_zs;
          if (dout) dout[(size_t)cls * k + t] = dv;
          if (iout) iout[(size_t)cls * k + t] = besti;
          s_d[besti] = 0xffffu;                                              // taken (above every finite half / inf)
        }
        __syncthreads();
      }
      if (tid == 0)
        for (int t = kk; t < k; ++t) {
          float* dout = prob ? dist_tuned : dist_zs;
          int* iout = prob ? idx_tuned : idx_zs;
          if (dout) dout[(size_t)cls * k + t] = CUDART_INF_F;
          if (iout) iout[(size_t)cls * k + t] = -1;
        }
    }
    if (tid == 0) {
      float score[2];
      for (int prob = 0; prob < 2; ++prob) {
        const __half sum16 = __float2half_rn(numpy_sum_f32(s_top[prob], kk));            // np.sum on float16
        const __half q16 = __float2half_rn(__fdiv_rn(-__half2float(sum16), (float)k));   // -sum / k (k <= 16 is exact in half)
        score[prob] = __half2float(__float2half_rn((float)exp((double)__half2float(q16))));   // np.exp on float16
      }
      const float thr = __half2float(__float2half_rn(0.05f));      // NEP 50: the Python float is compared as float16
      const __half ratio = __float2half_rn(__fdiv_rn(score[1], score[0]));
      class_conf[cls] = (s_top[1][0] < thr) ? 1.0f : __half2float(ratio);
    }
  }
}

static bool small_fit_applies(int b, int c, int d) {
  if (getenv("CCAL_DAC_NO_SMALL_FIT")) return false;
  return (long long)b * (long long)c <= (1ll << 21) && b <= kTile * 64 && d % 4 == 0;
}

static int launch_dac_fit_small(const float* base_zs, const float* cur_zs, const float* base_tuned, const float* cur_tuned,
                                int b, int c, int d, int k, float* class_conf, int32_t* idx_zs, int32_t* idx_tuned,
                                float* dist_zs, float* dist_tuned, cudaStream_t stream) {
  DacSmallParams P{};
  P.ref[0] = base_zs; P.ref[1] = base_tuned;
  P.qry[0] = cur_zs; P.qry[1] = cur_tuned;
  P.dist[0] = dist_zs; P.dist[1] = dist_tuned;
  P.idx[0] = idx_zs; P.idx[1] = idx_tuned;
  P.b = b; P.c = c; P.d = d; P.k = k;
  P.cap = k < b ? k : b;
  P.n_qt = (c + kTile - 1) / kTile;
  P.n_rt = (b + kTile - 1) / kTile;
  P.class_conf = class_conf;
  const size_t cells = (size_t)2 * c * P.n_rt * P.cap;
  const size_t o_i = (cells * sizeof(float) + 255) & ~(size_t)255;
  const size_t o_t = (o_i + cells * sizeof(int) + 255) & ~(size_t)255;
  const size_t ticket_bytes = (size_t)3 * P.n_qt * sizeof(unsigned int);
  AsyncWorkspace ws;
  CCAL_CUDA_OK(ws.alloc(o_t + ticket_bytes, stream));
  P.part_d = reinterpret_cast<float*>(ws.ptr);
  P.part_i = reinterpret_cast<int*>(ws.ptr + o_i);
  P.tickets = reinterpret_cast<unsigned int*>(ws.ptr + o_t);
  CCAL_CUDA_OK(cudaMemsetAsync(P.tickets, 0, ticket_bytes, stream));
  dac_fit_small_kernel<<<2 * P.n_qt * P.n_rt, 256, 0, stream>>>(P);
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

int launch_knn_exact(const float* ref, const float* query, int64_t nr, int64_t nq, int d, int k, int drop_first,
                     float* dist_out, int32_t* idx_out, const int* qlist, const int* qcount, cudaStream_t stream) {
  AsyncWorkspace ws;
  if (qlist != nullptr) {
    const int cap = (int)std::min<int64_t>(k + (drop_first ? 1 : 0), nr);
    const int parts = (int)std::min<int64_t>(num_sms(), (nr + kRedoThreads - 1) / kRedoThreads);
    const size_t cells = (size_t)kRedoSmallMax * parts * cap;
    CCAL_CUDA_OK(ws.alloc(cells * (sizeof(float) + sizeof(int)), stream));
    float* part_d = reinterpret_cast<float*>(ws.ptr);
    int* part_i = reinterpret_cast<int*>(ws.ptr + cells * sizeof(float));
    trace_mark(stream, 30);
    knn_redo_scan_kernel<<<dim3((unsigned)parts, 16), kRedoThreads, (size_t)d * sizeof(float), stream>>>(
        ref, query, (long long)nr, d, cap, qlist, qcount, part_d, part_i);
    note_launch();
    trace_mark(stream, 31);
    knn_redo_merge_kernel<<<kRedoSmallMax, 32, 0, stream>>>(part_d, part_i, parts, cap, k, drop_first, dist_out, idx_out,
                                                            qlist, qcount);
    note_launch();
    trace_mark(stream, 32);
  }
  long long grid = (nq + kTile - 1) / kTile;
  const long long cap = (long long)num_sms() * 4;
  if (qlist != nullptr && grid > cap) grid = cap;     // list mode: rows beyond the first kRedoSmallRows (normally none)
  if (grid > 2147483647ll) grid = 2147483647ll;
  knn_l2_kernel<<<(int)grid, 256, 0, stream>>>(ref, query, (long long)nr, (long long)nq, d, k, drop_first,
                                               dist_out, idx_out, qlist, qcount);
  note_launch();
  trace_mark(stream, 33);
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

int knn_l2_tensor(const float* ref, const float* query, int64_t nr, int64_t nq, int d, int k, int drop_first,
                  float* dist_out, int32_t* idx_out, cudaStream_t stream);

// tensor-core filter + exact verification when the problem is big enough to amortise it
static int launch_knn(const float* ref, const float* query, int64_t nr, int64_t nq, int d, int k, int drop_first,
                      float* dist_out, int32_t* idx_out, cudaStream_t stream) {
  // below ~64k (query, reference) pairs the exhaustive scan's single launch wins
  const bool tc_ok = (d % 64 == 0) && d <= 1024 && nr >= 64 && nq >= 128 && nr * nq >= (1ll << 16);
  if (tc_ok) return knn_l2_tensor(ref, query, nr, nq, d, k, drop_first, dist_out, idx_out, stream);
  return launch_knn_exact(ref, query, nr, nq, d, k, drop_first, dist_out, idx_out, nullptr, nullptr, stream);
}

}  // namespace ccal

using namespace ccal;

extern "C" int ccal_knn_l2(const float* ref, const float* query, int64_t nr, int64_t nq, int d, int k,
                           int drop_first, float* dist_out, int32_t* idx_out, ccal_stream_t stream) {
  CCAL_REQUIRE(nr >= 1 && nq >= 0, "ccal_knn_l2: bad row counts nr=%lld nq=%lld", (long long)nr, (long long)nq);
  CCAL_REQUIRE(nr < 2147483647ll, "ccal_knn_l2: nr must fit int32");
  CCAL_REQUIRE(d >= 4 && d % 4 == 0, "ccal_knn_l2: d must be a positive multiple of 4 (got %d)", d);
  CCAL_REQUIRE(k >= 1 && k <= CCAL_MAX_K, "ccal_knn_l2: k must be in 1..%d (got %d)", CCAL_MAX_K, k);
  if (nq == 0) return CCAL_OK;
  CCAL_REQUIRE(ref && query, "ccal_knn_l2: NULL input");
  CCAL_REQUIRE(((uintptr_t)ref % 16 == 0) && ((uintptr_t)query % 16 == 0), "ccal_knn_l2: 16-byte alignment required");
  return launch_knn(ref, query, nr, nq, d, k, drop_first, dist_out, idx_out, (cudaStream_t)stream);
}

extern "C" int ccal_knn_l2_exhaustive(const float* ref, const float* query, int64_t nr, int64_t nq, int d, int k,
                                      int drop_first, float* dist_out, int32_t* idx_out, ccal_stream_t stream) {
  CCAL_REQUIRE(nr >= 1 && nq >= 0 && nr < 2147483647ll, "ccal_knn_l2_exhaustive: bad row counts");
  CCAL_REQUIRE(d >= 4 && d % 4 == 0, "ccal_knn_l2_exhaustive: d must be a positive multiple of 4 (got %d)", d);
  CCAL_REQUIRE(k >= 1 && k <= CCAL_MAX_K, "ccal_knn_l2_exhaustive: k must be in 1..%d (got %d)", CCAL_MAX_K, k);
  if (nq == 0) return CCAL_OK;
  CCAL_REQUIRE(ref && query, "ccal_knn_l2_exhaustive: NULL input");
  CCAL_REQUIRE(((uintptr_t)ref % 16 == 0) && ((uintptr_t)query % 16 == 0), "ccal_knn_l2_exhaustive: 16-byte alignment required");
  return launch_knn_exact(ref, query, nr, nq, d, k, drop_first, dist_out, idx_out, nullptr, nullptr, (cudaStream_t)stream);
}

// One library-owned non-blocking stream per device for the second kNN chain of ccal_dac_fit (created on first use,
// never destroyed).  OPT-IN (CCAL_DAC_TWO_STREAMS=1): it buys 0.09 ms per large fit on an otherwise idle GPU, but a
// stream the caller does not know about is one more stream sharing the device's hardware queues.  Measured on 2 x
// B200 (round 2, profiles/r02i_stall_trace.txt): in a pipelined evaluation loop - copy stream blocked on staging
// buffers that wait for the scoring kernels that wait for THIS fit - the second chain stopped between its
// split_rows launch and its filter kernel in 3 of 26 processes and never resumed (the chain on the caller's stream
// had finished): work of the extra stream queued behind another stream's blocked wait.  With one stream the fit is
// ordered purely by the caller's stream, which 40+ multi-GPU runs never stalled on.
static cudaStream_t fit_side_stream() {
  static std::mutex mu;
  static cudaStream_t streams[64] = {};
  const char* two = getenv("CCAL_DAC_TWO_STREAMS");
  if (two == nullptr || two[0] == '0' || getenv("CCAL_DAC_ONE_STREAM")) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (streams[dev] == nullptr && cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess) {
    streams[dev] = nullptr;
    cudaGetLastError();
  }
  return streams[dev];
}

extern "C" int ccal_dac_fit(const float* base_zs, const float* cur_zs, const float* base_tuned,
                            const float* cur_tuned, int b, int c, int d, int k,
                            float* class_conf_out, int32_t* knn_idx_zs_out, int32_t* knn_idx_tuned_out,
                            float* knn_dist_zs_out, float* knn_dist_tuned_out, ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE(b >= 1 && c >= 0, "ccal_dac_fit: bad class counts b=%d c=%d", b, c);
  CCAL_REQUIRE(d >= 4 && d % 4 == 0, "ccal_dac_fit: d must be a positive multiple of 4 (got %d)", d);
  CCAL_REQUIRE(k >= 1 && k <= CCAL_MAX_K, "ccal_dac_fit: k must be in 1..%d (got %d)", CCAL_MAX_K, k);
  if (c == 0) return CCAL_OK;
  CCAL_REQUIRE(base_zs && cur_zs && base_tuned && cur_tuned && class_conf_out, "ccal_dac_fit: NULL input");
  CCAL_REQUIRE(knn_dist_zs_out && knn_dist_tuned_out,
               "ccal_dac_fit: the two [c,k] distance buffers are required (they double as workspace)");
  if (small_fit_applies(b, c, d))
    return launch_dac_fit_small(base_zs, cur_zs, base_tuned, cur_tuned, b, c, d, k, class_conf_out, knn_idx_zs_out,
                                knn_idx_tuned_out, knn_dist_zs_out, knn_dist_tuned_out, stream);
  // The two kNN problems (zero-shot space, tuned space) are independent chains of ~9 launches each, several of them
  // short or narrow (operand scan, split, verify, the redo of a handful of rows).  With CCAL_DAC_TWO_STREAMS=1 they run
  // on two streams - the caller's and a library-owned one, forked and joined by events - so that one chain's narrow
  // kernels run underneath the other's tensor-core filter (see fit_side_stream for why that is not the default).
  cudaStream_t side = fit_side_stream();
  int rc;
  if (side != nullptr) {
    cudaEvent_t fork, join;
    CCAL_CUDA_OK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    CCAL_CUDA_OK(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
    trace_mark(stream, 1);
    CCAL_CUDA_OK(cudaEventRecord(fork, stream));
    CCAL_CUDA_OK(cudaStreamWaitEvent(side, fork, 0));
    trace_mark(side, 2);
    rc = launch_knn(base_tuned, cur_tuned, b, c, d, k, 0, knn_dist_tuned_out, knn_idx_tuned_out, side);
    trace_mark(side, 3);
    int rc2 = launch_knn(base_zs, cur_zs, b, c, d, k, 0, knn_dist_zs_out, knn_idx_zs_out, stream);
    trace_mark(stream, 4);
    cudaEventRecord(join, side);                    // joined even on failure: the side stream must not outlive the call
    cudaStreamWaitEvent(stream, join, 0);
    trace_mark(stream, 5);
    cudaEventDestroy(fork);
    cudaEventDestroy(join);
    if (rc) return rc;
    if (rc2) return rc2;
  } else {
    rc = launch_knn(base_zs, cur_zs, b, c, d, k, 0, knn_dist_zs_out, knn_idx_zs_out, stream);
    if (rc) return rc;
    rc = launch_knn(base_tuned, cur_tuned, b, c, d, k, 0, knn_dist_tuned_out, knn_idx_tuned_out, stream);
    if (rc) return rc;
  }
  const int kk = k < b ? k : b;
  dac_map_kernel<<<(c + 127) / 128, 128, 0, stream>>>(knn_dist_zs_out, knn_dist_tuned_out, c, k, kk, class_conf_out);
  note_launch();
  trace_mark(stream, 6);
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}

extern "C" int ccal_dac_fit_f16(const void* base_zs, const void* cur_zs, const void* base_tuned, const void* cur_tuned,
                                int b, int c, int d, int k, float* class_conf_out, int32_t* knn_idx_zs_out,
                                int32_t* knn_idx_tuned_out, float* knn_dist_zs_out, float* knn_dist_tuned_out,
                                ccal_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CCAL_REQUIRE(b >= 1 && c >= 0, "ccal_dac_fit_f16: bad class counts b=%d c=%d", b, c);
  CCAL_REQUIRE(b <= kHalfFitMaxBase, "ccal_dac_fit_f16: at most %d base classes (got %d)", kHalfFitMaxBase, b);
  CCAL_REQUIRE(d >= 1 && d <= 4096, "ccal_dac_fit_f16: d must be in 1..4096 (got %d)", d);
  CCAL_REQUIRE(k >= 1 && k <= CCAL_MAX_K, "ccal_dac_fit_f16: k must be in 1..%d (got %d)", CCAL_MAX_K, k);
  if (c == 0) return CCAL_OK;
  CCAL_REQUIRE(base_zs && cur_zs && base_tuned && cur_tuned && class_conf_out, "ccal_dac_fit_f16: NULL input");
  int rc = ccal_check_device();
  if (rc) return rc;
  const size_t smem = (size_t)d * sizeof(float) + (size_t)b * sizeof(unsigned short) + 16;
  CCAL_CUDA_OK(cudaFuncSetAttribute(dac_fit_half_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = c < 8 * num_sms() ? c : 8 * num_sms();
  dac_fit_half_kernel<<<grid, 256, smem, stream>>>((const __half*)base_zs, (const __half*)cur_zs, (const __half*)base_tuned,
                                                   (const __half*)cur_tuned, b, c, d, k, class_conf_out, knn_idx_zs_out,
                                                   knn_idx_tuned_out, knn_dist_zs_out, knn_dist_tuned_out);
  note_launch();
  CCAL_CUDA_OK(cudaGetLastError());
  return CCAL_OK;
}
