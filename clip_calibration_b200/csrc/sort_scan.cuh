// Device-wide primitives of the isotonic / histogram-binning calibrators, written out here instead of calling a
// library: a stable LSD radix sort of (float64 key, uint8 payload) pairs, an int32 prefix sum, and a flagged
// compaction.  All three are HBM-bound byte/integer work: coalesced loads, shared-memory staging so that the
// scattered writes of a sort pass leave the CTA as runs of consecutive addresses, grids that cover the data once.
//
//   sort   8 digits of 8 bits.  One read pass takes the AND and the OR of all keys; a digit on which every key agrees
//          (sign + high exponent bits of probabilities, typically 1-2 of the 8) is skipped.  Per remaining digit:
//          per-tile digit counts (digit-major) -> exclusive prefix sum -> scatter.  A tile is 2048 consecutive keys:
//          each warp owns 256 consecutive keys and ranks them 32 at a time with `match.any` (rank = keys of the same
//          digit in earlier rounds of this warp + lower lanes of this round), so equal digits keep their input order.
//          Keys are staged in shared memory at their tile-local sorted position and written out digit run by digit
//          run.  float64 keys are mapped to order-preserving unsigned integers on the way in and back on the way out
//          (-0.0 sorts before +0.0; NaNs with the sign bit clear sort last).
//   scan   reduce-then-scan over 4096-element tiles: tile sums, one CTA scans the tile sums, tiles are rescanned with
//          their offset.  Striped (coalesced) global accesses, blocked thread-local scans through padded shared memory.
#pragma once

#include "ccal_common.cuh"

namespace ccal {

// ------------------------------------------------------------------------------------------------ prefix sum (int32)
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;          // 4096

__device__ __forceinline__ int block_reduce_256(int v, int* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = __reduce_add_sync(0xffffffffu, v);
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  int tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) tot += s_warp[w];
  __syncthreads();
  return tot;
}

// exclusive prefix of one value per thread over the CTA (256 threads); *total receives the CTA sum
__device__ __forceinline__ int block_exclusive_256(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += o;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    const int c = s_warp[w];
    if (w < warp) base += c;
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_tile_sums_kernel(const int* __restrict__ in, long long n, int* __restrict__ sums) {
  __shared__ int s_warp[kScanThreads / 32];
  const long long base = (long long)blockIdx.x * kScanTile;
  int v = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const long long i = base + j * kScanThreads + threadIdx.x;
    if (i < n) v += in[i];
  }
  const int tot = block_reduce_256(v, s_warp);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// one CTA: sums[] -> exclusive prefix in place
__global__ void __launch_bounds__(kScanThreads)
scan_sums_kernel(int* __restrict__ sums, int nt) {
  __shared__ int s_warp[kScanThreads / 32];
  int carry = 0;
  for (int base = 0; base < nt; base += kScanThreads) {
    const int i = base + threadIdx.x;
    const int v = i < nt ? sums[i] : 0;
    int tot;
    const int ex = block_exclusive_256(v, s_warp, &tot);
    if (i < nt) sums[i] = carry + ex;
    carry += tot;
  }
}

template <bool kInclusive>
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const int* in, int* out, long long n, const int* __restrict__ offsets) {   // in may be out
  __shared__ int s_tile[kScanTile + kScanTile / 16];           // thread t's 16 items start at t * 17
  __shared__ int s_warp[kScanThreads / 32];
  const long long base = (long long)blockIdx.x * kScanTile;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const int p = j * kScanThreads + threadIdx.x;
    const long long i = base + p;
    s_tile[p + (p >> 4)] = i < n ? in[i] : 0;
  }
  __syncthreads();
  int v[kScanItems];
  int sum = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    v[j] = s_tile[threadIdx.x * 17 + j];
    sum += v[j];
  }
  int tot;
  int run = offsets[blockIdx.x] + block_exclusive_256(sum, s_warp, &tot);
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const int x = v[j];
    s_tile[threadIdx.x * 17 + j] = kInclusive ? run + x : run;
    run += x;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const int p = j * kScanThreads + threadIdx.x;
    const long long i = base + p;
    if (i < n) out[i] = s_tile[p + (p >> 4)];
  }
}

inline size_t scan_workspace_bytes(long long n) {
  return (size_t)((n + kScanTile - 1) / kScanTile + 1) * sizeof(int);
}

// out[i] = in[0] + ... + in[i] (inclusive) or in[0] + ... + in[i-1] (exclusive); in == out is allowed.  Sums must fit
// int32.  ws: scan_workspace_bytes(n).
inline cudaError_t prefix_sum_i32(const int* in, int* out, long long n, bool inclusive, void* ws, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const long long nt = (n + kScanTile - 1) / kScanTile;
  int* sums = (int*)ws;
  scan_tile_sums_kernel<<<(unsigned)nt, kScanThreads, 0, stream>>>(in, n, sums);
  scan_sums_kernel<<<1, kScanThreads, 0, stream>>>(sums, (int)nt);
  if (inclusive) scan_apply_kernel<true><<<(unsigned)nt, kScanThreads, 0, stream>>>(in, out, n, sums);
  else scan_apply_kernel<false><<<(unsigned)nt, kScanThreads, 0, stream>>>(in, out, n, sums);
  note_launch(3);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ flagged compaction
// pos = inclusive prefix sum of keep (as int): element i with keep[i] goes to slot pos[i] - 1 of both outputs
__global__ void compact_flags_kernel(const unsigned char* __restrict__ keep, int n, int* __restrict__ as_int) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) as_int[i] = keep[i] ? 1 : 0;
}

__global__ void compact_pair_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                    const unsigned char* __restrict__ keep, const int* __restrict__ pos, int n,
                                    double* __restrict__ a_out, double* __restrict__ b_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && keep[i]) {
    a_out[pos[i] - 1] = a[i];
    b_out[pos[i] - 1] = b[i];
  }
}

// ------------------------------------------------------------------------------------------------ radix sort
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;           // 2048 keys per CTA
constexpr int kSortWarpSpan = 32 * kSortItems;                 // 256 consecutive keys per warp

__device__ __forceinline__ unsigned long long f64_to_ordered(double x) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_to_f64(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// bits[0] &= every key, bits[1] |= every key (ordered form): a digit on which AND and OR agree is the same in all keys
__global__ void __launch_bounds__(kSortThreads)
sort_and_or_kernel(const double* __restrict__ keys, long long n, unsigned long long* __restrict__ bits) {
  unsigned long long a = ~0ull, o = 0ull;
  const long long stride = (long long)gridDim.x * kSortThreads;
  for (long long i = (long long)blockIdx.x * kSortThreads + threadIdx.x; i < n; i += stride) {
    const unsigned long long k = f64_to_ordered(keys[i]);
    a &= k;
    o |= k;
  }
  const unsigned full = 0xffffffffu;
  const unsigned a_lo = __reduce_and_sync(full, (unsigned)a), a_hi = __reduce_and_sync(full, (unsigned)(a >> 32));
  const unsigned o_lo = __reduce_or_sync(full, (unsigned)o), o_hi = __reduce_or_sync(full, (unsigned)(o >> 32));
  if ((threadIdx.x & 31) == 0) {
    atomicAnd(&bits[0], ((unsigned long long)a_hi << 32) | a_lo);
    atomicOr(&bits[1], ((unsigned long long)o_hi << 32) | o_lo);
  }
}

// keys as they are held in a pass buffer: raw float64 bits in the caller's input (kRawIn), ordered integers otherwise
template <bool kRawIn>
__device__ __forceinline__ unsigned long long sort_load_key(const unsigned long long* __restrict__ keys, long long i) {
  const unsigned long long u = keys[i];
  return kRawIn ? f64_to_ordered(__longlong_as_double((long long)u)) : u;
}

// counts[value * n_tiles + tile]
template <bool kRawIn>
__global__ void __launch_bounds__(kSortThreads)
sort_tile_counts_kernel(const unsigned long long* __restrict__ keys, long long n, int shift, int n_tiles,
                        int* __restrict__ counts) {
  __shared__ unsigned s_cnt[256];
  s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kSortTile;
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const long long i = base + j * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&s_cnt[(unsigned)((sort_load_key<kRawIn>(keys, i) >> shift) & 255)], 1u);
  }
  __syncthreads();
  counts[(long long)threadIdx.x * n_tiles + blockIdx.x] = (int)s_cnt[threadIdx.x];
}

// offsets = exclusive prefix sum of counts (digit-major), i.e. the first output slot of (value, tile)
template <bool kRawIn, bool kRawOut>
__global__ void __launch_bounds__(kSortThreads)
sort_scatter_kernel(const unsigned long long* __restrict__ keys_in, const unsigned char* __restrict__ vals_in, long long n,
                    int shift, int n_tiles, const int* __restrict__ offsets, unsigned long long* __restrict__ keys_out,
                    unsigned char* __restrict__ vals_out) {
  __shared__ unsigned s_warp_cnt[kSortWarps][256];
  __shared__ int s_excl[256];
  __shared__ int s_gbase[256];
  __shared__ int s_scan[kSortWarps];
  __shared__ unsigned long long s_key[kSortTile];
  __shared__ unsigned char s_val[kSortTile];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (int j = threadIdx.x; j < kSortWarps * 256; j += kSortThreads) (&s_warp_cnt[0][0])[j] = 0;
  __syncthreads();

  const long long base = (long long)blockIdx.x * kSortTile;
  unsigned long long key[kSortItems];
  unsigned char val[kSortItems];
  unsigned rank[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const long long i = base + warp * kSortWarpSpan + r * 32 + lane;
    const bool live = i < n;
    // padding sorts after every live key of the tile (largest value, latest position) and is never written
    key[r] = live ? sort_load_key<kRawIn>(keys_in, i) : ~0ull;
    val[r] = live ? vals_in[i] : (unsigned char)0;
    const unsigned d = (unsigned)((key[r] >> shift) & 255);
    const unsigned same = __match_any_sync(0xffffffffu, d);
    const unsigned before = s_warp_cnt[warp][d];
    __syncwarp();
    rank[r] = before + __popc(same & lt);
    if ((same & lt) == 0) s_warp_cnt[warp][d] = before + __popc(same);
    __syncwarp();
  }
  __syncthreads();
  {
    // thread t owns digit value t: turn the per-warp counts into per-warp offsets, then scan the tile's counts
    int run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const int c = (int)s_warp_cnt[w][threadIdx.x];
      s_warp_cnt[w][threadIdx.x] = (unsigned)run;
      run += c;
    }
    int tot;
    const int ex = block_exclusive_256(run, s_scan, &tot);
    s_excl[threadIdx.x] = ex;
    s_gbase[threadIdx.x] = offsets[(long long)threadIdx.x * n_tiles + blockIdx.x] - ex;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const unsigned d = (unsigned)((key[r] >> shift) & 255);
    const int p = s_excl[d] + (int)s_warp_cnt[warp][d] + (int)rank[r];
    s_key[p] = key[r];
    s_val[p] = val[r];
  }
  __syncthreads();
  const int live = (int)((n - base) < (long long)kSortTile ? (n - base) : (long long)kSortTile);
  for (int p = threadIdx.x; p < live; p += kSortThreads) {
    const unsigned long long k = s_key[p];
    const long long dst = (long long)s_gbase[(unsigned)((k >> shift) & 255)] + p;
    keys_out[dst] = kRawOut ? (unsigned long long)__double_as_longlong(ordered_to_f64(k)) : k;
    vals_out[dst] = s_val[p];
  }
}

struct SortPlan {
  size_t bits_off, counts_off, scan_off, key_tmp_off, val_tmp_off, total;
  int n_tiles;
};

inline SortPlan sort_plan(long long n) {
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  SortPlan p;
  p.n_tiles = (int)((n + kSortTile - 1) / kSortTile);
  size_t o = 0;
  p.bits_off = o; o += up(2 * sizeof(unsigned long long));
  p.counts_off = o; o += up((size_t)256 * p.n_tiles * sizeof(int));
  p.scan_off = o; o += up(scan_workspace_bytes((long long)256 * p.n_tiles));
  p.key_tmp_off = o; o += up((size_t)n * 8);
  p.val_tmp_off = o; o += up((size_t)n);
  p.total = o;
  return p;
}

// Stable ascending sort of (keys_in[i], vals_in[i]) into (keys_out, vals_out); n < 2^31.  Synchronises the stream
// (the AND / OR of the keys decide on the host which passes run).  ws: sort_plan(n).total bytes.
inline cudaError_t sort_pairs_f64_u8(const double* keys_in, const unsigned char* vals_in, long long n, double* keys_out,
                                     unsigned char* vals_out, unsigned char* ws, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const SortPlan plan = sort_plan(n);
  unsigned long long* bits = (unsigned long long*)(ws + plan.bits_off);
  int* counts = (int*)(ws + plan.counts_off);
  void* scan_ws = ws + plan.scan_off;
  unsigned long long* key_tmp = (unsigned long long*)(ws + plan.key_tmp_off);
  unsigned char* val_tmp = ws + plan.val_tmp_off;
  cudaError_t e;
  if ((e = cudaMemsetAsync(bits, 0xff, sizeof(unsigned long long), stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(bits + 1, 0, sizeof(unsigned long long), stream)) != cudaSuccess) return e;
  const long long want = (n + kSortThreads - 1) / kSortThreads;
  const long long cap = (long long)num_sms() * 8;
  sort_and_or_kernel<<<(unsigned)(want < cap ? want : cap), kSortThreads, 0, stream>>>(keys_in, n, bits);
  note_launch();
  unsigned long long h[2];
  if ((e = cudaMemcpyAsync(h, bits, sizeof(h), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
  int digits[8], nd = 0;
  for (int d = 0; d < 8; ++d)
    if (((h[0] ^ h[1]) >> (8 * d)) & 255) digits[nd++] = d;
  if (nd == 0) {                                   // every key is the same value: the input order is the sorted order
    if ((e = cudaMemcpyAsync(keys_out, keys_in, (size_t)n * 8, cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
    return cudaMemcpyAsync(vals_out, vals_in, (size_t)n, cudaMemcpyDeviceToDevice, stream);
  }
  // ping-pong so that the last pass lands in the caller's output
  const unsigned long long* src_k = (const unsigned long long*)keys_in;
  const unsigned char* src_v = vals_in;
  for (int pass = 0; pass < nd; ++pass) {
    const bool first = pass == 0, last = pass == nd - 1;
    const bool to_out = ((nd - 1 - pass) % 2) == 0;
    unsigned long long* dst_k = to_out ? (unsigned long long*)keys_out : key_tmp;
    unsigned char* dst_v = to_out ? vals_out : val_tmp;
    const int shift = 8 * digits[pass];
    if (first) sort_tile_counts_kernel<true><<<plan.n_tiles, kSortThreads, 0, stream>>>(src_k, n, shift, plan.n_tiles, counts);
    else sort_tile_counts_kernel<false><<<plan.n_tiles, kSortThreads, 0, stream>>>(src_k, n, shift, plan.n_tiles, counts);
    note_launch();
    if ((e = prefix_sum_i32(counts, counts, (long long)256 * plan.n_tiles, false, scan_ws, stream)) != cudaSuccess) return e;
#define CCAL_SORT_SCATTER(RI, RO) \
    sort_scatter_kernel<RI, RO><<<plan.n_tiles, kSortThreads, 0, stream>>>(src_k, src_v, n, shift, plan.n_tiles, counts, dst_k, dst_v)
    if (first && last) CCAL_SORT_SCATTER(true, true);
    else if (first) CCAL_SORT_SCATTER(true, false);
    else if (last) CCAL_SORT_SCATTER(false, true);
    else CCAL_SORT_SCATTER(false, false);
#undef CCAL_SORT_SCATTER
    note_launch();
    src_k = dst_k;
    src_v = dst_v;
  }
  return cudaGetLastError();
}

}  // namespace ccal
