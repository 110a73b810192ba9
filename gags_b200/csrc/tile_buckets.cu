// tile_buckets.cu — K4-K6 as a tile-bucketed segmented sort (the "warp/block-level radix sort with
// a global bucketing pass" of north_star), bit-identical in its outputs to the global path of
// tiles.cu (gsplat isect_tiles + cub::DeviceRadixSort::SortPairs + isect_offset_encode, reached
// from /root/reference/gaussian_renderer/__init__.py:56-70).
//
// The reference order is a stable sort of (tile << 32 | depth_bits) over intersections emitted in
// ascending Gaussian index, i.e. inside a tile: ascending depth bits, ties by ascending Gaussian
// index.  A Gaussian meets a tile at most once, so (depth_bits, gaussian) is a unique key inside a
// tile and the same order is obtained without any global sort:
//   1. tile_scatter  slot = cursor[t]++ (L2 atomic) ; bucket[t][slot] = (depth bits, gaussian) — a
//                    fixed-capacity slab per tile, so no histogram pass has to come first
//   2. bucket_scan   offsets[t] = exclusive sum of the counts -> this IS isect_offsets; n_isects and
//                    the largest bucket come out as two extra ints (one host readback, as before)
//   3. bucket_sort   one CTA per tile: cub::BlockRadixSort of (32-bit depth key, Gaussian value) in
//                    shared memory, 6 passes of 6 bits, then runs of bit-identical depths are put
//                    in Gaussian order; writes flatten_ids (and isect_ids for the info dict)
// HBM traffic: 24 B/Gaussian + 8 B/intersection written, read, then 12 B written — about a sixth
// of the 6-pass global radix sort.  Buckets larger than BUCKET_MAX make the caller fall back
// to the global path.
#include "common.cuh"
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

namespace {

constexpr int SORT_THREADS = 256;
constexpr int BUCKET_MAX = SORT_THREADS * 16;

// steps 1+3 fused: slot = cursor[t]++ ; the tile's bucket is the fixed-capacity slab
// bucket[t * BUCKET_MAX ...] (entries beyond the capacity are dropped; the caller sees the count and
// falls back to the global sort), so no histogram pass is needed before the scatter.
//
// The kernel is bound by the latency of the returning atomic (ncu: 84 % of the stall samples wait
// for it), so the (Gaussian, tile) pairs of a warp are spread evenly over its lanes — a warp scan
// of the per-Gaussian tile counts, then lane l takes pairs l, l + 32, ... and finds each pair's
// Gaussian by a 5-step search over the scanned counts — and every lane keeps SCATTER_ILP atomics in
// flight before it touches the first result.  A warp with 32 small footprints thus needs
// sum/(32 SCATTER_ILP) round trips instead of max(count).
constexpr int SCATTER_ILP = 4;
__global__ void __launch_bounds__(256)
tile_scatter_kernel(const float2 *__restrict__ means2d, const int *__restrict__ radii,
                    const float *__restrict__ depths, long long N, int tile_w, int tile_h,
                    int *__restrict__ cursor, uint2 *__restrict__ bucket) {
  constexpr unsigned FULL = 0xffffffffu;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int x0 = 0, nx = 0, y0 = 0, cnt = 0;
  unsigned dbits = 0u;
  if (i < N && radii[i] > 0) {
    const float2 m = means2d[i];
    int x1, y1;
    tile_bounds(m.x, m.y, radii[i], tile_w, tile_h, x0, x1, y0, y1);
    nx = x1 - x0;
    cnt = nx * (y1 - y0);
    dbits = __float_as_uint(__ldg(depths + i));
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(FULL, incl, 31);
  const int excl = incl - cnt;
  const unsigned gid0 = (unsigned)(i - lane);
  for (int k0 = 0; k0 < total; k0 += 32 * SCATTER_ILP) {
    int tile[SCATTER_ILP], slot[SCATTER_ILP];
    uint2 rec[SCATTER_ILP];
#pragma unroll
    for (int u = 0; u < SCATTER_ILP; ++u) {
      const int k = k0 + u * 32 + lane;
      // owner = the last lane whose exclusive count is <= k (lanes without tiles never win: the
      // next lane with tiles has the same exclusive count and a larger index)
      int src = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int e = __shfl_sync(FULL, excl, (src + step) & 31);
        if (e <= k) src += step;
      }
      const int local = k - __shfl_sync(FULL, excl, src);
      const int ox0 = __shfl_sync(FULL, x0, src), oy0 = __shfl_sync(FULL, y0, src);
      const int onx = max(1, __shfl_sync(FULL, nx, src));
      rec[u] = make_uint2(__shfl_sync(FULL, dbits, src), gid0 + (unsigned)src);
      const int row = local / onx;
      tile[u] = k < total ? (oy0 + row) * tile_w + ox0 + (local - row * onx) : -1;
    }
#pragma unroll
    for (int u = 0; u < SCATTER_ILP; ++u) slot[u] = tile[u] >= 0 ? atomicAdd(cursor + tile[u], 1) : 0;
#pragma unroll
    for (int u = 0; u < SCATTER_ILP; ++u)
      if (tile[u] >= 0 && slot[u] < BUCKET_MAX) bucket[(size_t)tile[u] * BUCKET_MAX + slot[u]] = rec[u];
  }
}

// single CTA: exclusive scan of count[n_tiles] -> offsets[n_tiles + 1]; stats = {n_isects, max}
__global__ void __launch_bounds__(1024)
bucket_scan_kernel(const int *__restrict__ count, int n_tiles, int *__restrict__ offsets,
                   int *__restrict__ stats) {
  using Scan = cub::BlockScan<int, 1024>;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ int s_max[32];
  int base = 0, mx = 0;
  for (int t0 = 0; t0 < n_tiles; t0 += 1024) {
    const int t = t0 + threadIdx.x;
    const int c = t < n_tiles ? count[t] : 0;
    mx = max(mx, c);
    int ex, total;
    Scan(tmp).ExclusiveSum(c, ex, total);
    if (t < n_tiles) offsets[t] = base + ex;
    base += total;
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = 0;
    for (int w = 0; w < 32; ++w) m = max(m, s_max[w]);
    offsets[n_tiles] = base;
    stats[0] = base;
    stats[1] = m;
  }
}

// One CTA sorts one tile's bucket in shared memory: radix sort of (32-bit depth key, Gaussian id
// value) — 6 passes of 6 bits — then the rare runs of bit-identical depths are put in ascending
// Gaussian order (the reference's stable sort emits ties in Gaussian order).
template <int ITEMS>
__device__ __forceinline__ void sort_bucket(const uint2 *__restrict__ src, int cnt, int tile,
                                            long long *__restrict__ keys_out,
                                            int *__restrict__ ids_out, unsigned char *smem,
                                            unsigned *s_u) {
  using Sort = cub::BlockRadixSort<unsigned, SORT_THREADS, ITEMS, unsigned, 6>;
  typename Sort::TempStorage &tmp = *reinterpret_cast<typename Sort::TempStorage *>(smem);
  unsigned k[ITEMS], v[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int idx = j * SORT_THREADS + threadIdx.x;          // striped load (coalesced)
    const uint2 e = idx < cnt ? src[idx] : make_uint2(0xffffffffu, 0xffffffffu);
    k[j] = e.x;
    v[j] = e.y;
  }
  // Sort key = depth bits minus the tile's smallest key: the depths of one tile span ~25 of the 32
  // bits, i.e. 5 radix passes instead of 6.  Padding entries become range + 1 and stay last.
  unsigned kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
    if (j * SORT_THREADS + (int)threadIdx.x < cnt) { kmin = min(kmin, k[j]); kmax = max(kmax, k[j]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&s_u[1], kmin); atomicMax(&s_u[2], kmax); }
  __syncthreads();
  kmin = s_u[1];
  const unsigned pad = s_u[2] - kmin + 1u;                    // > every real key; no overflow: the
#pragma unroll                                                // keys are finite positive floats
  for (int j = 0; j < ITEMS; ++j)
    k[j] = (j * SORT_THREADS + (int)threadIdx.x < cnt) ? k[j] - kmin : pad;
  Sort(tmp).SortBlockedToStriped(k, v, 0, 32 - __clz(pad));    // result: striped, ascending depth
  __syncthreads();
  // sorted lists in shared memory (aliases the sort's storage): tie repair + coalesced output
  unsigned *sk = reinterpret_cast<unsigned *>(smem);
  unsigned *sv = sk + SORT_THREADS * ITEMS;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int idx = j * SORT_THREADS + threadIdx.x;
    sk[idx] = k[j] + kmin;
    sv[idx] = v[j];
  }
  __syncthreads();
  bool tie = false;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int idx = j * SORT_THREADS + threadIdx.x;
    if (idx > 0 && idx < cnt && sk[idx] == sk[idx - 1]) tie = true;
  }
  if (tie) s_u[0] = 1u;
  __syncthreads();
  if (s_u[0]) {
    if (threadIdx.x == 0) {                                   // rare: order each run of equal depths
      int i = 0;
      while (i < cnt) {
        int j = i + 1;
        while (j < cnt && sk[j] == sk[i]) ++j;
        for (int a = i + 1; a < j; ++a) {                    // insertion sort of sv[i, j)
          const unsigned x = sv[a];
          int b = a - 1;
          while (b >= i && sv[b] > x) { sv[b + 1] = sv[b]; --b; }
          sv[b + 1] = x;
        }
        i = j;
      }
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < cnt; idx += SORT_THREADS) {
    ids_out[idx] = (int)sv[idx];
    if (keys_out) keys_out[idx] = ((long long)tile << 32) | (long long)sk[idx];
  }
}

// Two register classes, two launches over the same grid (a CTA whose tile belongs to the other class
// leaves at once): the 16-items-per-thread variant needs 126 registers, which held every tile to two
// CTAs per SM although the typical tile (~1200 intersections at config 3) sorts with 5-8 per thread.
constexpr int SMALL_ITEMS = 8;
template <bool LARGE>
__global__ void __launch_bounds__(SORT_THREADS, LARGE ? 2 : 3)
bucket_sort_kernel(const uint2 *__restrict__ bucket, const int *__restrict__ offsets, long long cap,
                   long long *__restrict__ isect_ids, int *__restrict__ flatten_ids) {
  constexpr int MAXI = LARGE ? 16 : SMALL_ITEMS;
  using SortMax = cub::BlockRadixSort<unsigned, SORT_THREADS, MAXI, unsigned, 6>;
  constexpr int kSmem = sizeof(typename SortMax::TempStorage) > SORT_THREADS * MAXI * 8
                            ? (int)sizeof(typename SortMax::TempStorage)
                            : SORT_THREADS * MAXI * 8;
  __shared__ __align__(16) unsigned char smem[kSmem];
  __shared__ unsigned s_u[3];                     // tie flag, smallest key, largest key
  const int tile = blockIdx.x;
  const int s = offsets[tile], cnt = offsets[tile + 1] - s;
  if (cnt <= 0) return;
  if ((cnt > SORT_THREADS * SMALL_ITEMS) != LARGE) return;
  // speculative launch (before the host has read the counts): leave tiles alone that do not fit
  // the slab or the output arrays — the caller sees the counts and redoes the view
  if (cnt > BUCKET_MAX || (long long)s + cnt > cap) return;
  if (threadIdx.x == 0) { s_u[0] = 0u; s_u[1] = 0xffffffffu; s_u[2] = 0u; }
  __syncthreads();
  const uint2 *src = bucket + (size_t)tile * BUCKET_MAX;
  long long *ko = isect_ids ? isect_ids + s : nullptr;
  int *io = flatten_ids + s;
  if constexpr (LARGE) {
    if (cnt <= SORT_THREADS * 11) sort_bucket<11>(src, cnt, tile, ko, io, smem, s_u);
    else sort_bucket<16>(src, cnt, tile, ko, io, smem, s_u);
  } else {
    if (cnt <= SORT_THREADS * 2) sort_bucket<2>(src, cnt, tile, ko, io, smem, s_u);
    else if (cnt <= SORT_THREADS * 4) sort_bucket<4>(src, cnt, tile, ko, io, smem, s_u);
    else if (cnt <= SORT_THREADS * 5) sort_bucket<5>(src, cnt, tile, ko, io, smem, s_u);
    else if (cnt <= SORT_THREADS * 6) sort_bucket<6>(src, cnt, tile, ko, io, smem, s_u);
    else sort_bucket<8>(src, cnt, tile, ko, io, smem, s_u);
  }
}

int launch_bucket_sort(const void *bucket, int n_tiles, const int *offsets, long long cap,
                       int64_t *isect_ids, int32_t *flatten_ids, cudaStream_t st) {
  bucket_sort_kernel<false><<<n_tiles, SORT_THREADS, 0, st>>>(
      reinterpret_cast<const uint2 *>(bucket), offsets, cap, reinterpret_cast<long long *>(isect_ids),
      flatten_ids);
  bucket_sort_kernel<true><<<n_tiles, SORT_THREADS, 0, st>>>(
      reinterpret_cast<const uint2 *>(bucket), offsets, cap, reinterpret_cast<long long *>(isect_ids),
      flatten_ids);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int32_t gags_tile_bucket_max(void) { return BUCKET_MAX; }

// Steps 1-2: scatter every (Gaussian, tile) pair into the tile's slab of bucket[n_tiles *
// gags_tile_bucket_max()] (8 B each), count[n_tiles] = pairs per tile (zeroed here), exclusive scan
// -> offsets[n_tiles + 1], stats_dev[2] = {n_isects, largest bucket}.
extern "C" int gags_tile_bucket_count(const float *means2d, const int32_t *radii, const float *depths,
                                      int64_t N, int32_t tile_w, int32_t tile_h, int32_t *count,
                                      void *bucket, int32_t *offsets, int32_t *stats_dev,
                                      void *stream) {
  if (!means2d || !radii || !depths || !count || !bucket || !offsets || !stats_dev || N < 0 ||
      tile_w <= 0 || tile_h <= 0)
    return GAGS_EINVAL;
  if ((long long)tile_w * tile_h > 0x3fffffffLL || N > 0x7fffffffLL) return GAGS_ERANGE;
  if (((uintptr_t)means2d) & 7u || ((uintptr_t)bucket) & 7u) return GAGS_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const int n_tiles = tile_w * tile_h;
  GAGS_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)n_tiles, st));
  if (N > 0) {
    tile_scatter_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float2 *>(means2d), radii, depths, (long long)N, tile_w, tile_h, count,
        reinterpret_cast<uint2 *>(bucket));
    GAGS_CHECK_LAUNCH();
  }
  bucket_scan_kernel<<<1, 1024, 0, st>>>(count, n_tiles, offsets, stats_dev);
  GAGS_CHECK_LAUNCH();
  return 0;
}

// Step 3: sort every bucket by (depth bits, Gaussian index) and write the contiguous lists.
// Requires max_bucket <= gags_tile_bucket_max() (GAGS_ERANGE is the caller's cue to use the global
// sort instead).  isect_ids may be NULL.
extern "C" int gags_tile_bucket_sort(const void *bucket, int32_t tile_w, int32_t tile_h,
                                     const int32_t *offsets, int32_t max_bucket, int64_t *isect_ids,
                                     int32_t *flatten_ids, void *stream) {
  if (!bucket || !offsets || !flatten_ids || tile_w <= 0 || tile_h <= 0) return GAGS_EINVAL;
  if (max_bucket > BUCKET_MAX) return GAGS_ERANGE;
  return launch_bucket_sort(bucket, tile_w * tile_h, offsets, 0x7fffffffffffffffLL, isect_ids,
                            flatten_ids, (cudaStream_t)stream);
}

// The same sort launched BEFORE the host has read stats_dev (so the device sorts while the host
// waits for the counts): `capacity` = elements the output arrays hold.  Tiles that overflow the slab
// or the arrays are skipped on the device; the caller must check stats_dev afterwards and, if
// n_isects > capacity or the largest bucket > gags_tile_bucket_max(), discard the result.
extern "C" int gags_tile_bucket_sort_guarded(const void *bucket, int32_t tile_w, int32_t tile_h,
                                             const int32_t *offsets, int64_t capacity,
                                             int64_t *isect_ids, int32_t *flatten_ids,
                                             void *stream) {
  if (!bucket || !offsets || !flatten_ids || tile_w <= 0 || tile_h <= 0 || capacity < 0)
    return GAGS_EINVAL;
  return launch_bucket_sort(bucket, tile_w * tile_h, offsets, (long long)capacity, isect_ids,
                            flatten_ids, (cudaStream_t)stream);
}
