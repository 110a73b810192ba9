// tiles.cu — K4 tile intersection (count / scan / emit), K5 radix sort, K6 tile offsets.
// Replaces gsplat isect_tiles + cub::DeviceRadixSort::SortPairs + isect_offset_encode reached from
// /root/reference/gaussian_renderer/__init__.py:56-70; integer semantics = SURVEY.md App. A.3/A.4.
// Integer stages: results are bit-exact functions of (means2d, radii, depths).
//
// Roofline: HBM.  Algorithmic bytes: count 12 B/Gaussian in + 4 out; emit 20 B/Gaussian in +
// 12 B/intersection out; sort 6 passes x 24 B/intersection (46-bit keys at 1080p); offsets 8 B/isect.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace {

__global__ void tile_count_kernel(const float2 *__restrict__ means2d, const int *__restrict__ radii,
                                  long long N, int tile_w, int tile_h, int *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int r = radii[i];
  int n = 0;
  if (r > 0) {
    const float2 m = means2d[i];
    int x0, x1, y0, y1;
    tile_bounds(m.x, m.y, r, tile_w, tile_h, x0, x1, y0, y1);
    n = (x1 - x0) * (y1 - y0);
  }
  out[i] = n;
}

__global__ void write_total_kernel(const int *__restrict__ cum, long long N, int *__restrict__ total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *total = cum[N - 1];
}

// One thread per Gaussian; Gaussians touching many tiles are emitted cooperatively by the warp so
// that a screen-filling Gaussian does not serialise thousands of stores in one lane.
constexpr int EMIT_COOP = 32;

__global__ void __launch_bounds__(256)
tile_emit_kernel(const float2 *__restrict__ means2d, const int *__restrict__ radii,
                 const float *__restrict__ depths, const int *__restrict__ cum, long long N,
                 int tile_w, int tile_h, long long *__restrict__ keys, int *__restrict__ vals) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int x0 = 0, x1 = 0, y0 = 0, y1 = 0, cnt = 0, start = 0;
  unsigned dbits = 0;
  if (i < N && radii[i] > 0) {
    const float2 m = means2d[i];
    tile_bounds(m.x, m.y, radii[i], tile_w, tile_h, x0, x1, y0, y1);
    cnt = (x1 - x0) * (y1 - y0);
    start = cum[i] - cnt;
    dbits = __float_as_uint(depths[i]);
  }
  if (cnt > 0 && cnt < EMIT_COOP) {
    int k = start;
    for (int ty = y0; ty < y1; ++ty)
      for (int tx = x0; tx < x1; ++tx, ++k) {
        keys[k] = ((long long)(ty * tile_w + tx) << 32) | (long long)dbits;
        vals[k] = (int)i;
      }
  }
  unsigned big = __ballot_sync(0xffffffffu, cnt >= EMIT_COOP);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int bx0 = __shfl_sync(0xffffffffu, x0, src);
    const int bx1 = __shfl_sync(0xffffffffu, x1, src);
    const int by0 = __shfl_sync(0xffffffffu, y0, src);
    const int bcnt = __shfl_sync(0xffffffffu, cnt, src);
    const int bstart = __shfl_sync(0xffffffffu, start, src);
    const unsigned bd = __shfl_sync(0xffffffffu, dbits, src);
    const int gid = (int)(i - lane + src);
    const int nx = bx1 - bx0;
    for (int k = lane; k < bcnt; k += 32) {
      const int ty = by0 + k / nx, tx = bx0 + k % nx;
      keys[bstart + k] = ((long long)(ty * tile_w + tx) << 32) | (long long)bd;
      vals[bstart + k] = gid;
    }
  }
}

__global__ void tile_offsets_kernel(const long long *__restrict__ keys, long long n, int n_tiles,
                                    int *__restrict__ offsets) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0) {
    if (i <= n_tiles) offsets[i] = 0;
    return;
  }
  if (i >= n) return;
  const int cur = (int)(keys[i] >> 32);
  if (i == 0) {
    for (int t = 0; t <= cur; ++t) offsets[t] = 0;
  } else {
    const int prev = (int)(keys[i - 1] >> 32);
    for (int t = prev + 1; t <= cur; ++t) offsets[t] = (int)i;
  }
  if (i == n - 1) {
    for (int t = cur + 1; t <= n_tiles; ++t) offsets[t] = (int)n;
  }
}

}  // namespace

extern "C" int gags_tile_count(const float *means2d, const int32_t *radii, int64_t N,
                               int32_t tile_w, int32_t tile_h, int32_t *tiles_touched,
                               void *stream) {
  if (!means2d || !radii || !tiles_touched || N < 0 || tile_w <= 0 || tile_h <= 0) return GAGS_EINVAL;
  if (N == 0) return 0;
  if (((uintptr_t)means2d) & 7u) return GAGS_EALIGN;
  tile_count_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2 *>(means2d), radii, (long long)N, tile_w, tile_h, tiles_touched);
  GAGS_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t gags_tile_scan_workspace_bytes(int64_t N) {
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, bytes, (const int *)nullptr, (int *)nullptr, (int)N);
  return bytes + 256;
}

extern "C" int gags_tile_scan(const int32_t *tiles_touched, int64_t N, int32_t *cum_tiles,
                              int32_t *n_isects_dev, void *workspace, size_t workspace_bytes,
                              void *stream) {
  if (!tiles_touched || !cum_tiles || !n_isects_dev || N < 0 || N > 0x7fffffffLL) return GAGS_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    GAGS_CUDA(cudaMemsetAsync(n_isects_dev, 0, sizeof(int), st));
    return 0;
  }
  size_t need = 0;
  cub::DeviceScan::InclusiveSum(nullptr, need, tiles_touched, cum_tiles, (int)N);
  if (!workspace || workspace_bytes < need) return GAGS_ESMALL;
  GAGS_CUDA(cub::DeviceScan::InclusiveSum(workspace, workspace_bytes, tiles_touched, cum_tiles,
                                          (int)N, st));
  write_total_kernel<<<1, 32, 0, st>>>(cum_tiles, (long long)N, n_isects_dev);
  GAGS_CHECK_LAUNCH();
  return 0;
}

extern "C" int gags_tile_emit(const float *means2d, const int32_t *radii, const float *depths,
                              const int32_t *cum_tiles, int64_t N, int32_t tile_w, int32_t tile_h,
                              int64_t *isect_ids, int32_t *flatten_ids, void *stream) {
  if (!means2d || !radii || !depths || !cum_tiles || N < 0 || tile_w <= 0 || tile_h <= 0)
    return GAGS_EINVAL;
  if (N == 0) return 0;
  if (!isect_ids || !flatten_ids) return GAGS_EINVAL;
  if ((long long)tile_w * tile_h > 0x7fffffffLL) return GAGS_ERANGE;
  tile_emit_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2 *>(means2d), radii, depths, cum_tiles, (long long)N, tile_w,
      tile_h, reinterpret_cast<long long *>(isect_ids), flatten_ids);
  GAGS_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t gags_sort_pairs_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  size_t bytes = 0;
  cub::DoubleBuffer<long long> k(nullptr, nullptr);
  cub::DoubleBuffer<int> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int)n, 0, 64);
  return bytes + 256;
}

extern "C" int gags_sort_pairs(int64_t *keys_a, int64_t *keys_b, int32_t *vals_a, int32_t *vals_b,
                               int64_t n, int32_t end_bit, void *workspace, size_t workspace_bytes,
                               int32_t *selector_host, void *stream) {
  if (!selector_host || n < 0 || n > 0x7fffffffLL || end_bit < 1 || end_bit > 64) return GAGS_EINVAL;
  *selector_host = 0;
  if (n == 0) return 0;
  if (!keys_a || !keys_b || !vals_a || !vals_b) return GAGS_EINVAL;
  cub::DoubleBuffer<long long> k(reinterpret_cast<long long *>(keys_a),
                                 reinterpret_cast<long long *>(keys_b));
  cub::DoubleBuffer<int> v(vals_a, vals_b);
  size_t need = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, need, k, v, (int)n, 0, end_bit);
  if (!workspace || workspace_bytes < need) return GAGS_ESMALL;
  GAGS_CUDA(cub::DeviceRadixSort::SortPairs(workspace, workspace_bytes, k, v, (int)n, 0, end_bit,
                                            (cudaStream_t)stream));
  *selector_host = (k.Current() == reinterpret_cast<long long *>(keys_a)) ? 0 : 1;
  return 0;
}

extern "C" int gags_tile_offsets(const int64_t *isect_ids_sorted, int64_t n, int32_t n_tiles,
                                 int32_t *offsets, void *stream) {
  if (!offsets || n < 0 || n_tiles <= 0 || n > 0x7fffffffLL) return GAGS_EINVAL;
  if (n > 0 && !isect_ids_sorted) return GAGS_EINVAL;
  const long long work = n > 0 ? n : (long long)n_tiles + 1;
  tile_offsets_kernel<<<(unsigned)((work + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long *>(isect_ids_sorted), (long long)n, n_tiles, offsets);
  GAGS_CHECK_LAUNCH();
  return 0;
}
