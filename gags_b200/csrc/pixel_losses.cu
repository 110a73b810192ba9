// pixel_losses.cu — the per-pixel (reduce-over-channels) losses of the reference on the channel-last
// raster:  l1_loss_map  (/root/reference/utils/loss_utils.py:23-24: mean |a - b| over dim 0 -> [H,W],
// used by the L_r-distill branch, /root/reference/train.py:165-166)  and  cos_loss (:29-30:
// 1 - mean_px cosine_similarity(a, b, dim=0)).  With [H*W, D] rows the channel reduction is a
// contiguous row reduction: one warp per pixel, 16-byte loads, one pass forward and one backward
// instead of the 3-4 elementwise passes + strided reduction of the eager form.
//   mode 0 (L1 map): out[p] = (1/D) sum_c |a - b|
//   mode 1 (cosine): out[p] = sum_c (a / max(|a|, eps)) (b / max(|b|, eps)),  stats[p] = {|a|, |b|}
// Backward (w.r.t. a; b is the target): mode 0  v_a = g[p]/D * sign(a - b);
//   mode 1  v_a = g[p] * (b / (na nb) - out[p] * a / na^2)  (0 where |a| <= eps, as autograd's clamp).
#include "common.cuh"

namespace {

constexpr float COS_EPS = 1e-8f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(256)
pixel_loss_fwd(const float4 *__restrict__ a, const float4 *__restrict__ b, long long hw, int d4,
               float *__restrict__ out, float2 *__restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < hw; p += warps) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < d4; c += 32) {
      const float4 x = ldg_nc4(a + p * d4 + c), y = ldg_nc4(b + p * d4 + c);
      if (MODE == 0) {
        s0 += fabsf(x.x - y.x) + fabsf(x.y - y.y) + fabsf(x.z - y.z) + fabsf(x.w - y.w);
      } else {
        s0 += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
        s1 += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        s2 += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
      }
    }
    s0 = warp_sum(s0);
    if (MODE == 1) { s1 = warp_sum(s1); s2 = warp_sum(s2); }
    if (lane == 0) {
      if (MODE == 0) {
        out[p] = s0 / (float)(4 * d4);
      } else {
        const float na = sqrtf(s1), nb = sqrtf(s2);
        out[p] = s0 / (fmaxf(na, COS_EPS) * fmaxf(nb, COS_EPS));
        stats[p] = make_float2(na, nb);
      }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256)
pixel_loss_bwd(const float4 *__restrict__ a, const float4 *__restrict__ b,
               const float *__restrict__ g, const float *__restrict__ out,
               const float2 *__restrict__ stats, long long n4, int d4, float4 *__restrict__ va) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const long long p = i / d4;
    const float4 x = ldg_nc4(a + i), y = ldg_nc4(b + i);
    const float gp = __ldg(g + p);
    float4 v;
    if (MODE == 0) {
      const float s = gp / (float)(4 * d4);
      v.x = x.x > y.x ? s : (x.x < y.x ? -s : 0.f);
      v.y = x.y > y.y ? s : (x.y < y.y ? -s : 0.f);
      v.z = x.z > y.z ? s : (x.z < y.z ? -s : 0.f);
      v.w = x.w > y.w ? s : (x.w < y.w ? -s : 0.f);
    } else {
      const float2 st = __ldg(stats + p);
      const float na = fmaxf(st.x, COS_EPS), nb = fmaxf(st.y, COS_EPS);
      const float k1 = gp / (na * nb);
      // d/da of a / max(|a|, eps): the clamp has zero derivative below eps
      const float k2 = st.x > COS_EPS ? gp * __ldg(out + p) / (na * na) : 0.f;
      v.x = k1 * y.x - k2 * x.x; v.y = k1 * y.y - k2 * x.y;
      v.z = k1 * y.z - k2 * x.z; v.w = k1 * y.w - k2 * x.w;
    }
    va[i] = v;
  }
}

}  // namespace

extern "C" int gags_pixel_loss_fwd(int32_t mode, const float *a, const float *b, int64_t HW,
                                   int32_t D, float *out, float *stats, void *stream) {
  if (!a || !b || !out || HW < 0 || D < 4 || (D & 3) || (mode != 0 && mode != 1)) return GAGS_EINVAL;
  if (mode == 1 && !stats) return GAGS_EINVAL;
  if (!gags_aligned16(a) || !gags_aligned16(b)) return GAGS_EALIGN;
  if (HW == 0) return 0;
  long long blocks = (HW + 7) / 8;                   // 8 warps = 8 pixels per CTA
  if (blocks > gags_sm_count() * 8) blocks = gags_sm_count() * 8;
  const float4 *a4 = reinterpret_cast<const float4 *>(a), *b4 = reinterpret_cast<const float4 *>(b);
  if (mode == 0)
    pixel_loss_fwd<0><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a4, b4, HW, D / 4, out,
                                                                         nullptr);
  else
    pixel_loss_fwd<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        a4, b4, HW, D / 4, out, reinterpret_cast<float2 *>(stats));
  GAGS_CHECK_LAUNCH();
  return 0;
}

extern "C" int gags_pixel_loss_bwd(int32_t mode, const float *a, const float *b, const float *g,
                                   const float *out, const float *stats, int64_t HW, int32_t D,
                                   float *v_a, void *stream) {
  if (!a || !b || !g || !v_a || HW < 0 || D < 4 || (D & 3) || (mode != 0 && mode != 1))
    return GAGS_EINVAL;
  if (mode == 1 && (!out || !stats)) return GAGS_EINVAL;
  if (!gags_aligned16(a) || !gags_aligned16(b) || !gags_aligned16(v_a)) return GAGS_EALIGN;
  if (HW == 0) return 0;
  const long long n4 = (long long)HW * (D / 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > gags_sm_count() * 8) blocks = gags_sm_count() * 8;
  const float4 *a4 = reinterpret_cast<const float4 *>(a), *b4 = reinterpret_cast<const float4 *>(b);
  if (mode == 0)
    pixel_loss_bwd<0><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        a4, b4, g, out, nullptr, n4, D / 4, reinterpret_cast<float4 *>(v_a));
  else
    pixel_loss_bwd<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        a4, b4, g, out, reinterpret_cast<const float2 *>(stats), n4, D / 4,
        reinterpret_cast<float4 *>(v_a));
  GAGS_CHECK_LAUNCH();
  return 0;
}
