// common.cuh — shared device helpers for the gags_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gags_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "gags_b200 kernels are written for sm_100a only"
#endif

#define GAGS_ALPHA_MIN (1.0f / 255.0f)
#define GAGS_ALPHA_MAX 0.999f
#define GAGS_T_STOP 1e-4f

#define GAGS_CHECK_LAUNCH()                          \
  do {                                               \
    cudaError_t _e = cudaGetLastError();             \
    if (_e != cudaSuccess) return (int)_e;           \
  } while (0)

#define GAGS_CUDA(expr)                              \
  do {                                               \
    cudaError_t _e = (expr);                         \
    if (_e != cudaSuccess) return (int)_e;           \
  } while (0)

static inline bool gags_aligned16(const void *p) { return (((uintptr_t)p) & 15u) == 0; }

// SM count of the current device (148 on B200), cached per device: grids are sized in multiples of it.
static inline long long gags_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Device-side copy of gags_camera_t plus the tile grid (passed by value as a kernel parameter).
struct CamDev {
  float R[9];
  float t[3];
  float fx, fy, cx, cy;
  int W, H;
  float eps2d, near_plane, far_plane, radius_clip, smod;
  int flags;
  int tile_w, tile_h;
  const float *vm_dev;   // optional device copy of the 4x4 view matrix
};

// resolve R,t from the device view matrix when one was given (uniform broadcast loads)
__device__ __forceinline__ void cam_resolve(CamDev &c) {
  if (c.vm_dev) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) c.R[i * 3 + j] = __ldg(c.vm_dev + i * 4 + j);
      c.t[i] = __ldg(c.vm_dev + i * 4 + 3);
    }
  }
}

static inline CamDev make_camdev(const gags_camera_t *c, int tile_w, int tile_h) {
  CamDev d;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) d.R[i * 3 + j] = c->viewmat[i * 4 + j];
    d.t[i] = c->viewmat[i * 4 + 3];
  }
  d.fx = c->fx; d.fy = c->fy; d.cx = c->cx; d.cy = c->cy;
  d.W = c->width; d.H = c->height;
  d.eps2d = c->eps2d; d.near_plane = c->near_plane; d.far_plane = c->far_plane;
  d.radius_clip = c->radius_clip; d.smod = c->scaling_modifier; d.flags = c->flags;
  d.tile_w = tile_w; d.tile_h = tile_h;
  d.vm_dev = c->viewmat_dev;
  return d;
}

// ---- tile bounds: fp32, division by 16 is exact -----------------------------
__device__ __forceinline__ void tile_bounds(float mx, float my, int radius, int tile_w, int tile_h,
                                            int &x0, int &x1, int &y0, int &y1) {
  const float r = (float)radius * (1.0f / GAGS_TILE);
  const float tx = mx * (1.0f / GAGS_TILE);
  const float ty = my * (1.0f / GAGS_TILE);
  // negative -> 0 (float->uint32 conversion saturates), then clamp to the grid
  x0 = min(max(0, (int)fminf(floorf(tx - r), 1e9f)), tile_w);
  x1 = min(max(0, (int)fminf(ceilf(tx + r), 1e9f)), tile_w);
  y0 = min(max(0, (int)fminf(floorf(ty - r), 1e9f)), tile_h);
  y1 = min(max(0, (int)fminf(ceilf(ty + r), 1e9f)), tile_h);
}

// ---- blend weight evaluation ------------------------------------------------
// Returns alpha (0 when the Gaussian is rejected for this pixel: sigma < 0 or alpha < 1/255).
__device__ __forceinline__ float eval_alpha(float mx, float my, float ca, float cb, float cc,
                                            float op, float px, float py) {
  const float dx = mx - px;
  const float dy = my - py;
  const float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
  const float alpha = fminf(GAGS_ALPHA_MAX, op * __expf(-sigma));
  return (sigma < 0.f || alpha < GAGS_ALPHA_MIN) ? 0.f : alpha;
}

// ---- mbarrier + bulk async copy (TMA 1-D, SASS: UBLKCP / SYNCS) --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-B aligned.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                         uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// make generic-proxy smem writes visible to the async proxy (before reusing a TMA buffer)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ float4 ldg_nc4(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_cs4(float4 *p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_cs1(float *p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add4(float *p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
