// blend_tc_common.cuh — pieces shared by the tensor-core forward and feature-backward blend kernels
// (blend_fwd_tc.cu, blend_bwd_tc.cu): the exact tile-level cull, the per-batch Gaussian record and
// the per-pixel weight evaluation that fills one [hi(32) | lo(32)] bf16
// weight row per pixel.
#pragma once
#include "umma.cuh"

constexpr int TC_KB = 32;      // Gaussians per batch (= one 128-B row of [hi|lo] bf16 weights)
constexpr int TC_RING = 256;   // survivor ring capacity (power of two)

// Conservative reach of the alpha >= 1/255 ellipse {sigma <= ln(255 op)} of a projected Gaussian;
// returns false when the Gaussian can never pass the alpha test (op < 1/255).
__device__ __forceinline__ bool tc_alpha_extent(float a, float b, float c, float op, float &hx,
                                                float &hy) {
  const float L = __logf(255.f * op);
  if (!(L > -1e-3f)) return false;
  const float Lm = fmaxf(L, 0.f) + 2e-3f;
  const float det = a * c - b * b;
  if (det > 0.f) {
    const float inv = 2.f * Lm / det;
    hx = sqrtf(inv * c) * 1.0005f + 0.02f;
    hy = sqrtf(inv * a) * 1.0005f + 0.02f;
  } else {
    hx = hy = 1e9f;
  }
  return true;
}

// 4-bit mask of the 8x4 pixel blocks of the 16x8 half tile at (hx0, hy0) (= first pixel CENTRE)
// that the Gaussian's alpha >= 1/255 bounding box touches.
__device__ __forceinline__ unsigned tc_block_mask(const float4 &g0, const float4 &g1, float hx0,
                                                  float hy0) {
  float hx, hy;
  unsigned mask = 0;
  if (tc_alpha_extent(g0.z, g0.w, g1.x, g1.y, hx, hy)) {
    const float lx = g0.x - hx, ux = g0.x + hx, ly = g0.y - hy, uy = g0.y + hy;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float bx = hx0 + (float)((b & 1) << 3), by = hy0 + (float)((b >> 1) << 2);
      if (ux >= bx && lx <= bx + 7.f && uy >= by && ly <= by + 3.f) mask |= 1u << b;
    }
  }
  return mask;
}

// Batch record read by the pixel threads: exponent coefficients pre-scaled by log2(e) so that
//   alpha = op * exp2(A dx^2 + B dx dy + C dy^2),  A = -a/2 log2e, B = -b log2e, C = -c/2 log2e.
// A null record (op = 0) yields alpha = 0 and pads a short batch.
struct TcRec {
  float4 q0;   // mx, my, A, B
  float4 q1;   // C, op, list index (bits), unused
};
__device__ __forceinline__ TcRec tc_make_rec(const float4 &g0, const float4 &g1) {
  constexpr float LOG2E = 1.4426950408889634f;
  TcRec r;
  r.q0 = make_float4(g0.x, g0.y, -0.5f * LOG2E * g0.z, -LOG2E * g0.w);
  r.q1 = make_float4(-0.5f * LOG2E * g1.x, g1.y, g1.z, 0.f);
  return r;
}
__device__ __forceinline__ TcRec tc_null_rec() {
  TcRec r;
  r.q0 = make_float4(0.f, 0.f, 0.f, 0.f);
  r.q1 = make_float4(0.f, 0.f, 0.f, 0.f);
  return r;
}

__device__ __forceinline__ float tc_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct TcPixel {
  float px, py;   // pixel centre
  float T;        // transmittance in front of the next Gaussian
  int last;       // list index of the last contributor
  bool done;      // early-stopped (T' <= 1e-4) or outside the image
};

// Weights of 8 consecutive Gaussians of the batch (records rec0[8], rec1[8] in shared memory) at
// one pixel: 8 independent alpha evaluations, then the sequential transmittance chain written with
// selects only.  Returns the bf16 hi / lo halves packed for one 16-byte chunk each.
__device__ __forceinline__ void tc_weights8(const float4 *__restrict__ rec0,
                                            const float4 *__restrict__ rec1, TcPixel &st, uint4 &hi,
                                            uint4 &lo) {
  float a[8];
  int gi[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 r0 = rec0[k];
    const float4 r1 = rec1[k];
    const float dx = r0.x - st.px, dy = r0.y - st.py;
    const float t = fmaf(r0.w, dy, r0.z * dx);            // A dx + B dy
    const float u = (r1.x * dy) * dy;                     // C dy^2
    const float q = fmaf(t, dx, u);                       // = -sigma * log2(e)
    const float al = fminf(GAGS_ALPHA_MAX, r1.y * tc_ex2(q));
    a[k] = (q <= 0.f && al >= GAGS_ALPHA_MIN) ? al : 0.f;
    gi[k] = __float_as_int(r1.z);
  }
  float w[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float ae = st.done ? 0.f : a[k];
    const float Tn = fmaf(-ae, st.T, st.T);               // T (1 - alpha); == T when alpha == 0
    const bool stopnow = Tn <= GAGS_T_STOP;               // T > 1e-4 is invariant while !done
    const float wk = stopnow ? 0.f : ae * st.T;
    st.T = stopnow ? st.T : Tn;
    st.done = st.done || stopnow;
    st.last = (wk > 0.f) ? gi[k] : st.last;
    w[k] = wk;
  }
  split_pack2(w[0], w[1], hi.x, lo.x);
  split_pack2(w[2], w[3], hi.y, lo.y);
  split_pack2(w[4], w[5], hi.z, lo.z);
  split_pack2(w[6], w[7], hi.w, lo.w);
}
