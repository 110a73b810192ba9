// blend_common.cuh — pieces shared by the forward and backward wide-feature blend kernels.
//
// Work decomposition (wide path, D > 32): one CTA = one 16x8 "half tile" (128 pixels) x up to 256
// channels.  The tile's depth-sorted Gaussian list is walked in batches of 32:
//   phase A1  all 256 threads evaluate alpha(g, px) for the batch           -> wbuf[g][px]
//   phase A2  128 pixel threads run the sequential transmittance chain      -> wbuf = w = alpha*T
//             and publish per-(g, 32-pixel warp) contribution masks (ballots)
//   compact   warp 0 drops Gaussians that contribute to no pixel of the half tile and issues one
//             bulk async copy (TMA 1-D, UBLKCP) per surviving feature row
//   phase B   (one batch behind, overlapping the copies) register-tiled accumulate.
// Pixel index p in [0,128): 4x4 blocks, block b = p>>4 (4 across, 2 down), q = p&15 row-major in
// the block, so each accumulate warp owns one 4x4 block and can skip Gaussians that miss it.
#pragma once
#include "common.cuh"

constexpr int WB = 32;    // Gaussians per batch (one per lane of the copy-issuing warp)
constexpr int HP = 128;   // pixels per CTA
constexpr int HROWS = 8;  // pixel rows per CTA

__device__ __forceinline__ void hp_pixel(int p, int &dx, int &dy) {
  const int b = p >> 4, q = p & 15;
  dx = ((b & 3) << 2) + (q & 3);
  dy = ((b >> 2) << 2) + (q >> 2);
}

struct PixelState {
  float T;
  int last;
  int done;
};

// Phase A1: alpha for 16 Gaussians of the batch at this thread's pixel.
__device__ __forceinline__ void phase_a1(const float4 *sg0, const float4 *sg1, int nb, int half,
                                         int p, float px, float py, bool inside, float *wrow) {
  const int g0 = half * (WB / 2);
#pragma unroll 4
  for (int k = 0; k < WB / 2; ++k) {
    const int g = g0 + k;
    float a = 0.f;
    if (g < nb && inside) {
      const float4 r0 = sg0[g];
      const float4 r1 = sg1[g];
      a = eval_alpha(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, px, py);
    }
    wrow[g * HP + p] = a;
  }
}

// Phase A2 (threads with tid < 128): sequential chain; overwrites alpha with w = alpha * T and
// writes the ballot of "w > 0" per Gaussian for this 32-pixel warp.
__device__ __forceinline__ void phase_a2(float *wrow, unsigned *masks, int nb, int p, int base_idx,
                                         PixelState &st) {
  const int pw = p >> 5;
  const int lane = p & 31;
  for (int g = 0; g < nb; ++g) {
    const float a = wrow[g * HP + p];
    float w = 0.f;
    if (!st.done && a > 0.f) {
      const float Tn = st.T * (1.0f - a);
      if (Tn <= GAGS_T_STOP) {
        st.done = 1;
      } else {
        w = a * st.T;
        st.T = Tn;
        st.last = base_idx + g;
      }
    }
    wrow[g * HP + p] = w;
    const unsigned m = __ballot_sync(0xffffffffu, w > 0.f);
    if (lane == 0) masks[g * 4 + pw] = m;
  }
}
