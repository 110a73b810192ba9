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

// Exact (conservative only by the fp32 slack) test "does the ellipse {sigma <= Lm} reach the
// rectangle of pixel centres [x_lo, x_hi] x [y_lo, y_hi]": sigma is a convex quadratic, so its minimum
// over the rectangle is 0 when the mean is inside and otherwise sits on one of the four edges, where
// it is a clamped 1-D minimisation.  Culls ~15 % of the bounding-box survivors at config 3
// (tools/workload_stats.py): diagonal / elongated footprints whose box clips a corner.
__device__ __forceinline__ bool tc_ellipse_hits_rect(float mx, float my, float a, float b, float c,
                                                     float Lm, float x_lo, float x_hi, float y_lo,
                                                     float y_hi) {
  if (mx >= x_lo && mx <= x_hi && my >= y_lo && my <= y_hi) return true;
  if (!(a > 0.f && c > 0.f)) return true;
  float best = 3.0e38f;
  const float ia = __fdividef(1.f, a), ic = __fdividef(1.f, c);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float dx = (k ? x_hi : x_lo) - mx;
    const float dy = fminf(fmaxf(-b * dx * ic, y_lo - my), y_hi - my);
    best = fminf(best, 0.5f * (a * dx * dx + c * dy * dy) + b * dx * dy);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float dy = (k ? y_hi : y_lo) - my;
    const float dx = fminf(fmaxf(-b * dy * ia, x_lo - mx), x_hi - mx);
    best = fminf(best, 0.5f * (a * dx * dx + c * dy * dy) + b * dx * dy);
  }
  return best <= Lm * 1.0002f + 2e-3f;
}

// 4-bit mask of the 8x4 pixel blocks of the 16x8 half tile at (hx0, hy0) (= first pixel CENTRE)
// that the Gaussian's alpha >= 1/255 bounding box touches; 0 when the bounding box misses the half
// tile or the ellipse itself does.
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
    if (mask != 0u) {
      const float Lm = fmaxf(__logf(255.f * g1.y), 0.f) + 2e-3f;
      if (!tc_ellipse_hits_rect(g0.x, g0.y, g0.z, g0.w, g1.x, Lm, hx0, hx0 + 15.f, hy0, hy0 + 7.f))
        mask = 0u;
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
  float P;        // running product of (1 - alpha); the pixel is finished once P <= 1e-4 (an
                  // outside-image pixel starts at 0).  P keeps shrinking after the stop, which is
                  // what zeroes every later weight without a sequential select.
  float T;        // transmittance in front of the next Gaussian, frozen at the stop (A.5: the
                  // stopping Gaussian does not contribute and does not update T)
  int last;       // list index of the last contributor
};
__device__ __forceinline__ void tc_pixel_init(TcPixel &st, float px, float py, bool inside) {
  st.px = px; st.py = py; st.P = inside ? 1.f : 0.f; st.T = 1.f; st.last = 0;
}
__device__ __forceinline__ bool tc_pixel_done(const TcPixel &st) { return st.P <= GAGS_T_STOP; }

// Weights of 8 consecutive Gaussians of the batch (records rec0[8], rec1[8] in shared memory) at
// one pixel.  The only loop-carried dependency is ONE fma per Gaussian (P' = P - alpha P): with
// T_k = P_k while the pixel is alive and P non-increasing, "T_k (1 - alpha_k) <= 1e-4 stops the
// pixel" is the monotone predicate P_{k+1} <= 1e-4, so
//     w_k = (P_{k+1} > 1e-4) ? alpha_k P_k : 0
// reproduces the sequential chain of Appendix A.5 bit for bit while every other operation (alpha
// evaluation, products, selects, bf16 split) is independent across the 8 Gaussians.
// Returns the bf16 hi / lo halves packed for one 16-byte chunk each.
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
  float P[9];
  P[0] = st.P;
#pragma unroll
  for (int k = 0; k < 8; ++k) P[k + 1] = fmaf(-a[k], P[k], P[k]);   // == P when alpha == 0
  float w[8];
  float T = st.T;
  int last = st.last;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool alive = P[k + 1] > GAGS_T_STOP;
    w[k] = alive ? a[k] * P[k] : 0.f;
    T = alive ? P[k + 1] : T;
    last = (w[k] > 0.f) ? gi[k] : last;
  }
  st.P = P[8];
  st.T = T;
  st.last = last;
  split_pack2(w[0], w[1], hi.x, lo.x);
  split_pack2(w[2], w[3], hi.y, lo.y);
  split_pack2(w[4], w[5], hi.z, lo.z);
  split_pack2(w[6], w[7], hi.w, lo.w);
}

// ---- v3 split of the weight evaluation -----------------------------------------------------------
// The per-pixel work above has two very different halves: the alpha evaluation (independent across
// Gaussians and pixels, ~60 % of the instructions, needs the Gaussian record) and the transmittance
// chain (sequential in the Gaussian index, needs nothing but the alphas).  The v3 forward runs them
// in two warp groups, one batch apart:
//   front warps (lane = GAUSSIAN of the batch, record in registers, no shared-memory reads): the 32
//       alphas of the warp's 8x4 pixel block, written as fp32 into the pixel rows of the A stage;
//   chain warps (lane = PIXEL): read their own 128-B row (32 alphas), run the chain and overwrite
//       the row in place with the [hi(32) | lo(32)] bf16 weights the MMA reads.
// Same operations in the same order as tc_weights8, so the two variants are bit-identical.
//
// alpha_k of pixel row r lives at logical byte 4k of the row; rows are stored with the A tile's
// SWIZZLE_128B pattern (16-B chunk index XOR row % 8) so that both the lane-per-Gaussian stores and
// the lane-per-pixel 16-B loads are bank-conflict free.
struct TcFrontRec {
  float mx, my, A, B, C, op;
};
// Position (16-B chunk, before the swizzle) of alpha_k inside its pixel row: the 8 alphas the chain
// consumes in round c (k = 8c .. 8c+7) sit exactly where that round's hi chunk (c) and lo chunk
// (c + 4) are written afterwards, so the in-place rewrite never clobbers an alpha not yet read.
__device__ __forceinline__ uint32_t tc3_alpha_chunk(int k) {
  return (uint32_t)((k >> 3) + (((k >> 2) & 1) << 2));
}
// `rowbase` = shared address of pixel row 0 of this warp's block (rows are 128 B, block = 32 rows:
// row y*8 + x <-> pixel (x, y) of the 8x4 block); (pxc, pyc) = centre of the block's first pixel.
// The y loop is deliberately NOT unrolled: with five roles running side by side the kernel lives
// or dies by its instruction-cache footprint (ncu: 60-70 % no_inst stalls when every role body was
// fully unrolled), so each role's steady-state body is kept around 1-2 KB of SASS.
__device__ __forceinline__ void tc3_front_alphas(const TcFrontRec &g, float pxc, float pyc,
                                                 unsigned char *rowbase, int lane) {
  float dx[8], m[8];
  uint32_t off[8];
  const uint32_t cpos = tc3_alpha_chunk(lane);
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    dx[x] = g.mx - (pxc + (float)x);
    m[x] = g.A * dx[x];
    off[x] = (uint32_t)x * 128u + ((cpos ^ (uint32_t)x) << 4) + ((uint32_t)(lane & 3) << 2);
  }
#pragma unroll 1
  for (int y = 0; y < 4; ++y) {
    const float dy = g.my - (pyc + (float)y);
    const float u = (g.C * dy) * dy;
    unsigned char *rb = rowbase + y * 1024;
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const float t = fmaf(g.B, dy, m[x]);
      const float q = fmaf(t, dx[x], u);
      const float al = fminf(GAGS_ALPHA_MAX, g.op * tc_ex2(q));
      const float a = (q <= 0.f && al >= GAGS_ALPHA_MIN) ? al : 0.f;
      *reinterpret_cast<float *>(rb + off[x]) = a;
    }
  }
}

struct TcChain {
  float P, T;   // EXACT: T = transmittance frozen at the stop.  FAST: T accumulates sum_k w_k.
  int last;
};
// 8 consecutive alphas of the batch (positions k0 .. k0+7) at one pixel -> bf16 hi / lo chunks.
// EXACT (the caller wants last_ids, i.e. the full geometry backward may follow): the same compare /
// select chain as tc_weights8, bit-identical to it; `lastk` receives the in-batch position of the
// last contributor (unchanged when there is none).
// FAST (frozen-geometry training and inference, where nobody reads last_ids): ncu showed the ALU
// pipe (compares, selects, integer ops: half rate) as the kernel's busiest unit, and 6 of the
// chain's 12 instructions per Gaussian sat on it.  Here the "pixel still alive" mask is formed on
// the FMA pipe, s = saturate((P' - 1e-4) * 2^60): the product is exact, so s is exactly 1 or 0, and
// w = (alpha * P) * s is bit-identical to the select it replaces; the final transmittance is
// recovered as 1 - sum_k w_k (one more FADD per Gaussian) instead of a select per step, which
// changes render_alpha by a few 1e-7 (it is mathematically the same quantity).
template <bool FAST>
__device__ __forceinline__ void tc3_chain8(const float4 &a0, const float4 &a1, int k0, TcChain &st,
                                           int &lastk, uint4 &hi, uint4 &lo) {
  const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  float P[9];
  P[0] = st.P;
#pragma unroll
  for (int k = 0; k < 8; ++k) P[k + 1] = fmaf(-a[k], P[k], P[k]);
  float w[8];
  if (FAST) {
    constexpr float BIG = 1152921504606846976.f;          // 2^60
    float sum = st.T;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float s = __saturatef(fmaf(P[k + 1], BIG, -GAGS_T_STOP * BIG));
      w[k] = (a[k] * P[k]) * s;
      sum += w[k];
    }
    st.T = sum;
  } else {
    float T = st.T;
    int lk = lastk;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const bool alive = P[k + 1] > GAGS_T_STOP;
      w[k] = alive ? a[k] * P[k] : 0.f;
      T = alive ? P[k + 1] : T;
      lk = (w[k] > 0.f) ? (k0 + k) : lk;
    }
    st.T = T;
    lastk = lk;
  }
  st.P = P[8];
  split_pack2(w[0], w[1], hi.x, lo.x);
  split_pack2(w[2], w[3], hi.y, lo.y);
  split_pack2(w[4], w[5], hi.z, lo.z);
  split_pack2(w[6], w[7], hi.w, lo.w);
}

// ------------------------------------------------------------------------------------------------
// Scanner: per-thread state of the NT-thread warp group that walks a tile's depth-sorted list NT
// entries at a time, culls every Gaussian whose alpha >= 1/255 bounding box misses the half tile
// (exact: such a Gaussian contributes to no pixel of it) and appends the survivors, in list order,
// to a ring in shared memory.  One round = ids load (prefetched a round ahead) -> geometry gather
// -> cull -> warp ballots + 4-counter prefix -> ring insert.  `issue` starts a round's loads,
// `finish` consumes them, so a caller can put other work between the two.
// Uses named barrier 1 (128 threads).
// ------------------------------------------------------------------------------------------------
template <int NT>
struct TcScannerT {
  const float4 *geom;
  const int *ids;
  float4 *rg0, *rg1;
  int *rgid, *wcnt;
  float hx0, hy0;
  int e, p, pw, lane;
  int scan, qtail, qhead;
  bool pending;
  int pend_idx, pend_gid, nxt_gid;
  float4 pa0, pa1;

  __device__ __forceinline__ void init(const float4 *geom_, const int *ids_, int s, int e_, float hx0_,
                                       float hy0_, float4 *rg0_, float4 *rg1_, int *rgid_,
                                       int *wcnt_, int p_) {
    geom = geom_; ids = ids_; e = e_; hx0 = hx0_; hy0 = hy0_;
    rg0 = rg0_; rg1 = rg1_; rgid = rgid_; wcnt = wcnt_;
    p = p_; pw = p_ >> 5; lane = p_ & 31;
    scan = s; qtail = 0; qhead = 0; pending = false;
    pend_idx = 0; pend_gid = -1;
    pa0 = make_float4(0.f, 0.f, 0.f, 0.f); pa1 = pa0;
    nxt_gid = (scan + p < e) ? __ldg(ids + scan + p) : -1;
  }
  __device__ __forceinline__ bool more() const { return pending || scan < e; }
  __device__ __forceinline__ int queued() const { return qtail - qhead; }
  __device__ __forceinline__ void issue() {
    pend_idx = scan + p;
    pend_gid = nxt_gid;
    if (pend_gid >= 0) {
      pa0 = __ldg(geom + pend_gid * 2);
      pa1 = __ldg(geom + pend_gid * 2 + 1);
    }
    scan += NT;
    nxt_gid = (scan + p < e) ? __ldg(ids + scan + p) : -1;
    pending = true;
  }
  __device__ __forceinline__ void finish() {
    const unsigned mask = (pend_gid >= 0) ? tc_block_mask(pa0, pa1, hx0, hy0) : 0u;
    const bool keep = mask != 0u;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wcnt[pw] = __popc(bal);
    named_bar_sync(1, NT);
    int basec = qtail, total = 0;
#pragma unroll
    for (int k = 0; k < NT / 32; ++k) {
      const int c = wcnt[k];
      if (k < pw) basec += c;
      total += c;
    }
    if (keep) {
      const int slot = (basec + __popc(bal & ((1u << lane) - 1u))) & (TC_RING - 1);
      rg0[slot] = pa0;
      rg1[slot] = make_float4(pa1.x, pa1.y, __int_as_float(pend_idx), __uint_as_float(mask));
      rgid[slot] = pend_gid;
    }
    qtail += total;
    pending = false;
    named_bar_sync(1, NT);
  }
};
using TcScanner = TcScannerT<128>;
