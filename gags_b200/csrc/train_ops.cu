// train_ops.cu — the two elementwise passes either side of the rasteriser in the training step
// (SURVEY.md §8f rows 1 and 2), each fused to ONE pass over its 2 GB operands.
//   gags_l1_loss_fused : loss = sum(|m| * |render - target|) and v_render = scale * |m| * sign(r - t)
//       replaces  l1_loss(feature_map * seg_mask, gt * seg_mask)
//       (/root/reference/utils/loss_utils.py:20-21 called at /root/reference/train.py:162-163) and
//       the 4 elementwise autograd kernels behind it (mul, mul, sub, abs/sign, mean).
//       Algorithmic bytes: 8*HW*D read (+4*HW mask), 4*HW*D written.
//   gags_adam_step     : torch.optim.Adam(lr, betas, eps) single-tensor step
//       (/root/reference/scene/gaussian_model.py:199,208 stepped at /root/reference/train.py:222-223),
//       optionally zeroing the gradient in the same pass (zero_grad(set_to_none) analogue for a
//       persistent gradient buffer).  Algorithmic bytes: 16 read + 12 (+4) written per element.
#include <cstdlib>
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
l1_loss_kernel(const float4 *__restrict__ r, const float4 *__restrict__ t,
               const float *__restrict__ mask, long long n4, int d4, float scale,
               float *__restrict__ loss, float4 *__restrict__ vout) {
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // four independent 16-B loads in flight per thread (two elements of the grid-stride sequence)
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
    const long long i1 = i0 + stride;
    const bool two = i1 < n4;
    const float4 a0 = ldg_nc4(r + i0), b0 = ldg_nc4(t + i0);
    float4 a1 = a0, b1 = b0;
    if (two) { a1 = ldg_nc4(r + i1); b1 = ldg_nc4(t + i1); }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const long long i = u ? i1 : i0;
      const float4 a = u ? a1 : a0, b = u ? b1 : b0;
      const float m = mask ? fabsf(__ldg(mask + i / d4)) : 1.f;
      const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
      acc += m * (fabsf(dx) + fabsf(dy) + fabsf(dz) + fabsf(dw));
      const float s = scale * m;
      float4 g;
      g.x = dx > 0.f ? s : (dx < 0.f ? -s : 0.f);
      g.y = dy > 0.f ? s : (dy < 0.f ? -s : 0.f);
      g.z = dz > 0.f ? s : (dz < 0.f ? -s : 0.f);
      g.w = dw > 0.f ? s : (dw < 0.f ? -s : 0.f);
      vout[i] = g;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += s_part[w];
    atomicAdd(loss, v);
  }
}

// Same loss with the target given in the reference's compact form: a per-pixel segment id and a
// per-segment embedding table (what read_sam_clip_feature gathers into a dense [H,W,D] map every
// iteration, /root/reference/scene/dataset_readers.py:54-121 called at train.py:162).  The table
// (n_seg x D fp32, a few hundred KB) stays in L1/L2, so the pass reads 4 B per pixel instead of 4D.
// seg < 0 = pixel without a target (weight 0).
__global__ void __launch_bounds__(256)
l1_loss_segmap_kernel(const float4 *__restrict__ r, const int *__restrict__ seg,
                      const float4 *__restrict__ emb, const float *__restrict__ mask, long long n4,
                      int d4, int n_seg, float scale, float *__restrict__ loss,
                      float4 *__restrict__ vout) {
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
    const long long i1 = i0 + stride;
    const bool two = i1 < n4;
    const float4 a0 = ldg_nc4(r + i0);
    float4 a1 = a0;
    if (two) a1 = ldg_nc4(r + i1);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const long long i = u ? i1 : i0;
      const float4 a = u ? a1 : a0;
      const long long pix = i / d4;
      const int c4 = (int)(i - pix * d4);
      const int sg = __ldg(seg + pix);
      const bool ok = sg >= 0 && sg < n_seg;
      const float4 b = ok ? __ldg(emb + (size_t)sg * d4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float m = ok ? (mask ? fabsf(__ldg(mask + pix)) : 1.f) : 0.f;
      const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
      acc += m * (fabsf(dx) + fabsf(dy) + fabsf(dz) + fabsf(dw));
      const float s = scale * m;
      float4 g;
      g.x = dx > 0.f ? s : (dx < 0.f ? -s : 0.f);
      g.y = dy > 0.f ? s : (dy < 0.f ? -s : 0.f);
      g.z = dz > 0.f ? s : (dz < 0.f ? -s : 0.f);
      g.w = dw > 0.f ? s : (dw < 0.f ? -s : 0.f);
      vout[i] = g;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += s_part[w];
    atomicAdd(loss, v);
  }
}

// The reference's real target (read_sam_clip_feature, /root/reference/scene/dataset_readers.py:54-121,
// default mode, called at /root/reference/train.py:162-166): three SAM levels (s, m, l) of segment ids
// per pixel, one embedding table, and a per-pixel weight per level (the scale decoder's softmax):
//     gt[px, :] = sum_l scale[l, px] * emb[seg[l, px], :]        valid(px) = all three seg != -1
//     loss = mean(|render * valid - gt * valid|)                  (utils/loss_utils.py:20-21)
// (the reference's bilinear resize of the three maps is the identity when the scale map has the
// image's size, which is how train.py runs it; other sizes take the dense fallback in Python).
// One pass: loss, v_render = scale * valid * sign(render - gt) and, when requested, the gradient
// w.r.t. the scale map, v_scale[l, px] = -sum_ch v_render[px, ch] * emb[seg[l, px], ch], which is what
// trains the scale decoder through the target (train.py:149 -> :162).
template <bool VS>
__global__ void __launch_bounds__(256)
l1_loss_sam_kernel(const float4 *__restrict__ r, const int *__restrict__ seg3,
                   const float4 *__restrict__ emb, const float *__restrict__ scale3, long long hw,
                   long long n4, int d4, int n_seg, float scale, float *__restrict__ loss,
                   float4 *__restrict__ vout, float *__restrict__ v_scale) {
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // with VS the host guarantees d4 % 32 == 0: a warp's 32 float4 belong to ONE pixel
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = ldg_nc4(r + i);
    const long long pix = i / d4;
    const int c4 = (int)(i - pix * d4);
    int sg[3];
    float w[3];
    bool ok = true;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      sg[l] = __ldg(seg3 + l * hw + pix);
      w[l] = __ldg(scale3 + l * hw + pix);
      ok = ok && sg[l] >= 0 && sg[l] < n_seg;
    }
    float4 e[3];
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      e[l] = ok ? __ldg(emb + (size_t)sg[l] * d4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      t.x = fmaf(w[l], e[l].x, t.x); t.y = fmaf(w[l], e[l].y, t.y);
      t.z = fmaf(w[l], e[l].z, t.z); t.w = fmaf(w[l], e[l].w, t.w);
    }
    const float m = ok ? 1.f : 0.f;
    const float dx = a.x - t.x, dy = a.y - t.y, dz = a.z - t.z, dw = a.w - t.w;
    acc += m * (fabsf(dx) + fabsf(dy) + fabsf(dz) + fabsf(dw));
    const float s = scale * m;
    float4 g;
    g.x = dx > 0.f ? s : (dx < 0.f ? -s : 0.f);
    g.y = dy > 0.f ? s : (dy < 0.f ? -s : 0.f);
    g.z = dz > 0.f ? s : (dz < 0.f ? -s : 0.f);
    g.w = dw > 0.f ? s : (dw < 0.f ? -s : 0.f);
    vout[i] = g;
    if (VS) {
      float vs[3];
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        vs[l] = -(g.x * e[l].x + g.y * e[l].y + g.z * e[l].z + g.w * e[l].w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vs[l] += __shfl_xor_sync(0xffffffffu, vs[l], o);
      }
      if ((threadIdx.x & 31) == 0 && ok) {
#pragma unroll
        for (int l = 0; l < 3; ++l) atomicAdd(v_scale + l * hw + pix, vs[l]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += s_part[w];
    atomicAdd(loss, v);
  }
}

// ONE Adam update, spelled with explicit roundings: every kernel below (dense, row-sparse, lazy,
// peer, multicast, tail) must produce the same bits for the same operands — the row-sparse and lazy
// forms are specified as "bit-identical to the dense pass" — so nothing is left to the compiler's
// choice of which product of `b1*m + (1-b1)*g` to contract into an FMA.
__device__ __forceinline__ void adam_update1(float &p, float g, float &m, float &v, float step_size,
                                             float b1, float b2, float omb1, float omb2,
                                             float inv_sqrt_bc2, float eps) {
  m = fmaf(omb1, g, __fmul_rn(b1, m));
  v = fmaf(omb2, __fmul_rn(g, g), __fmul_rn(b2, v));
  const float denom = fmaf(sqrtf(v), inv_sqrt_bc2, eps);
  p = fmaf(-step_size, __fdiv_rn(m, denom), p);
}
#define GAGS_ADAM1(P, G, M, V, c) \
  adam_update1(P.c, G.c, M.c, V.c, step_size, b1, b2, omb1, omb2, inv_sqrt_bc2, eps);

__global__ void __launch_bounds__(256)
adam_kernel(float4 *__restrict__ p, float4 *__restrict__ g, float4 *__restrict__ m,
            float4 *__restrict__ v, long long n4, float step_size, float b1, float b2,
            float omb1, float omb2, float inv_sqrt_bc2, float eps, int zero_grad) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  // two independent 16-B streams per array in flight per thread (128 B per thread)
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
    const long long i1 = i0 + stride;
    const bool two = i1 < n4;
    const float4 ga = g[i0];
    float4 ma = m[i0], va = v[i0], pa = p[i0];
    float4 gb = ga, mb = ma, vb = va, pb = pa;
    if (two) { gb = g[i1]; mb = m[i1]; vb = v[i1]; pb = p[i1]; }
    GAGS_ADAM1(pa, ga, ma, va, x) GAGS_ADAM1(pa, ga, ma, va, y)
    GAGS_ADAM1(pa, ga, ma, va, z) GAGS_ADAM1(pa, ga, ma, va, w)
    m[i0] = ma; v[i0] = va; p[i0] = pa;
    if (zero_grad) g[i0] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (two) {
      GAGS_ADAM1(pb, gb, mb, vb, x) GAGS_ADAM1(pb, gb, mb, vb, y)
      GAGS_ADAM1(pb, gb, mb, vb, z) GAGS_ADAM1(pb, gb, mb, vb, w)
      m[i1] = mb; v[i1] = vb; p[i1] = pb;
      if (zero_grad) g[i1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// Row-sparse form of adam_kernel for a [rows, D] table whose gradient is mostly zero rows (one view
// touches ~1/6 of the Gaussians at config 3): `flags[row] == 0` PROMISES that gradient row `row` is
// all zero, so the row's gradient is neither read nor re-zeroed and the update runs with g = 0 —
// the same arithmetic on the same values as the dense pass, bit for bit (m and v still decay, p still
// moves by its momentum).  Touched rows are read, applied and zeroed in this pass, so the gradient
// buffer is all-zero again afterwards without a separate 4 B/parameter fill.  24 B per parameter on
// untouched rows, 32 B on touched ones (dense pass + fill: 28 + 4).  The flags are cleared by the
// caller behind the kernel (a row's threads may sit in different CTAs).
template <bool POW2>
__global__ void __launch_bounds__(256)
adam_rows_kernel(float4 *__restrict__ p, float4 *__restrict__ g, float4 *__restrict__ m,
                 float4 *__restrict__ v, const unsigned char *__restrict__ flags, long long n4,
                 int row4, int shift, float step_size, float b1, float b2, float omb1, float omb2,
                 float inv_sqrt_bc2, float eps) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
    const long long i1 = i0 + stride;
    const bool two = i1 < n4;
    const bool fa = flags[POW2 ? (i0 >> shift) : (i0 / row4)] != 0;
    const bool fb = two && flags[POW2 ? (i1 >> shift) : (i1 / row4)] != 0;
    float4 ma = m[i0], va = v[i0], pa = p[i0];
    float4 mb = ma, vb = va, pb = pa;
    if (two) { mb = m[i1]; vb = v[i1]; pb = p[i1]; }
    float4 ga = zero, gb = zero;
    if (fa) ga = g[i0];
    if (fb) gb = g[i1];
    GAGS_ADAM1(pa, ga, ma, va, x) GAGS_ADAM1(pa, ga, ma, va, y)
    GAGS_ADAM1(pa, ga, ma, va, z) GAGS_ADAM1(pa, ga, ma, va, w)
    m[i0] = ma; v[i0] = va; p[i0] = pa;
    if (fa) g[i0] = zero;
    if (two) {
      GAGS_ADAM1(pb, gb, mb, vb, x) GAGS_ADAM1(pb, gb, mb, vb, y)
      GAGS_ADAM1(pb, gb, mb, vb, z) GAGS_ADAM1(pb, gb, mb, vb, w)
      m[i1] = mb; v[i1] = vb; p[i1] = pb;
      if (fb) g[i1] = zero;
    }
  }
}

// ---- lazily evaluated Adam for a row-sparse gradient --------------------------------------------
// With g = 0 a row's Adam step depends on nothing but the row's own (p, m, v) and the step's two
// scalars, so it does not have to be TAKEN at that step: a row that no view touches for k steps can
// take those k zero-gradient steps later, in registers, in one visit — the same fp32 operations in
// the same order, hence bit-identical to the dense pass, for 1/k of its memory traffic.  `last[r]`
// is the optimiser step row r is current to; `consts[s]` = (lr / (1 - b1^s), 1 / sqrt(1 - b2^s)) as
// gags_adam_step computes them for step s (written by the host, one entry per step).  One call:
// every selected row (flags[r] != 0, or all rows when flags == NULL) first takes the zero-gradient
// steps last[r]+1 .. t_to; then, if t_apply != 0 (= t_to + 1), the step t_apply with its gradient
// row, which is re-zeroed.  The two-pass forward (gags_blend_fwd_weights -> mark rows -> here with
// t_apply = 0 -> gags_blend_fwd_from_cache) brings exactly the rows a view reads up to date before
// they are read; the optimiser step then visits only the rows the view (or, multi-GPU, any rank's
// view) touched; a flush (flags == NULL) materialises the whole table.
// A zero-gradient step runs the very same adam_update1 with the gradient operand `gzero` (a kernel
// ARGUMENT holding 0.f, so that nothing is folded away): e.g. fma(1-b1, +0, -0) is +0, exactly what
// the dense pass leaves in m.
struct AdamC { float b1, b2, omb1, omb2, eps; };

__global__ void __launch_bounds__(256)
adam_lazy_rows_kernel(float4 *__restrict__ p, float4 *__restrict__ g, float4 *__restrict__ m,
                      float4 *__restrict__ v, const unsigned char *__restrict__ flags,
                      int *__restrict__ last, const float2 *__restrict__ consts, long long rows,
                      int row4, int t_to, int t_apply, float gzero, AdamC c) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const float b1 = c.b1, b2 = c.b2, omb1 = c.omb1, omb2 = c.omb2, eps = c.eps;
#define GAGS_ADAM4(P, G, M, V)                                                        \
    GAGS_ADAM1(P, G, M, V, x) GAGS_ADAM1(P, G, M, V, y) GAGS_ADAM1(P, G, M, V, z) GAGS_ADAM1(P, G, M, V, w)
  for (long long base = warp * 32; base < rows; base += nwarps * 32) {
    const long long r = base + lane;
    bool f = r < rows && (flags == nullptr || flags[r] != 0);
    const int s0 = f ? last[r] : 0;
    if (f && t_apply == 0 && s0 >= t_to) f = false;          // already current
    unsigned mask = __ballot_sync(0xffffffffu, f);
    while (mask) {
      const int k = __ffs(mask) - 1;
      mask &= mask - 1;
      const long long row = base + k;
      const int sk = __shfl_sync(0xffffffffu, s0, k);
      const long long rb = row * (long long)row4;
      for (int c0 = 0; c0 < row4; c0 += 64) {
        const int ca = c0 + lane, cb = c0 + 32 + lane;
        const bool ha = ca < row4, hb = cb < row4;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 pa = zero, ma = zero, va = zero, pb = zero, mb = zero, vb = zero, ga = zero, gb = zero;
        if (ha) { pa = p[rb + ca]; ma = m[rb + ca]; va = v[rb + ca]; }
        if (hb) { pb = p[rb + cb]; mb = m[rb + cb]; vb = v[rb + cb]; }
        if (t_apply != 0) {
          if (ha) ga = g[rb + ca];
          if (hb) gb = g[rb + cb];
        }
        // a row whose moments are exactly zero does not move under zero-gradient steps
        const bool still = ma.x == 0.f && ma.y == 0.f && ma.z == 0.f && ma.w == 0.f && va.x == 0.f &&
                           va.y == 0.f && va.z == 0.f && va.w == 0.f && mb.x == 0.f && mb.y == 0.f &&
                           mb.z == 0.f && mb.w == 0.f && vb.x == 0.f && vb.y == 0.f && vb.z == 0.f &&
                           vb.w == 0.f;
        if (!__all_sync(0xffffffffu, still)) {
          const float4 gz = make_float4(gzero, gzero, gzero, gzero);
#pragma unroll 1
          for (int s = sk + 1; s <= t_to; ++s) {
            const float2 cs = __ldg(consts + s);
            const float step_size = cs.x, inv_sqrt_bc2 = cs.y;
            GAGS_ADAM4(pa, gz, ma, va)
            GAGS_ADAM4(pb, gz, mb, vb)
          }
        }
        if (t_apply != 0) {
          const float2 cs = __ldg(consts + t_apply);
          const float step_size = cs.x, inv_sqrt_bc2 = cs.y;
          GAGS_ADAM4(pa, ga, ma, va)
          GAGS_ADAM4(pb, gb, mb, vb)
          if (ha) g[rb + ca] = zero;
          if (hb) g[rb + cb] = zero;
        }
        if (ha) { p[rb + ca] = pa; m[rb + ca] = ma; v[rb + ca] = va; }
        if (hb) { p[rb + cb] = pb; m[rb + cb] = mb; v[rb + cb] = vb; }
      }
      if (lane == 0) last[row] = t_apply != 0 ? t_apply : t_to;
    }
  }
#undef GAGS_ADAM4
}

// ---- gradient all-reduce + Adam + parameter all-gather as ONE kernel over NVLink peer memory ---
// View-parallel training replicates the feature table and sums its gradient over the ranks every
// step (2 GB at config 3).  Here rank r owns the float4 range [start, start + count) of the table
// (and only that slice of the Adam moments): it LOADS that range of every rank's gradient buffer
// straight from peer memory, sums in rank order (every rank gets bit-identical parameters), applies
// Adam, and STORES the new parameters into every rank's replica.  Inbound gradients and outbound
// parameters use the two directions of the NVLink ports at the same time, the moments' HBM traffic
// is divided by the world size, and no staging copy of the gradient exists.  The caller brackets
// the launch with two inter-rank barriers (gradients final before; replicas complete after).
constexpr int GAGS_MAX_PEERS = 16;
struct PeerPtrs {
  const float4 *grad[GAGS_MAX_PEERS];
  float4 *param[GAGS_MAX_PEERS];
};

template <int MAXW>
__global__ void __launch_bounds__(256)
adam_peer_kernel(PeerPtrs pp, int world, int rank, float4 *__restrict__ m, float4 *__restrict__ v,
                 long long start, long long count, float step_size, float b1, float b2, float omb1,
                 float omb2, float inv_sqrt_bc2, float eps) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const long long i = start + k;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 gq[MAXW];
#pragma unroll
    for (int q = 0; q < MAXW; ++q)
      if (q < world) gq[q] = pp.grad[q][i];             // all loads in flight before the first add
    float4 pa = pp.param[rank][i];
    float4 ma = m[k], va = v[k];
#pragma unroll
    for (int q = 0; q < MAXW; ++q)
      if (q < world) { g.x += gq[q].x; g.y += gq[q].y; g.z += gq[q].z; g.w += gq[q].w; }
    GAGS_ADAM1(pa, g, ma, va, x) GAGS_ADAM1(pa, g, ma, va, y)
    GAGS_ADAM1(pa, g, ma, va, z) GAGS_ADAM1(pa, g, ma, va, w)
    m[k] = ma;
    v[k] = va;
#pragma unroll
    for (int q = 0; q < MAXW; ++q)
      if (q < world) pp.param[q][i] = pa;
  }
}

// The same step through the NVSwitch (NVLS): `mc_grad` / `mc_param` are MULTICAST addresses of the
// symmetric buffers.  multimem.ld_reduce returns the sum over all ranks' copies, formed inside the
// switch, and multimem.st writes all replicas with one store, so a rank moves 1/G of the table in
// each direction instead of (G-1)/G and the kernel is bound by HBM, not by the NVLink ports.
template <int U>
__global__ void __launch_bounds__(256)
adam_multicast_kernel(const float4 *mc_grad, float4 *mc_param, const float4 *__restrict__ p_local,
                      float4 *__restrict__ m, float4 *__restrict__ v, long long start,
                      long long count, float step_size, float b1, float b2, float omb1, float omb2,
                      float inv_sqrt_bc2, float eps) {
  // a small grid (the exchange runs beside the next view's projection / sort) with U switch
  // reductions in flight per thread to cover the NVLink round trip
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; k0 < count; k0 += U * stride) {
    float4 g[U], pa[U], ma[U], va[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long k = k0 + u * stride;
      if (k < count)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(g[u].x), "=f"(g[u].y), "=f"(g[u].z), "=f"(g[u].w)
                     : "l"(mc_grad + start + k)
                     : "memory");
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long k = k0 + u * stride;
      if (k < count) { pa[u] = p_local[start + k]; ma[u] = m[k]; va[u] = v[k]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long k = k0 + u * stride;
      if (k < count) {
        GAGS_ADAM1(pa[u], g[u], ma[u], va[u], x) GAGS_ADAM1(pa[u], g[u], ma[u], va[u], y)
        GAGS_ADAM1(pa[u], g[u], ma[u], va[u], z) GAGS_ADAM1(pa[u], g[u], ma[u], va[u], w)
        m[k] = ma[u];
        v[k] = va[u];
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(
                         mc_param + start + k),
                     "f"(pa[u].x), "f"(pa[u].y), "f"(pa[u].z), "f"(pa[u].w)
                     : "memory");
      }
    }
  }
}

// ---- row-sparse gradient all-reduce over NVLink peer memory -------------------------------------
// A view's feature backward touches a small part of the table's rows (7 % at config 3; the union
// over 8 consecutive views is 15 %), and every rank knows which ones (the row flags the backward
// sets, rasterization.row_flags).  Instead of exchanging the whole [N, D] gradient (2 GB) and the
// whole updated table (2 GB) per step, every rank keeps the full optimiser state and only the rows
// flagged on ANY rank are summed: rank r owns a contiguous range of 4-row flag words; for each
// owned word it ORs the ranks' flags, and for every flagged row it forms the sum of that row over
// all ranks and writes it back into ALL replicas' gradient buffers (in place).  Each rank also
// builds its own copy of the union flags for all rows (N bytes read through the fabric), which is
// what its local row-sparse Adam pass (adam_rows_kernel) consumes afterwards.  MC = NVLS:
// multimem.ld_reduce forms the sums inside the NVSwitch and multimem.st writes every replica;
// otherwise plain peer loads (summed in rank order) and stores.  Every replica receives the same
// bits for a row because exactly one rank computes it.  The caller brackets the launch with two
// inter-rank barriers (all backward passes done before; all sums landed after).
struct RowPeerPtrs {
  float4 *grad[GAGS_MAX_PEERS];
  const unsigned *flags[GAGS_MAX_PEERS];
};

__device__ __forceinline__ unsigned flags_or_mc(const unsigned *mc) {
  unsigned u;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.or.b32 %0, [%1];" : "=r"(u) : "l"(mc) : "memory");
  return u;
}
__device__ __forceinline__ float4 ld_sum_mc(const float4 *mc) {
  float4 g;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w)
               : "l"(mc)
               : "memory");
  return g;
}
__device__ __forceinline__ void st_mc(float4 *mc, const float4 &v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

template <bool MC, int MAXW>
__global__ void __launch_bounds__(256)
grad_rows_allreduce_kernel(RowPeerPtrs pp, int world, float4 *mc_grad, const unsigned *mc_flags,
                           unsigned *__restrict__ uflags, long long words, long long w0,
                           long long w1, int row4) {
  const int lane = threadIdx.x & 31;
  const long long gthreads = (long long)gridDim.x * blockDim.x;
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  auto union_word = [&](long long w) -> unsigned {
    if (MC) return flags_or_mc(mc_flags + w);
    unsigned u = 0;
#pragma unroll
    for (int q = 0; q < MAXW; ++q)
      if (q < world) u |= pp.flags[q][w];
    return u;
  };
  // (1) this rank's copy of the union flags of ALL rows
  for (long long w = gtid; w < words; w += gthreads) uflags[w] = union_word(w);
  // (2) the owned words: a warp takes 32 consecutive words with one coalesced flag load (the flag
  // round trip through the fabric is paid once per 128 rows, not once per word), then visits the
  // flagged words; the flagged rows of a word are summed together, lanes across the row, up to
  // eight 16-byte reductions in flight per lane
  const long long gwarp = gtid >> 5, nwarps = gthreads >> 5;
  for (long long wb = w0 + gwarp * 32; wb < w1; wb += nwarps * 32) {
    const long long wl = wb + lane;
    const unsigned u = wl < w1 ? union_word(wl) : 0u;
    unsigned wm = __ballot_sync(0xffffffffu, u != 0u);
    while (wm) {
      const int k = __ffs(wm) - 1;
      wm &= wm - 1;
      const unsigned uk = __shfl_sync(0xffffffffu, u, k);
      const long long row0 = (wb + k) * 4;
      for (int c0 = 0; c0 < row4; c0 += 64) {
        float4 acc[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (((uk >> (8 * r)) & 0xffu) == 0u) continue;            // warp-uniform
          const long long base = (row0 + r) * (long long)row4;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = c0 + h * 32 + lane;
            if (c < row4) {
              if (MC) {
                acc[r][h] = ld_sum_mc(mc_grad + base + c);
              } else {
                float4 gq[MAXW];
#pragma unroll
                for (int q = 0; q < MAXW; ++q)
                  if (q < world) gq[q] = pp.grad[q][base + c];
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < MAXW; ++q)
                  if (q < world) { a.x += gq[q].x; a.y += gq[q].y; a.z += gq[q].z; a.w += gq[q].w; }
                acc[r][h] = a;
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (((uk >> (8 * r)) & 0xffu) == 0u) continue;
          const long long base = (row0 + r) * (long long)row4;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = c0 + h * 32 + lane;
            if (c < row4) {
              if (MC) {
                st_mc(mc_grad + base + c, acc[r][h]);
              } else {
#pragma unroll
                for (int q = 0; q < MAXW; ++q)
                  if (q < world) pp.grad[q][base + c] = acc[r][h];
              }
            }
          }
        }
      }
    }
  }
}

__global__ void adam_tail_kernel(float *p, float *g, float *m, float *v, long long start,
                                 long long n, float step_size, float b1, float b2, float omb1,
                                 float omb2, float inv_sqrt_bc2, float eps, int zero_grad) {
  const long long i = start + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  float mi = m[i], vi = v[i], pi = p[i];
  adam_update1(pi, gi, mi, vi, step_size, b1, b2, omb1, omb2, inv_sqrt_bc2, eps);
  p[i] = pi; m[i] = mi; v[i] = vi;
  if (zero_grad) g[i] = 0.f;
}

// v *= *scale, skipped entirely (one 4-byte read per thread block) when *scale == 1: the usual
// loss.backward() seeds the graph with ones, so the fused loss's gradient is already final.
__global__ void __launch_bounds__(256)
scale_dev_kernel(float4 *__restrict__ v, const float *__restrict__ scale, long long n4) {
  const float s = __ldg(scale);
  if (s == 1.f) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 x = v[i];
    x.x *= s; x.y *= s; x.z *= s; x.w *= s;
    v[i] = x;
  }
}

}  // namespace

extern "C" int gags_scale_inplace(float *v, const float *scale_dev, int64_t numel, void *stream) {
  if (!v || !scale_dev || numel < 0 || (numel & 3)) return GAGS_EINVAL;
  if (!gags_aligned16(v)) return GAGS_EALIGN;
  if (numel == 0) return 0;
  const long long n4 = numel / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > gags_sm_count() * 4) blocks = gags_sm_count() * 4;
  scale_dev_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float4 *>(v), scale_dev, n4);
  GAGS_CHECK_LAUNCH();
  return 0;
}

extern "C" int gags_l1_loss_fused(const float *render, const float *target, const float *mask,
                                  int64_t HW, int32_t D, float grad_scale, float *loss_out,
                                  float *v_render, void *stream) {
  if (!render || !target || !loss_out || !v_render || HW < 0 || D < 1) return GAGS_EINVAL;
  if (D % 4 != 0) return GAGS_EINVAL;
  if (!gags_aligned16(render) || !gags_aligned16(target) || !gags_aligned16(v_render)) return GAGS_EALIGN;
  if (HW == 0) return 0;
  const long long n4 = (long long)HW * (D / 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > gags_sm_count() * 8) blocks = gags_sm_count() * 8;     // 8 resident 256-thread CTAs per SM, grid-stride
  l1_loss_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4 *>(render), reinterpret_cast<const float4 *>(target), mask, n4,
      D / 4, grad_scale, loss_out, reinterpret_cast<float4 *>(v_render));
  GAGS_CHECK_LAUNCH();
  return 0;
}

extern "C" int gags_l1_loss_segmap(const float *render, const int32_t *seg, const float *emb,
                                   const float *mask, int64_t HW, int32_t D, int32_t n_seg,
                                   float grad_scale, float *loss_out, float *v_render,
                                   void *stream) {
  if (!render || !seg || !emb || !loss_out || !v_render || HW < 0 || D < 1 || n_seg < 1)
    return GAGS_EINVAL;
  if (D % 4 != 0) return GAGS_EINVAL;
  if (!gags_aligned16(render) || !gags_aligned16(emb) || !gags_aligned16(v_render)) return GAGS_EALIGN;
  if (HW == 0) return 0;
  const long long n4 = (long long)HW * (D / 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > gags_sm_count() * 8) blocks = gags_sm_count() * 8;
  l1_loss_segmap_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4 *>(render), seg, reinterpret_cast<const float4 *>(emb), mask, n4,
      D / 4, n_seg, grad_scale, loss_out, reinterpret_cast<float4 *>(v_render));
  GAGS_CHECK_LAUNCH();
  return 0;
}

// v_scale_map (optional, [3][HW], zeroed by the caller) needs D % 128 == 0 (a warp per pixel slice).
extern "C" int gags_l1_loss_sam(const float *render, const int32_t *seg3, const float *emb,
                                const float *scale_map3, int64_t HW, int32_t D, int32_t n_seg,
                                float grad_scale, float *loss_out, float *v_render,
                                float *v_scale_map, void *stream) {
  if (!render || !seg3 || !emb || !scale_map3 || !loss_out || !v_render || HW < 0 || D < 1 || n_seg < 1)
    return GAGS_EINVAL;
  if (D % 4 != 0) return GAGS_EINVAL;
  if (v_scale_map && D % 128 != 0) return GAGS_ERANGE;
  if (!gags_aligned16(render) || !gags_aligned16(emb) || !gags_aligned16(v_render)) return GAGS_EALIGN;
  if (HW == 0) return 0;
  const long long n4 = (long long)HW * (D / 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > gags_sm_count() * 8) blocks = gags_sm_count() * 8;
  if (v_scale_map)
    l1_loss_sam_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(render), seg3, reinterpret_cast<const float4 *>(emb),
        scale_map3, HW, n4, D / 4, n_seg, grad_scale, loss_out,
        reinterpret_cast<float4 *>(v_render), v_scale_map);
  else
    l1_loss_sam_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(render), seg3, reinterpret_cast<const float4 *>(emb),
        scale_map3, HW, n4, D / 4, n_seg, grad_scale, loss_out,
        reinterpret_cast<float4 *>(v_render), nullptr);
  GAGS_CHECK_LAUNCH();
  return 0;
}

extern "C" int gags_adam_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq,
                              int64_t numel, double lr, double beta1, double beta2, double eps,
                              int32_t step, int32_t zero_grad, void *stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || numel < 0 || step < 1) return GAGS_EINVAL;
  if (!gags_aligned16(param) || !gags_aligned16(grad) || !gags_aligned16(exp_avg) ||
      !gags_aligned16(exp_avg_sq))
    return GAGS_EALIGN;
  if (numel == 0) return 0;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  // 1 - beta in double, as torch does (1.f - 0.999f is off by 5e-5 relative)
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
  const long long n4 = numel / 4;
  cudaStream_t st = (cudaStream_t)stream;
  if (n4 > 0) {
    long long blocks = (n4 + 255) / 256;
    // half of each SM's thread slots stay free next to this pure HBM stream
    if (blocks > gags_sm_count() * 4) blocks = gags_sm_count() * 4;
    adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(
        reinterpret_cast<float4 *>(param), reinterpret_cast<float4 *>(grad),
        reinterpret_cast<float4 *>(exp_avg), reinterpret_cast<float4 *>(exp_avg_sq), n4, step_size,
        (float)beta1, (float)beta2, omb1, omb2, inv_sqrt_bc2, (float)eps, zero_grad);
    GAGS_CHECK_LAUNCH();
  }
  if (n4 * 4 < numel) {
    adam_tail_kernel<<<1, 32, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n4 * 4, numel, step_size,
                                       (float)beta1, (float)beta2, omb1, omb2, inv_sqrt_bc2,
                                       (float)eps, zero_grad);
    GAGS_CHECK_LAUNCH();
  }
  return 0;
}

// gags_adam_step on a [rows, D] table with per-row gradient flags (see adam_rows_kernel): rows whose
// flag is 0 take the g = 0 update without their gradient being read; flagged rows are applied and
// their gradient zeroed; the flags are cleared behind the kernel.  D % 4 == 0.
extern "C" int gags_adam_step_rows(float *param, float *grad, float *exp_avg, float *exp_avg_sq,
                                   uint8_t *row_flags, int64_t rows, int32_t D, double lr,
                                   double beta1, double beta2, double eps, int32_t step,
                                   void *stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !row_flags || rows < 0 || D < 4 || step < 1)
    return GAGS_EINVAL;
  if (D % 4 != 0) return GAGS_EINVAL;
  if (!gags_aligned16(param) || !gags_aligned16(grad) || !gags_aligned16(exp_avg) ||
      !gags_aligned16(exp_avg_sq))
    return GAGS_EALIGN;
  if (rows == 0) return 0;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
  const int row4 = D / 4;
  const long long n4 = (long long)rows * row4;
  cudaStream_t st = (cudaStream_t)stream;
  long long blocks = (n4 + 255) / 256;
  if (blocks > gags_sm_count() * 4) blocks = gags_sm_count() * 4;
  int shift = -1;
  if ((row4 & (row4 - 1)) == 0) { shift = 0; while ((1 << shift) < row4) ++shift; }
#define GAGS_ROWS_ARGS reinterpret_cast<float4 *>(param), reinterpret_cast<float4 *>(grad),        \
      reinterpret_cast<float4 *>(exp_avg), reinterpret_cast<float4 *>(exp_avg_sq), row_flags, n4,  \
      row4, shift, step_size, (float)beta1, (float)beta2, omb1, omb2, inv_sqrt_bc2, (float)eps
  if (shift >= 0) adam_rows_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(GAGS_ROWS_ARGS);
  else adam_rows_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(GAGS_ROWS_ARGS);
#undef GAGS_ROWS_ARGS
  GAGS_CHECK_LAUNCH();
  cudaError_t e = cudaMemsetAsync(row_flags, 0, (size_t)rows, st);
  return e == cudaSuccess ? 0 : (int)e;
}

// (lr / (1 - beta1^step), 1 / sqrt(1 - beta2^step)) exactly as gags_adam_step forms them: the host
// writes one pair per optimiser step into the table gags_adam_lazy_rows reads.
extern "C" int gags_adam_step_consts(double lr, double beta1, double beta2, int32_t step,
                                     float *out2_host) {
  if (!out2_host || step < 1) return GAGS_EINVAL;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  out2_host[0] = (float)(lr / bc1);
  out2_host[1] = (float)(1.0 / sqrt(bc2));
  return 0;
}

// Lazily evaluated Adam on the rows selected by row_flags (NULL = every row), see
// adam_lazy_rows_kernel.  t_apply is 0 (catch up only; grad may be NULL) or t_to + 1.  With
// clear_flags the flags are zeroed behind the kernel.
extern "C" int gags_adam_lazy_rows(float *param, float *grad, float *exp_avg, float *exp_avg_sq,
                                   uint8_t *row_flags, int32_t *last_step, const float *step_consts,
                                   int64_t rows, int32_t D, int32_t t_to, int32_t t_apply,
                                   double beta1, double beta2, double eps, int32_t clear_flags,
                                   void *stream) {
  if (!param || !exp_avg || !exp_avg_sq || !last_step || !step_consts || rows < 0 || D < 4 ||
      t_to < 0)
    return GAGS_EINVAL;
  if (D % 4 != 0) return GAGS_EINVAL;
  if (t_apply != 0 && (t_apply != t_to + 1 || !grad)) return GAGS_EINVAL;
  if (clear_flags && !row_flags) return GAGS_EINVAL;
  if (!gags_aligned16(param) || !gags_aligned16(exp_avg) || !gags_aligned16(exp_avg_sq) ||
      (grad && !gags_aligned16(grad)) || ((uintptr_t)step_consts & 7))
    return GAGS_EALIGN;
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  AdamC c;
  c.b1 = (float)beta1; c.b2 = (float)beta2;
  c.omb1 = (float)(1.0 - beta1); c.omb2 = (float)(1.0 - beta2);
  c.eps = (float)eps;
  long long blocks = (rows + 255) / 256;                       // one row per lane-slot of flags
  static int lazy_ctas = -1;                                   // tuning: GAGS_B200_LAZY_GRID=CTAs/SM
  if (lazy_ctas < 0) {
    const char *e = getenv("GAGS_B200_LAZY_GRID");
    lazy_ctas = (e && atoi(e) > 0 && atoi(e) <= 32) ? atoi(e) : 8;
  }
  const long long cap = (long long)gags_sm_count() * lazy_ctas;
  if (blocks > cap) blocks = cap;
  adam_lazy_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(
      reinterpret_cast<float4 *>(param), reinterpret_cast<float4 *>(grad),
      reinterpret_cast<float4 *>(exp_avg), reinterpret_cast<float4 *>(exp_avg_sq), row_flags,
      last_step, reinterpret_cast<const float2 *>(step_consts), rows, D / 4, t_to, t_apply, 0.f, c);
  GAGS_CHECK_LAUNCH();
  if (clear_flags) {
    cudaError_t e = cudaMemsetAsync(row_flags, 0, (size_t)rows, st);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

// Tuning hook for the two exchange kernels (tools/peer_rate.py sweeps it): CTAs per SM of their
// grids; 0 = the built-in choice (4 for the unicast form, 2 for the multicast form).
static int g_peer_ctas_per_sm = 0;
static int g_peer_unroll = 4;
extern "C" int gags_set_peer_grid(int32_t ctas_per_sm) {
  if (ctas_per_sm < 0 || ctas_per_sm > 16) return GAGS_EINVAL;
  g_peer_ctas_per_sm = ctas_per_sm;
  return 0;
}
// reductions in flight per thread of the multicast form: 2, 4 (default) or 8
extern "C" int gags_set_peer_unroll(int32_t unroll) {
  if (unroll != 2 && unroll != 4 && unroll != 8) return GAGS_EINVAL;
  g_peer_unroll = unroll;
  return 0;
}

extern "C" int gags_adam_step_peer(int32_t world, int32_t rank, const uint64_t *grad_ptrs,
                                   const uint64_t *param_ptrs, float *exp_avg_shard,
                                   float *exp_avg_sq_shard, int64_t start, int64_t count, double lr,
                                   double beta1, double beta2, double eps, int32_t step,
                                   void *stream) {
  if (!grad_ptrs || !param_ptrs || world < 1 || world > GAGS_MAX_PEERS || rank < 0 || rank >= world ||
      start < 0 || count < 0 || step < 1)
    return GAGS_EINVAL;
  if (count == 0) return 0;
  if (!exp_avg_shard || !exp_avg_sq_shard) return GAGS_EINVAL;
  if ((start & 3) || (count & 3)) return GAGS_EALIGN;          // float4 granularity
  if (!gags_aligned16(exp_avg_shard) || !gags_aligned16(exp_avg_sq_shard)) return GAGS_EALIGN;
  PeerPtrs pp;
  for (int q = 0; q < GAGS_MAX_PEERS; ++q) {
    pp.grad[q] = nullptr;
    pp.param[q] = nullptr;
  }
  for (int q = 0; q < world; ++q) {
    if (!grad_ptrs[q] || !param_ptrs[q] || (grad_ptrs[q] & 15) || (param_ptrs[q] & 15))
      return GAGS_EALIGN;
    pp.grad[q] = reinterpret_cast<const float4 *>(grad_ptrs[q]);
    pp.param[q] = reinterpret_cast<float4 *>(param_ptrs[q]);
  }
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
  const long long c4 = count / 4;
  long long blocks = (c4 + 255) / 256;
  // Default 2 CTAs/SM: the exchange runs beside the NEXT view's projection / tile sort on the main
  // stream; with 4 CTAs/SM the projection kernel was starved (2.9 ms instead of 0.1 ms in the 2-GPU
  // trace, profiles/r02b_trace_dp2.txt) and the forward behind it started 0.6 ms late.
  const long long cap_u = gags_sm_count() * (g_peer_ctas_per_sm > 0 ? g_peer_ctas_per_sm : 2);
  if (blocks > cap_u) blocks = cap_u;
#define GAGS_PEER_LAUNCH(MAXW)                                                                   \
  adam_peer_kernel<MAXW><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(                      \
      pp, world, rank, reinterpret_cast<float4 *>(exp_avg_shard),                                  \
      reinterpret_cast<float4 *>(exp_avg_sq_shard), start / 4, c4, step_size, (float)beta1,        \
      (float)beta2, omb1, omb2, inv_sqrt_bc2, (float)eps)
  if (world <= 2) GAGS_PEER_LAUNCH(2);
  else if (world <= 4) GAGS_PEER_LAUNCH(4);
  else if (world <= 8) GAGS_PEER_LAUNCH(8);
  else GAGS_PEER_LAUNCH(16);
#undef GAGS_PEER_LAUNCH
  GAGS_CHECK_LAUNCH();
  return 0;
}

extern "C" int gags_adam_step_multicast(const float *mc_grad, float *mc_param,
                                        const float *param_local, float *exp_avg_shard,
                                        float *exp_avg_sq_shard, int64_t start, int64_t count,
                                        double lr, double beta1, double beta2, double eps,
                                        int32_t step, void *stream) {
  if (!mc_grad || !mc_param || !param_local || start < 0 || count < 0 || step < 1) return GAGS_EINVAL;
  if (count == 0) return 0;
  if (!exp_avg_shard || !exp_avg_sq_shard) return GAGS_EINVAL;
  if ((start & 3) || (count & 3)) return GAGS_EALIGN;
  if (!gags_aligned16(mc_grad) || !gags_aligned16(mc_param) || !gags_aligned16(param_local) ||
      !gags_aligned16(exp_avg_shard) || !gags_aligned16(exp_avg_sq_shard))
    return GAGS_EALIGN;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
  const long long c4 = count / 4;
  long long blocks = (c4 + 255) / 256;
  // Default 1 CTA/SM: the 8-GPU sweep (profiles/r02a_peer_rate_8gpu.txt) shows the NVLS form at the
  // same 4.2-4.4 ms from 1 to 8 CTAs/SM (it is bound by the switch, not by requests in flight), and
  // the smallest grid leaves the SMs to the next view's projection / sort running beside it.
  const long long cap_m = gags_sm_count() * (g_peer_ctas_per_sm > 0 ? g_peer_ctas_per_sm : 1);
  if (blocks > cap_m) blocks = cap_m;
#define GAGS_MC_LAUNCH(U)                                                                        \
  adam_multicast_kernel<U><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(                   \
      reinterpret_cast<const float4 *>(mc_grad), reinterpret_cast<float4 *>(mc_param),             \
      reinterpret_cast<const float4 *>(param_local), reinterpret_cast<float4 *>(exp_avg_shard),    \
      reinterpret_cast<float4 *>(exp_avg_sq_shard), start / 4, c4, step_size, (float)beta1,        \
      (float)beta2, omb1, omb2, inv_sqrt_bc2, (float)eps)
  if (g_peer_unroll == 2) GAGS_MC_LAUNCH(2);
  else if (g_peer_unroll == 8) GAGS_MC_LAUNCH(8);
  else GAGS_MC_LAUNCH(4);
#undef GAGS_MC_LAUNCH
  GAGS_CHECK_LAUNCH();
  return 0;
}

// Row-sparse gradient all-reduce (grad_rows_allreduce_kernel).  grad_ptrs[q] / flag_ptrs[q]: every
// rank's [rows, D] gradient buffer and its row flags (uint8 [4 * words], words = ceil(rows / 4),
// zero padded), mapped into this process; mc_grad / mc_flags: their NVLS multicast addresses, or
// NULL for the unicast form.  union_flags (local, 4 * words bytes) receives the OR of all ranks'
// flags; the gradient rows flagged there hold the sum over ranks afterwards, on every rank, for
// the words [w0, w1) this rank owns (all ranks together cover [0, words)).
extern "C" int gags_grad_allreduce_rows(int32_t world, int32_t rank, const uint64_t *grad_ptrs,
                                        const uint64_t *flag_ptrs, float *mc_grad,
                                        const uint8_t *mc_flags, uint8_t *union_flags, int64_t rows,
                                        int32_t D, void *stream) {
  if (!grad_ptrs || !flag_ptrs || !union_flags || world < 1 || world > GAGS_MAX_PEERS || rank < 0 ||
      rank >= world || rows < 0 || D < 4 || (D % 4) != 0)
    return GAGS_EINVAL;
  if ((mc_grad == nullptr) != (mc_flags == nullptr)) return GAGS_EINVAL;
  if (rows == 0) return 0;
  RowPeerPtrs pp;
  for (int q = 0; q < GAGS_MAX_PEERS; ++q) {
    pp.grad[q] = nullptr;
    pp.flags[q] = nullptr;
  }
  for (int q = 0; q < world; ++q) {
    pp.grad[q] = reinterpret_cast<float4 *>(grad_ptrs[q]);
    pp.flags[q] = reinterpret_cast<const unsigned *>(flag_ptrs[q]);
    if (!pp.grad[q] || !pp.flags[q]) return GAGS_EINVAL;
    if (!gags_aligned16(pp.grad[q]) || (flag_ptrs[q] & 3)) return GAGS_EALIGN;
  }
  if (mc_grad && (!gags_aligned16(mc_grad) || ((uintptr_t)mc_flags & 3))) return GAGS_EALIGN;
  if ((uintptr_t)union_flags & 3) return GAGS_EALIGN;
  const long long words = (rows + 3) / 4;
  const long long per = (words + world - 1) / world;
  long long w0 = per * rank, w1 = per * (rank + 1);
  if (w0 > words) w0 = words;
  if (w1 > words) w1 = words;
  // latency bound (a flag round trip, then one round trip per flagged word, per warp): many warps,
  // each with little to do — the kernel is short, what runs beside it waits at most that long
  // (gags_set_peer_grid overrides, as for the dense exchange kernels)
  const long long blocks = (long long)gags_sm_count() * (g_peer_ctas_per_sm > 0 ? g_peer_ctas_per_sm : 8);
  cudaStream_t st = (cudaStream_t)stream;
#define GAGS_ROWS_AR(MC, MAXW)                                                                     \
  grad_rows_allreduce_kernel<MC, MAXW><<<(unsigned)blocks, 256, 0, st>>>(                           \
      pp, world, reinterpret_cast<float4 *>(mc_grad), reinterpret_cast<const unsigned *>(mc_flags), \
      reinterpret_cast<unsigned *>(union_flags), words, w0, w1, D / 4)
  if (mc_grad) GAGS_ROWS_AR(true, 1);
  else if (world <= 2) GAGS_ROWS_AR(false, 2);
  else if (world <= 4) GAGS_ROWS_AR(false, 4);
  else if (world <= 8) GAGS_ROWS_AR(false, 8);
  else GAGS_ROWS_AR(false, GAGS_MAX_PEERS);
#undef GAGS_ROWS_AR
  GAGS_CHECK_LAUNCH();
  return 0;
}

// Zero-fill with a SMALL grid (two 256-thread CTAs per SM): stores are fire-and-forget, so a quarter
// of the thread slots already streams at HBM rate, and the other three quarters stay free for the
// latency-bound kernels this fill is meant to run beside (projection / tile scatter / bucket sort at
// the start of a view; torch's fill kernel and cudaMemsetAsync both occupy every slot and simply
// push those kernels behind them).
__global__ void __launch_bounds__(256)
zero_fill_kernel(float4 *__restrict__ p, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    p[i] = z;
    p[i + stride] = z;
    p[i + 2 * stride] = z;
    p[i + 3 * stride] = z;
  }
  for (; i < n4; i += stride) p[i] = z;
}

extern "C" int gags_zero_fill(void *ptr, int64_t bytes, void *stream) {
  if ((!ptr && bytes) || bytes < 0) return GAGS_EINVAL;
  if (bytes == 0) return 0;
  if (!gags_aligned16(ptr) || (bytes & 15)) return GAGS_EALIGN;
  const long long n4 = bytes / 16;
  long long blocks = (n4 + 255) / 256;
  if (blocks > gags_sm_count() * 2) blocks = gags_sm_count() * 2;
  zero_fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float4 *>(ptr), n4);
  GAGS_CHECK_LAUNCH();
  return 0;
}

// Zero-fill through cudaMemsetAsync (a driver memset, not one of this library's kernels): used for
// the 2 GB gradient accumulation buffer so that the fill does not compete for SM slots with the
// kernels running on the other stream.
extern "C" int gags_memset_zero(void *ptr, size_t bytes, void *stream) {
  if (!ptr && bytes) return GAGS_EINVAL;
  if (bytes == 0) return 0;
  GAGS_CUDA(cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream));
  return 0;
}
