// blend_bwd.cu — K8 alpha-blend backward.  Replaces gsplat rasterize_to_pixels_bwd<CDIM> and the
// 32-channel chunk loop + slice/cat autograd glue around it (reached from
// /root/reference/gaussian_renderer/__init__.py:56-70 via train.py:174 loss.backward());
// semantics = SURVEY.md Appendix A.6.
//
// Kernels
//   blend_bwd_feat_wide<NJ>   feature-only VJP (frozen geometry — the only gradient the shipped
//       training loop consumes, scene/gaussian_model.py:192-206):
//           v_colors[g,:] += sum_px w(g,px) * v_render[px,:],   w = the forward weight,
//       front-to-back (no last_ids, no colour re-read).  One CTA = 16x8 half tile; the half tile of
//       v_render is staged once in shared memory by 128 row-wise bulk async copies; per batch of 32
//       Gaussians the weights are recomputed (phases A1/A2 of blend_common.cuh), non-contributing
//       Gaussians are compacted away, and a register-tiled [g x ch] += w^T [g x px] * v [px x ch]
//       product is flushed with 16-byte vector reductions (red.global.add.v4.f32).
//   blend_bwd_narrow<CDIM,FULL>  D <= 32: per-pixel back-to-front loop of App. A.6; per-Gaussian
//       partial sums are reduced across the warp with a transposing butterfly (CDIM-1 shuffles
//       instead of 5*CDIM) and accumulated per block in shared memory before global atomics.
//
// Roofline (feature-only): HBM — H*W*4D (v_render once) + N_vis*4D*2 (reduction target, read +
// write in L2/HBM) + 12*n_isects_read; the SIMT FMA pipe is the co-bound.
#include "blend_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// transposing warp reduction: every lane holds v[0..NV); afterwards lane l holds, in v[0], the
// warp-wide sum of channel `chan` (returned); lanes with (lane & dupmask) != 0 hold duplicates.
// ------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ int warp_reduce_scatter(float (&v)[NV], int lane) {
  static_assert((NV & (NV - 1)) == 0 && NV <= 32, "NV must be a power of two <= 32");
  int chan = 0;
  int n = NV;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (n > 1) {
      const int h = n >> 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int k = 0; k < NV / 2; ++k) {
        if (k < h) {
          const float send = upper ? v[k] : v[k + h];
          const float keep = upper ? v[k + h] : v[k];
          v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      if (upper) chan += h;
      n = h;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
  return chan;
}
template <int NV>
__device__ __forceinline__ unsigned reduce_dupmask() {
  // lane bits consumed by halving steps are the top log2(NV) bits (16, 8, ...); the remaining low
  // bits index duplicates.
  unsigned m = 31u, off = 16u;
  for (int n = NV; n > 1; n >>= 1) { m &= ~off; off >>= 1; }
  return m;
}

// ------------------------------------------------------------------------------------------------
// narrow backward (D <= CDIM <= 32), back to front.  FULL: also geometry/opacity gradients.
// ------------------------------------------------------------------------------------------------
constexpr int NB = 128;   // Gaussians per batch (static smem budget: 48 KB)

template <int CDIM, bool FULL>
__global__ void __launch_bounds__(256)
blend_bwd_narrow(const float4 *__restrict__ geom, const float *__restrict__ colors, int D,
                 const float *__restrict__ bg, int W, int H, int tile_w,
                 const int *__restrict__ offsets, const int *__restrict__ ids,
                 const float *__restrict__ render_alphas, const int *__restrict__ last_ids,
                 const float *__restrict__ v_render, const float *__restrict__ v_alphas,
                 float *__restrict__ v_means2d, float *__restrict__ v_conics,
                 float *__restrict__ v_opac, float *__restrict__ v_colors) {
  __shared__ int s_id[NB];
  __shared__ float4 s_g0[NB];
  __shared__ float2 s_g1[NB];
  __shared__ __align__(16) float s_feat[FULL ? NB * CDIM : 4];
  __shared__ float s_vcol[NB * CDIM];
  __shared__ float s_vgeo[FULL ? NB * 8 : 8];
  __shared__ int s_maxlast;
  const int tile = blockIdx.y * tile_w + blockIdx.x;
  const int tid = threadIdx.y * 16 + threadIdx.x;
  const int lane = tid & 31;
  const int x = blockIdx.x * GAGS_TILE + threadIdx.x;
  const int y = blockIdx.y * GAGS_TILE + threadIdx.y;
  const bool inside = (x < W) && (y < H);
  const float px = (float)x + 0.5f, py = (float)y + 0.5f;
  const int s = offsets[tile], e = offsets[tile + 1];
  if (e <= s) return;
  const size_t pix = inside ? (size_t)y * W + x : 0;
  const float T_final = inside ? 1.f - render_alphas[pix] : 1.f;
  const int my_last = inside ? last_ids[pix] : -1;
  float vo[CDIM];
#pragma unroll
  for (int c = 0; c < CDIM; ++c) vo[c] = (inside && c < D) ? v_render[pix * D + c] : 0.f;
  float bgdot = 0.f;
  if (FULL && bg) {
#pragma unroll
    for (int c = 0; c < CDIM; ++c) if (c < D) bgdot = fmaf(bg[c], vo[c], bgdot);
  }
  const float va = (FULL && inside && v_alphas) ? v_alphas[pix] : 0.f;
  if (tid == 0) s_maxlast = -1;
  __syncthreads();
  // a pixel with no contributor has last == 0 and T_final == 1; treating index 0 as a candidate is
  // harmless (its alpha test fails or its weight is exactly what the forward used).
  if (inside) atomicMax(&s_maxlast, my_last);
  __syncthreads();
  const int maxlast = s_maxlast;
  if (maxlast < s) return;
  float T = T_final;
  float S[FULL ? CDIM : 1];
#pragma unroll
  for (int c = 0; c < (FULL ? CDIM : 1); ++c) S[c] = 0.f;
  const int nbatch = (maxlast - s) / NB + 1;
  for (int bi = nbatch - 1; bi >= 0; --bi) {
    const int b0 = s + bi * NB;
    const int nb = min(NB, e - b0);
    __syncthreads();
    if (tid < nb) {
      const int id = ids[b0 + tid];
      s_id[tid] = id;
      const float4 r0 = geom[id * 2], r1 = geom[id * 2 + 1];
      s_g0[tid] = r0;
      s_g1[tid] = make_float2(r1.x, r1.y);
    }
    for (int i = tid; i < NB * CDIM; i += 256) s_vcol[i] = 0.f;
    if (FULL) for (int i = tid; i < NB * 8; i += 256) s_vgeo[i] = 0.f;
    __syncthreads();
    if (FULL) {
      for (int i = tid; i < nb * CDIM; i += 256) {
        const int g = i / CDIM, c = i - g * CDIM;
        s_feat[i] = (c < D) ? __ldg(colors + (size_t)s_id[g] * D + c) : 0.f;
      }
      __syncthreads();
    }
    const int jhi = min(nb - 1, maxlast - b0);
    for (int j = jhi; j >= 0; --j) {
      const float4 r0 = s_g0[j];
      const float2 r1 = s_g1[j];
      float a = 0.f, vis = 0.f, dx = 0.f, dy = 0.f;
      if (inside && b0 + j <= my_last) {
        dx = r0.x - px; dy = r0.y - py;
        const float sigma = 0.5f * (r0.z * dx * dx + r1.x * dy * dy) + r0.w * dx * dy;
        vis = __expf(-sigma);
        a = fminf(GAGS_ALPHA_MAX, r1.y * vis);
        if (sigma < 0.f || a < GAGS_ALPHA_MIN) a = 0.f;
      }
      if (__ballot_sync(0xffffffffu, a > 0.f) == 0u) continue;
      float vc[CDIM];
      float vg[8];
      float w = 0.f, ra = 1.f;
      if (a > 0.f) {
        ra = 1.f / (1.f - a);
        T *= ra;
        w = a * T;
      }
#pragma unroll
      for (int c = 0; c < CDIM; ++c) vc[c] = w * vo[c];
      if (FULL) {
#pragma unroll
        for (int k = 0; k < 8; ++k) vg[k] = 0.f;
        if (a > 0.f) {
          const float *f = s_feat + j * CDIM;
          float v_al = 0.f;
#pragma unroll
          for (int c = 0; c < CDIM; ++c) v_al = fmaf(f[c] * T - S[c] * ra, vo[c], v_al);
          v_al += T_final * ra * (va - bgdot);
          if (r1.y * vis <= GAGS_ALPHA_MAX) {
            const float v_sig = -r1.y * vis * v_al;
            vg[0] = 0.5f * v_sig * dx * dx;
            vg[1] = v_sig * dx * dy;
            vg[2] = 0.5f * v_sig * dy * dy;
            vg[3] = v_sig * (r0.z * dx + r0.w * dy);
            vg[4] = v_sig * (r0.w * dx + r1.x * dy);
            vg[5] = vis * v_al;
          }
#pragma unroll
          for (int c = 0; c < CDIM; ++c) S[c] = fmaf(f[c], w, S[c]);
        }
      }
      const int ch = warp_reduce_scatter<CDIM>(vc, lane);
      if ((lane & reduce_dupmask<CDIM>()) == 0) atomicAdd(&s_vcol[j * CDIM + ch], vc[0]);
      if (FULL) {
        const int k = warp_reduce_scatter<8>(vg, lane);
        if ((lane & reduce_dupmask<8>()) == 0 && k < 6) atomicAdd(&s_vgeo[j * 8 + k], vg[0]);
      }
    }
    __syncthreads();
    // flush the block's partial sums
    for (int i = tid; i < nb * D; i += 256) {
      const int g = i / D, c = i - g * D;
      const float v = s_vcol[g * CDIM + c];
      if (v != 0.f && v_colors) atomicAdd(v_colors + (size_t)s_id[g] * D + c, v);
    }
    if (FULL) {
      for (int i = tid; i < nb * 6; i += 256) {
        const int g = i / 6, k = i - g * 6;
        const float v = s_vgeo[g * 8 + k];
        if (v == 0.f) continue;
        const size_t id = (size_t)s_id[g];
        if (k < 3) atomicAdd(v_conics + id * 3 + k, v);
        else if (k < 5) atomicAdd(v_means2d + id * 2 + (k - 3), v);
        else atomicAdd(v_opac + id, v);
      }
    }
  }
}

// narrow feature-only backward, front to back (same chain as the forward; no alphas/last_ids)
template <int CDIM>
__global__ void __launch_bounds__(256)
blend_bwd_feat_narrow(const float4 *__restrict__ geom, int D, int W, int H, int tile_w,
                      const int *__restrict__ offsets, const int *__restrict__ ids,
                      const float *__restrict__ v_render, float *__restrict__ v_colors) {
  __shared__ int s_id[NB];
  __shared__ float4 s_g0[NB];
  __shared__ float2 s_g1[NB];
  __shared__ float s_vcol[NB * CDIM];
  const int tile = blockIdx.y * tile_w + blockIdx.x;
  const int tid = threadIdx.y * 16 + threadIdx.x;
  const int lane = tid & 31;
  const int x = blockIdx.x * GAGS_TILE + threadIdx.x;
  const int y = blockIdx.y * GAGS_TILE + threadIdx.y;
  const bool inside = (x < W) && (y < H);
  const float px = (float)x + 0.5f, py = (float)y + 0.5f;
  const int s = offsets[tile], e = offsets[tile + 1];
  if (e <= s) return;
  const size_t pix = inside ? (size_t)y * W + x : 0;
  float vo[CDIM];
#pragma unroll
  for (int c = 0; c < CDIM; ++c) vo[c] = (inside && c < D) ? v_render[pix * D + c] : 0.f;
  float T = 1.f;
  bool done = !inside;
  for (int b0 = s; b0 < e; b0 += NB) {
    const int nb = min(NB, e - b0);
    __syncthreads();
    if (tid < nb) {
      const int id = ids[b0 + tid];
      s_id[tid] = id;
      const float4 r0 = geom[id * 2], r1 = geom[id * 2 + 1];
      s_g0[tid] = r0;
      s_g1[tid] = make_float2(r1.x, r1.y);
    }
    for (int i = tid; i < NB * CDIM; i += 256) s_vcol[i] = 0.f;
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
      const float4 r0 = s_g0[j];
      const float2 r1 = s_g1[j];
      float w = 0.f;
      if (!done) {
        const float a = eval_alpha(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, px, py);
        if (a > 0.f) {
          const float Tn = T * (1.f - a);
          if (Tn <= GAGS_T_STOP) done = true;
          else { w = a * T; T = Tn; }
        }
      }
      if (__ballot_sync(0xffffffffu, w > 0.f) == 0u) continue;
      float vc[CDIM];
#pragma unroll
      for (int c = 0; c < CDIM; ++c) vc[c] = w * vo[c];
      const int ch = warp_reduce_scatter<CDIM>(vc, lane);
      if ((lane & reduce_dupmask<CDIM>()) == 0) atomicAdd(&s_vcol[j * CDIM + ch], vc[0]);
    }
    __syncthreads();
    for (int i = tid; i < nb * D; i += 256) {
      const int g = i / D, c = i - g * D;
      const float v = s_vcol[g * CDIM + c];
      if (v != 0.f) atomicAdd(v_colors + (size_t)s_id[g] * D + c, v);
    }
    if (__syncthreads_count(done) == 256) break;
  }
}

// ------------------------------------------------------------------------------------------------
// wide feature-only backward
// ------------------------------------------------------------------------------------------------
template <int NJ>
struct BwdSmem {
  static constexpr int PW = 64 * NJ;
  float vbuf[HP][PW];          // v_render half tile, pixel-major (our 4x4-block pixel order)
  float wbuf[2][WB * HP];
  float4 g0[2][WB];
  float4 g1[2][WB];
  int id[2][WB];
  unsigned masks[2][WB * 4];
  int clist[2][WB];
  int ccount[2];
  uint64_t mbar;
};

template <int NJ>
__global__ void __launch_bounds__(256, 1)
blend_bwd_feat_wide(const float4 *__restrict__ geom, int D, int ch0, int W, int H, int tile_w,
                    const int *__restrict__ offsets, const int *__restrict__ ids,
                    const float *__restrict__ v_render, float *__restrict__ v_colors) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdSmem<NJ> &sm = *reinterpret_cast<BwdSmem<NJ> *>(smem_raw);
  constexpr int PW = 64 * NJ;
  constexpr int GPT = 2 * NJ;            // Gaussians per thread
  constexpr int NCQ = 16 * NJ;           // float4 channel groups
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int tile = (blockIdx.y >> 1) * tile_w + blockIdx.x;
  const int x0 = blockIdx.x * GAGS_TILE, y0 = blockIdx.y * HROWS;
  const int p = tid & (HP - 1), half = tid >> 7;
  int pdx, pdy;
  hp_pixel(p, pdx, pdy);
  const int pxi = x0 + pdx, pyi = y0 + pdy;
  const bool inside = (pxi < W) && (pyi < H);
  const float px = (float)pxi + 0.5f, py = (float)pyi + 0.5f;
  const int nch = min(PW, D - ch0);
  const unsigned rowbytes = (unsigned)nch * 4u;
  const int s = offsets[tile], e = offsets[tile + 1];
  if (e <= s) return;
  const int nbatches = (e - s + WB - 1) / WB;
  // GEMM identity
  const int cq = tid % NCQ;              // float4 channel group
  const int gg = tid / NCQ;              // Gaussian group: compacted entries [gg*GPT, gg*GPT+GPT)

  if (tid == 0) { mbar_init(&sm.mbar, 1); mbar_fence_init(); }
  __syncthreads();
  // stage v_render: one bulk copy per in-image pixel row of `nch` floats; zero-fill the rest
  if (tid == 0) {
    int nin = 0;
    for (int q = 0; q < HP; ++q) {
      int dx, dy; hp_pixel(q, dx, dy);
      nin += ((x0 + dx) < W && (y0 + dy) < H) ? 1 : 0;
    }
    mbar_expect_tx(&sm.mbar, (unsigned)nin * rowbytes);
  }
  if (tid < HP) {
    if (!inside) {
      for (int c = 0; c < PW; ++c) sm.vbuf[p][c] = 0.f;
    }
  }
  __syncthreads();
  if (tid < HP && inside)
    bulk_g2s(&sm.vbuf[p][0], v_render + ((size_t)pyi * W + pxi) * D + ch0, rowbytes, &sm.mbar);

  PixelState st;
  st.T = 1.f; st.last = 0; st.done = inside ? 0 : 1;
  if (warp == 7) {
    const int nb = min(WB, e - s);
    if (lane < nb) {
      const int id = ids[s + lane];
      sm.id[0][lane] = id;
      sm.g0[0][lane] = geom[id * 2];
      sm.g1[0][lane] = geom[id * 2 + 1];
    }
  }
  __syncthreads();
  bool vready = false;
  bool all_done = false;
  for (int i = 0; i < nbatches && !all_done; ++i) {
    const int b = i & 1;
    const int base = s + i * WB;
    const int nb = min(WB, e - base);
    int nid = 0; float4 n0, n1; bool pf = false;
    if (warp == 7 && i + 1 < nbatches) {
      const int nnb = min(WB, e - base - WB);
      if (lane < nnb) { nid = ids[base + WB + lane]; n0 = geom[nid * 2]; n1 = geom[nid * 2 + 1]; pf = true; }
    }
    phase_a1(sm.g0[b], sm.g1[b], nb, half, p, px, py, inside, sm.wbuf[b]);
    __syncthreads();
    if (pf) { sm.id[b ^ 1][lane] = nid; sm.g0[b ^ 1][lane] = n0; sm.g1[b ^ 1][lane] = n1; }
    if (tid < HP) phase_a2(sm.wbuf[b], sm.masks[b], nb, p, base, st);
    all_done = (__syncthreads_count(st.done || tid >= HP) == 256);
    if (warp == 0) {
      bool any = false;
      if (lane < nb) {
        const uint4 m = *reinterpret_cast<const uint4 *>(&sm.masks[b][lane * 4]);
        any = (m.x | m.y | m.z | m.w) != 0u;
      }
      const unsigned ball = __ballot_sync(0xffffffffu, any);
      if (lane == 0) sm.ccount[b] = __popc(ball);
      if (any) sm.clist[b][__popc(ball & ((1u << lane) - 1u))] = lane;
    }
    __syncthreads();
    const int cnt = sm.ccount[b];
    if (cnt > 0) {
      if (!vready) { mbar_wait(&sm.mbar, 0); vready = true; }
      if (gg * GPT < cnt && cq * 4 < nch) {
        float acc[GPT][4];
        int grow[GPT];
#pragma unroll
        for (int k = 0; k < GPT; ++k) {
          acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
          const int c = gg * GPT + k;
          grow[k] = sm.clist[b][c < cnt ? c : cnt - 1];
        }
        const float *wb = sm.wbuf[b];
#pragma unroll 2
        for (int q = 0; q < HP; q += 4) {
          float4 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const float4 *>(&sm.vbuf[q + k][cq * 4]);
#pragma unroll
          for (int k = 0; k < GPT; ++k) {
            const float4 w = *reinterpret_cast<const float4 *>(&wb[grow[k] * HP + q]);
            acc[k][0] = fmaf(w.x, v[0].x, acc[k][0]); acc[k][1] = fmaf(w.x, v[0].y, acc[k][1]);
            acc[k][2] = fmaf(w.x, v[0].z, acc[k][2]); acc[k][3] = fmaf(w.x, v[0].w, acc[k][3]);
            acc[k][0] = fmaf(w.y, v[1].x, acc[k][0]); acc[k][1] = fmaf(w.y, v[1].y, acc[k][1]);
            acc[k][2] = fmaf(w.y, v[1].z, acc[k][2]); acc[k][3] = fmaf(w.y, v[1].w, acc[k][3]);
            acc[k][0] = fmaf(w.z, v[2].x, acc[k][0]); acc[k][1] = fmaf(w.z, v[2].y, acc[k][1]);
            acc[k][2] = fmaf(w.z, v[2].z, acc[k][2]); acc[k][3] = fmaf(w.z, v[2].w, acc[k][3]);
            acc[k][0] = fmaf(w.w, v[3].x, acc[k][0]); acc[k][1] = fmaf(w.w, v[3].y, acc[k][1]);
            acc[k][2] = fmaf(w.w, v[3].z, acc[k][2]); acc[k][3] = fmaf(w.w, v[3].w, acc[k][3]);
          }
        }
#pragma unroll
        for (int k = 0; k < GPT; ++k) {
          if (gg * GPT + k < cnt) {
            float *dst = v_colors + (size_t)sm.id[b][grow[k]] * D + ch0 + cq * 4;
            red_add4(dst, make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]));
          }
        }
      }
    }
    __syncthreads();
  }
  // never leave with the staging copies in flight
  if (!vready) mbar_wait(&sm.mbar, 0);
}

template <int CDIM, bool FULL>
int launch_bwd_narrow(const float *geom, const float *colors, int D, const float *bg, int W, int H,
                      const int *offsets, const int *ids, const float *ra, const int *last_ids,
                      const float *v_render, const float *v_alphas, float *v_m, float *v_c,
                      float *v_o, float *v_col, cudaStream_t st) {
  const int tw = (W + GAGS_TILE - 1) / GAGS_TILE, th = (H + GAGS_TILE - 1) / GAGS_TILE;
  blend_bwd_narrow<CDIM, FULL><<<dim3(tw, th), dim3(16, 16), 0, st>>>(
      reinterpret_cast<const float4 *>(geom), colors, D, bg, W, H, tw, offsets, ids, ra, last_ids,
      v_render, v_alphas, v_m, v_c, v_o, v_col);
  return (int)cudaGetLastError();
}

template <int CDIM>
int launch_bwd_feat_narrow(const float *geom, int D, int W, int H, const int *offsets,
                           const int *ids, const float *v_render, float *v_colors, cudaStream_t st) {
  const int tw = (W + GAGS_TILE - 1) / GAGS_TILE, th = (H + GAGS_TILE - 1) / GAGS_TILE;
  blend_bwd_feat_narrow<CDIM><<<dim3(tw, th), dim3(16, 16), 0, st>>>(
      reinterpret_cast<const float4 *>(geom), D, W, H, tw, offsets, ids, v_render, v_colors);
  return (int)cudaGetLastError();
}

template <int NJ>
int launch_bwd_feat_wide(const float *geom, int D, int ch0, int W, int H, const int *offsets,
                         const int *ids, const float *v_render, float *v_colors, cudaStream_t st) {
  const int tw = (W + GAGS_TILE - 1) / GAGS_TILE;
  const int hh = (H + HROWS - 1) / HROWS;
  const size_t smem = sizeof(BwdSmem<NJ>);
  {   // per-device attribute: set on every launch (a process may drive several GPUs)
    cudaError_t e = cudaFuncSetAttribute(blend_bwd_feat_wide<NJ>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  blend_bwd_feat_wide<NJ><<<dim3(tw, hh), 256, smem, st>>>(reinterpret_cast<const float4 *>(geom), D,
                                                          ch0, W, H, tw, offsets, ids, v_render,
                                                          v_colors);
  return (int)cudaGetLastError();
}

template <bool FULL>
int dispatch_narrow(const float *geom, const float *colors, int D, const float *bg, int W, int H,
                    const int *offsets, const int *ids, const float *ra, const int *last_ids,
                    const float *v_render, const float *v_alphas, float *v_m, float *v_c, float *v_o,
                    float *v_col, cudaStream_t st) {
  if (D <= 4) return launch_bwd_narrow<4, FULL>(geom, colors, D, bg, W, H, offsets, ids, ra, last_ids, v_render, v_alphas, v_m, v_c, v_o, v_col, st);
  if (D <= 8) return launch_bwd_narrow<8, FULL>(geom, colors, D, bg, W, H, offsets, ids, ra, last_ids, v_render, v_alphas, v_m, v_c, v_o, v_col, st);
  if (D <= 16) return launch_bwd_narrow<16, FULL>(geom, colors, D, bg, W, H, offsets, ids, ra, last_ids, v_render, v_alphas, v_m, v_c, v_o, v_col, st);
  return launch_bwd_narrow<32, FULL>(geom, colors, D, bg, W, H, offsets, ids, ra, last_ids, v_render, v_alphas, v_m, v_c, v_o, v_col, st);
}

}  // namespace

// defined in blend_bwd_geom.cu
int gags_blend_bwd_geom_wide(const float *geom, const float *colors, int32_t D,
                             const float *background, int32_t width, int32_t height,
                             const int32_t *offsets, const int32_t *flatten_ids,
                             const float *render_alphas, const int32_t *last_ids,
                             const float *v_render, const float *v_alphas, float *v_means2d,
                             float *v_conics, float *v_opacities, cudaStream_t st);

// defined in blend_bwd_tc.cu / api.cu
int gags_blend_bwd_features_tc(const float *geom, int32_t D, int32_t width, int32_t height,
                               const int32_t *offsets, const int32_t *flatten_ids,
                               const float *v_render, float *v_colors, cudaStream_t st);
extern int g_gags_blend_impl;

extern "C" int gags_blend_bwd_features(const float *geom, int32_t D, int32_t width, int32_t height,
                                       const int32_t *offsets, const int32_t *flatten_ids,
                                       const float *v_render, float *v_colors, void *stream) {
  if (!geom || !offsets || !v_render || !v_colors) return GAGS_EINVAL;
  if (D < 1 || width <= 0 || height <= 0) return GAGS_EINVAL;
  if (!gags_aligned16(geom)) return GAGS_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  if (D <= 32) {
    if (D <= 4) return launch_bwd_feat_narrow<4>(geom, D, width, height, offsets, flatten_ids, v_render, v_colors, st);
    if (D <= 8) return launch_bwd_feat_narrow<8>(geom, D, width, height, offsets, flatten_ids, v_render, v_colors, st);
    if (D <= 16) return launch_bwd_feat_narrow<16>(geom, D, width, height, offsets, flatten_ids, v_render, v_colors, st);
    return launch_bwd_feat_narrow<32>(geom, D, width, height, offsets, flatten_ids, v_render, v_colors, st);
  }
  if (D % 4 != 0) return GAGS_EINVAL;
  if (!gags_aligned16(v_render) || !gags_aligned16(v_colors)) return GAGS_EALIGN;
  if (g_gags_blend_impl != 1 && D % 16 == 0)
    return gags_blend_bwd_features_tc(geom, D, width, height, offsets, flatten_ids, v_render,
                                      v_colors, st);
  if (g_gags_blend_impl == 2) return GAGS_EINVAL;
  for (int ch0 = 0; ch0 < D; ch0 += 256) {
    const int nch = (D - ch0) < 256 ? (D - ch0) : 256;
    const int nj = (nch + 63) / 64;
    int rc;
    switch (nj) {
      case 1: rc = launch_bwd_feat_wide<1>(geom, D, ch0, width, height, offsets, flatten_ids, v_render, v_colors, st); break;
      case 2: rc = launch_bwd_feat_wide<2>(geom, D, ch0, width, height, offsets, flatten_ids, v_render, v_colors, st); break;
      default: rc = launch_bwd_feat_wide<4>(geom, D, ch0, width, height, offsets, flatten_ids, v_render, v_colors, st); break;
    }
    if (rc != 0) return rc;
  }
  return 0;
}

extern "C" int gags_blend_bwd_full(const float *geom, const float *colors, int32_t D,
                                   const float *background, int32_t width, int32_t height,
                                   const int32_t *offsets, const int32_t *flatten_ids,
                                   const float *render_alphas, const int32_t *last_ids,
                                   const float *v_render, const float *v_alphas, float *v_means2d,
                                   float *v_conics, float *v_opacities, float *v_colors,
                                   void *stream) {
  if (!geom || !offsets || !render_alphas || !last_ids || !v_render) return GAGS_EINVAL;
  if (D < 1 || width <= 0 || height <= 0) return GAGS_EINVAL;
  if (!gags_aligned16(geom)) return GAGS_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const bool want_geo = v_means2d || v_conics || v_opacities;
  if (want_geo && !(v_means2d && v_conics && v_opacities && colors)) return GAGS_EINVAL;
  if (D <= 32) {
    if (want_geo)
      return dispatch_narrow<true>(geom, colors, D, background, width, height, offsets, flatten_ids,
                                   render_alphas, last_ids, v_render, v_alphas, v_means2d, v_conics,
                                   v_opacities, v_colors, st);
    if (!v_colors) return 0;
    return gags_blend_bwd_features(geom, D, width, height, offsets, flatten_ids, v_render, v_colors,
                                   stream);
  }
  if (v_colors) {
    int rc = gags_blend_bwd_features(geom, D, width, height, offsets, flatten_ids, v_render,
                                     v_colors, stream);
    if (rc != 0) return rc;
  }
  if (want_geo)
    return gags_blend_bwd_geom_wide(geom, colors, D, background, width, height, offsets, flatten_ids,
                                    render_alphas, last_ids, v_render, v_alphas, v_means2d, v_conics,
                                    v_opacities, st);
  return 0;
}
