// blend_bwd_geom.cu — K8b for wide features (D > 32, in channel blocks of <= 256): gradients of the blend w.r.t. the
// projected geometry and opacity (v_means2d, v_conics, v_opacities), SURVEY.md Appendix A.6.
// Replaces the geometry half of gsplat rasterize_to_pixels_bwd<CDIM> (x ceil(D/32) chunk launches
// in the reference, /root/reference/gaussian_renderer/__init__.py:56-70 -> train.py:174).
// v_colors comes from blend_bwd_feat_wide; the shipped GAGS loop (frozen geometry) never runs this.
//
// A.6 needs S[d] = sum_{j behind k} c_j[d] w_j per pixel — D running sums.  Only its contraction
// with v_out is used, so the kernel carries ONE scalar per pixel instead:
//     v_alpha_k = T_k * dot_k - ra_k * Sdot + T_final * ra_k * (v_alpha_out - bg . v_out)
//     dot_k = c_k . v_out[px]          Sdot = sum_{j behind k} w_j dot_j
// One CTA = 16x8 half tile, one thread per pixel, back to front in batches of 32 Gaussians:
// v_out rows live in shared memory (padded stride, conflict-free float4 reads), the batch's
// feature rows arrive by bulk async copies, dot products are register-blocked 1 pixel x 4 Gaussians
// and skipped per warp when no lane sees the Gaussian.
#include "blend_common.cuh"

namespace {

constexpr int GT = 128;   // threads per CTA (= pixels)

template <int NV>
__device__ __forceinline__ int warp_reduce_scatter8(float (&v)[NV], int lane) {
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

__global__ void __launch_bounds__(GT)
blend_bwd_geom_wide(const float4 *__restrict__ geom, const float *__restrict__ colors, int D,
                    int ld, const float *__restrict__ bg, int W, int H, int tile_w,
                    const int *__restrict__ offsets, const int *__restrict__ ids,
                    const float *__restrict__ render_alphas, const int *__restrict__ last_ids,
                    const float *__restrict__ v_render, const float *__restrict__ v_alphas,
                    float *__restrict__ v_means2d, float *__restrict__ v_conics,
                    float *__restrict__ v_opac) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int VS = D + 4;                                   // padded v_out row stride (floats)
  float *vbuf = reinterpret_cast<float *>(smem_raw);      // [GT][VS]
  float *fbuf = vbuf + GT * VS;                           // [WB][D]
  float4 *s_g0 = reinterpret_cast<float4 *>(fbuf + WB * D);
  float4 *s_g1 = s_g0 + WB;
  float *s_vgeo = reinterpret_cast<float *>(s_g1 + WB);   // [WB][8]
  int *s_id = reinterpret_cast<int *>(s_vgeo + WB * 8);
  int *s_maxlast = s_id + WB;
  uint64_t *mbar = reinterpret_cast<uint64_t *>(s_maxlast + 2);

  const int tid = threadIdx.x, lane = tid & 31;
  const int tile = (blockIdx.y >> 1) * tile_w + blockIdx.x;
  const int x0 = blockIdx.x * GAGS_TILE, y0 = blockIdx.y * HROWS;
  int pdx, pdy;
  hp_pixel(tid, pdx, pdy);
  const int pxi = x0 + pdx, pyi = y0 + pdy;
  const bool inside = (pxi < W) && (pyi < H);
  const float px = (float)pxi + 0.5f, py = (float)pyi + 0.5f;
  const int s = offsets[tile], e = offsets[tile + 1];
  if (e <= s) return;
  const size_t pix = inside ? (size_t)pyi * W + pxi : 0;
  const float T_final = inside ? 1.f - render_alphas[pix] : 1.f;
  const int my_last = inside ? last_ids[pix] : -1;
  const float va = (inside && v_alphas) ? v_alphas[pix] : 0.f;
  const unsigned rowbytes = (unsigned)D * 4u;

  if (tid == 0) { *s_maxlast = -1; mbar_init(mbar, 1); mbar_fence_init(); }
  __syncthreads();
  if (inside) atomicMax(s_maxlast, my_last);
  // stage this pixel's v_out row
  if (tid == 0) {
    int nin = 0;
    for (int q = 0; q < GT; ++q) {
      int dx, dy; hp_pixel(q, dx, dy);
      nin += ((x0 + dx) < W && (y0 + dy) < H) ? 1 : 0;
    }
    mbar_expect_tx(mbar, (unsigned)nin * rowbytes);
  }
  if (!inside) for (int c = 0; c < D; ++c) vbuf[tid * VS + c] = 0.f;
  __syncthreads();
  if (inside) bulk_g2s(vbuf + tid * VS, v_render + pix * ld, rowbytes, mbar);
  unsigned phase = 0;
  mbar_wait(mbar, phase); phase ^= 1u;
  const int maxlast = *s_maxlast;
  if (maxlast < s) return;
  const float *vrow = vbuf + tid * VS;
  float bgdot = 0.f;
  if (bg) for (int c = 0; c < D; ++c) bgdot = fmaf(bg[c], vrow[c], bgdot);

  float T = T_final, Sdot = 0.f;
  const int nbatch = (maxlast - s) / WB + 1;
  for (int bi = nbatch - 1; bi >= 0; --bi) {
    const int b0 = s + bi * WB;
    const int nb = min(WB, e - b0);
    __syncthreads();                                      // previous batch fully consumed
    if (tid < nb) {
      const int id = ids[b0 + tid];
      s_id[tid] = id;
      s_g0[tid] = geom[id * 2];
      s_g1[tid] = geom[id * 2 + 1];
    }
    for (int i = tid; i < WB * 8; i += GT) s_vgeo[i] = 0.f;
    if (tid == 0) mbar_expect_tx(mbar, (unsigned)nb * rowbytes);
    __syncthreads();
    if (tid < nb) bulk_g2s(fbuf + tid * D, colors + (size_t)s_id[tid] * ld, rowbytes, mbar);
    // alphas for the batch (registers), while the feature rows are in flight
    float al[WB], vi[WB];
#pragma unroll
    for (int g = 0; g < WB; ++g) {
      float a = 0.f, vis = 0.f;
      if (g < nb && inside && b0 + g <= my_last) {
        const float4 r0 = s_g0[g];
        const float4 r1 = s_g1[g];
        const float dx = r0.x - px, dy = r0.y - py;
        const float sigma = 0.5f * (r0.z * dx * dx + r1.x * dy * dy) + r0.w * dx * dy;
        vis = __expf(-sigma);
        a = fminf(GAGS_ALPHA_MAX, r1.y * vis);
        if (sigma < 0.f || a < GAGS_ALPHA_MIN) a = 0.f;
      }
      al[g] = a; vi[g] = vis;
    }
    mbar_wait(mbar, phase); phase ^= 1u;
    // dot products, 1 pixel x 4 Gaussians per pass over the channels
    float dot[WB];
#pragma unroll
    for (int g4 = 0; g4 < WB; g4 += 4) {
      dot[g4] = dot[g4 + 1] = dot[g4 + 2] = dot[g4 + 3] = 0.f;
      const bool mine = (al[g4] > 0.f) || (al[g4 + 1] > 0.f) || (al[g4 + 2] > 0.f) || (al[g4 + 3] > 0.f);
      if (__ballot_sync(0xffffffffu, mine) == 0u) continue;
      const float *f0 = fbuf + g4 * D;
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      for (int c = 0; c < D; c += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(vrow + c);
        const float4 a0 = *reinterpret_cast<const float4 *>(f0 + c);
        const float4 a1 = *reinterpret_cast<const float4 *>(f0 + D + c);
        const float4 a2 = *reinterpret_cast<const float4 *>(f0 + 2 * D + c);
        const float4 a3 = *reinterpret_cast<const float4 *>(f0 + 3 * D + c);
        d0 = fmaf(a0.x, v.x, d0); d0 = fmaf(a0.y, v.y, d0); d0 = fmaf(a0.z, v.z, d0); d0 = fmaf(a0.w, v.w, d0);
        d1 = fmaf(a1.x, v.x, d1); d1 = fmaf(a1.y, v.y, d1); d1 = fmaf(a1.z, v.z, d1); d1 = fmaf(a1.w, v.w, d1);
        d2 = fmaf(a2.x, v.x, d2); d2 = fmaf(a2.y, v.y, d2); d2 = fmaf(a2.z, v.z, d2); d2 = fmaf(a2.w, v.w, d2);
        d3 = fmaf(a3.x, v.x, d3); d3 = fmaf(a3.y, v.y, d3); d3 = fmaf(a3.z, v.z, d3); d3 = fmaf(a3.w, v.w, d3);
      }
      dot[g4] = d0; dot[g4 + 1] = d1; dot[g4 + 2] = d2; dot[g4 + 3] = d3;
    }
    // sequential back-to-front pass
#pragma unroll
    for (int g = WB - 1; g >= 0; --g) {
      const float a = al[g];
      if (__ballot_sync(0xffffffffu, a > 0.f) == 0u) continue;
      float vg[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) vg[k] = 0.f;
      if (a > 0.f) {
        const float4 r0 = s_g0[g];
        const float4 r1 = s_g1[g];
        const float dx = r0.x - px, dy = r0.y - py;
        const float ra = 1.f / (1.f - a);
        T *= ra;
        const float w = a * T;
        const float v_al = T * dot[g] - ra * Sdot + T_final * ra * (va - bgdot);
        const float ov = r1.y * vi[g];
        if (ov <= GAGS_ALPHA_MAX) {
          const float v_sig = -ov * v_al;
          vg[0] = 0.5f * v_sig * dx * dx;
          vg[1] = v_sig * dx * dy;
          vg[2] = 0.5f * v_sig * dy * dy;
          vg[3] = v_sig * (r0.z * dx + r0.w * dy);
          vg[4] = v_sig * (r0.w * dx + r1.x * dy);
          vg[5] = vi[g] * v_al;
        }
        Sdot = fmaf(w, dot[g], Sdot);
      }
      const int k = warp_reduce_scatter8<8>(vg, lane);
      if ((lane & 3) == 0 && k < 6) atomicAdd(&s_vgeo[g * 8 + k], vg[0]);
    }
    __syncthreads();
    for (int i = tid; i < nb * 6; i += GT) {
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

}  // namespace

int gags_blend_bwd_geom_wide(const float *geom, const float *colors, int32_t D,
                             const float *background, int32_t width, int32_t height,
                             const int32_t *offsets, const int32_t *flatten_ids,
                             const float *render_alphas, const int32_t *last_ids,
                             const float *v_render, const float *v_alphas, float *v_means2d,
                             float *v_conics, float *v_opacities, cudaStream_t st) {
  if (D % 4 != 0 || D <= 32) return GAGS_ERANGE;
  if (!gags_aligned16(colors) || !gags_aligned16(v_render)) return GAGS_EALIGN;
  const int tw = (width + GAGS_TILE - 1) / GAGS_TILE;
  const int hh = (height + HROWS - 1) / HROWS;
  // The geometry gradients are linear in the channel contractions (dot_k, Sdot, bg . v_out), so a
  // feature width beyond what one CTA's shared memory holds is processed in channel blocks of
  // <= 256 that all accumulate into the same outputs; the channel-free v_alpha_out term rides with
  // the first block only.  (D <= 513 is what gsplat compiles; config 5 uses D = 512.)
  for (int ch0 = 0; ch0 < D; ch0 += 256) {
    const int nch = (D - ch0) < 256 ? (D - ch0) : 256;
    const size_t smem = (size_t)(GT * (nch + 4) + WB * nch) * 4 + WB * 32 + WB * 32 + WB * 4 + 8 + 16;
    cudaError_t e = cudaFuncSetAttribute(blend_bwd_geom_wide,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    blend_bwd_geom_wide<<<dim3(tw, hh), GT, smem, st>>>(
        reinterpret_cast<const float4 *>(geom), colors + ch0, nch, D,
        background ? background + ch0 : nullptr, width, height, tw, offsets, flatten_ids,
        render_alphas, last_ids, v_render + ch0, ch0 == 0 ? v_alphas : nullptr, v_means2d,
        v_conics, v_opacities);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}
