// blend_fwd.cu — K7 alpha-blend forward.  Replaces gsplat rasterize_to_pixels_fwd<CDIM> plus the
// channel_chunk=32 loop and torch.cat around it (reached from
// /root/reference/gaussian_renderer/__init__.py:56-70); semantics = SURVEY.md Appendix A.5.
//
// Two kernels:
//   blend_fwd_narrow<CDIM>  D <= 32 : one thread per pixel, CDIM register accumulators, the batch's
//                           feature rows staged in shared memory (RGB, RGB+ED, D=16, D=32 cases).
//   blend_fwd_wide<NJ>      D  > 32 : ONE launch for up to 256 channels (no 32-channel chunking);
//                           see blend_common.cuh for the decomposition.  Feature rows arrive by
//                           per-row bulk async copies (cp.async.bulk -> UBLKCP) completing on an
//                           mbarrier, double-buffered one batch ahead of the FMA loop; the raster
//                           is written channel-last with 16-byte streaming stores (256 B / half warp).
//
// Roofline: HBM is the compulsory bound — algorithmic bytes per launch
//   N_vis*4D (feature rows, once) + H*W*(4D+8) (render + alpha + last_ids) + 12*n_isects_read;
// the SIMT FMA pipe is the co-bound: dense work = 2*128*D flop per (half tile, g).
#include "blend_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// narrow: D <= CDIM <= 32
// ------------------------------------------------------------------------------------------------
constexpr int NB = 256;  // Gaussians staged per batch (= threads per block)

template <int CDIM>
__global__ void __launch_bounds__(256)
blend_fwd_narrow(const float4 *__restrict__ geom, const float *__restrict__ colors, int D,
                 const float *__restrict__ bg, int W, int H, int tile_w,
                 const int *__restrict__ offsets, const int *__restrict__ ids,
                 float *__restrict__ render, float *__restrict__ alphas, int *__restrict__ last_ids) {
  __shared__ int s_id[NB];
  __shared__ float4 s_g0[NB];
  __shared__ float2 s_g1[NB];
  __shared__ __align__(16) float s_feat[NB * CDIM];
  const int tile = blockIdx.y * tile_w + blockIdx.x;
  const int tid = threadIdx.y * 16 + threadIdx.x;
  const int x = blockIdx.x * GAGS_TILE + threadIdx.x;
  const int y = blockIdx.y * GAGS_TILE + threadIdx.y;
  const bool inside = (x < W) && (y < H);
  const float px = (float)x + 0.5f, py = (float)y + 0.5f;
  const int s = offsets[tile], e = offsets[tile + 1];
  float acc[CDIM];
#pragma unroll
  for (int c = 0; c < CDIM; ++c) acc[c] = 0.f;
  float T = 1.f;
  int last = 0;
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
    __syncthreads();
    for (int i = tid; i < nb * D; i += 256) {
      const int g = i / D, c = i - g * D;
      s_feat[g * CDIM + c] = __ldg(colors + (size_t)s_id[g] * D + c);
    }
    __syncthreads();
    if (!done) {
      for (int j = 0; j < nb; ++j) {
        const float4 r0 = s_g0[j];
        const float2 r1 = s_g1[j];
        const float a = eval_alpha(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, px, py);
        if (a == 0.f) continue;
        const float Tn = T * (1.f - a);
        if (Tn <= GAGS_T_STOP) { done = true; break; }
        const float w = a * T;
        const float *f = s_feat + j * CDIM;
#pragma unroll
        for (int c = 0; c < CDIM; ++c) acc[c] = fmaf(w, f[c], acc[c]);
        last = b0 + j;
        T = Tn;
      }
    }
    if (__syncthreads_count(done) == 256) break;
  }
  if (inside) {
    const size_t pix = (size_t)y * W + x;
    alphas[pix] = 1.f - T;
    last_ids[pix] = last;
    float *o = render + pix * D;
    for (int c = 0; c < CDIM; ++c)
      if (c < D) o[c] = acc[c] + T * (bg ? bg[c] : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// wide: 32 < D, D % 4 == 0, up to 64*NJ channels per launch starting at channel ch0
// ------------------------------------------------------------------------------------------------
template <int NJ>
struct WideSmem {
  static constexpr int PW = 64 * NJ;  // floats per staged feature row
  float fbuf[2][WB][PW];
  float wbuf[2][WB * HP];
  float4 g0[2][WB];
  float4 g1[2][WB];
  int id[2][WB];
  unsigned masks[2][WB * 4];
  int clist[2][WB];
  int ccount[2];
  float Tfin[HP];
  uint64_t mbar[2];
};

template <int NJ>
__global__ void __launch_bounds__(256, 1)
blend_fwd_wide(const float4 *__restrict__ geom, const float *__restrict__ colors, int D, int ch0,
               const float *__restrict__ bg, int W, int H, int tile_w,
               const int *__restrict__ offsets, const int *__restrict__ ids,
               float *__restrict__ render, float *__restrict__ alphas, int *__restrict__ last_ids) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WideSmem<NJ> &sm = *reinterpret_cast<WideSmem<NJ> *>(smem_raw);
  constexpr int PW = 64 * NJ;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int tile = (blockIdx.y >> 1) * tile_w + blockIdx.x;
  const int x0 = blockIdx.x * GAGS_TILE, y0 = blockIdx.y * HROWS;
  // pixel-phase identity
  const int p = tid & (HP - 1), half = tid >> 7;
  int pdx, pdy;
  hp_pixel(p, pdx, pdy);
  const int pxi = x0 + pdx, pyi = y0 + pdy;
  const bool inside = (pxi < W) && (pyi < H);
  const float px = (float)pxi + 0.5f, py = (float)pyi + 0.5f;
  // accumulate-phase identity
  const int cg = lane & 15;
  const int pg = warp * 2 + (lane >> 4);
  const int nch = min(PW, D - ch0);          // channels of this pass
  const unsigned rowbytes = (unsigned)nch * 4u;

  const int s = offsets[tile], e = offsets[tile + 1];
  const int nbatches = (e - s + WB - 1) / WB;

  if (tid == 0) {
    mbar_init(&sm.mbar[0], 1);
    mbar_init(&sm.mbar[1], 1);
    mbar_fence_init();
  }
  float acc[8][4 * NJ];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int c = 0; c < 4 * NJ; ++c) acc[k][c] = 0.f;

  PixelState st;
  st.T = 1.f; st.last = 0; st.done = inside ? 0 : 1;
  unsigned phase0 = 0, phase1 = 0;

  // prologue: geometry of batch 0
  if (warp == 7 && nbatches > 0) {
    const int nb = min(WB, e - s);
    if (lane < nb) {
      const int id = ids[s + lane];
      sm.id[0][lane] = id;
      sm.g0[0][lane] = geom[id * 2];
      sm.g1[0][lane] = geom[id * 2 + 1];
    }
  }
  __syncthreads();

  bool all_done = false;
  for (int i = 0;; ++i) {
    const bool have_cur = (i < nbatches) && !all_done;
    const int b = i & 1;
    if (have_cur) {
      const int base = s + i * WB;
      const int nb = min(WB, e - base);
      // prefetch next batch's geometry (global latency overlaps phase A1 of the other warps)
      int nid = 0; float4 n0, n1; bool pf = false;
      if (warp == 7 && i + 1 < nbatches) {
        const int nnb = min(WB, e - base - WB);
        if (lane < nnb) { nid = ids[base + WB + lane]; n0 = geom[nid * 2]; n1 = geom[nid * 2 + 1]; pf = true; }
      }
      phase_a1(sm.g0[b], sm.g1[b], nb, half, p, px, py, inside, sm.wbuf[b]);
      if (pf) { sm.id[b ^ 1][lane] = nid; sm.g0[b ^ 1][lane] = n0; sm.g1[b ^ 1][lane] = n1; }
      __syncthreads();
      if (tid < HP) phase_a2(sm.wbuf[b], sm.masks[b], nb, p, base, st);
      all_done = (__syncthreads_count(st.done || tid >= HP) == 256);
      if (warp == 0) {
        // compact: keep Gaussians that contribute to at least one pixel of this half tile
        bool any = false;
        if (lane < nb) {
          const uint4 m = *reinterpret_cast<const uint4 *>(&sm.masks[b][lane * 4]);
          any = (m.x | m.y | m.z | m.w) != 0u;
        }
        const unsigned ball = __ballot_sync(0xffffffffu, any);
        const int cnt = __popc(ball);
        if (lane == 0) {
          sm.ccount[b] = cnt;
          if (cnt > 0) mbar_expect_tx(&sm.mbar[b], (unsigned)cnt * rowbytes);
        }
        __syncwarp();
        if (any) {
          const int pos = __popc(ball & ((1u << lane) - 1u));
          sm.clist[b][pos] = lane;
          bulk_g2s(&sm.fbuf[b][pos][0], colors + (size_t)sm.id[b][lane] * D + ch0, rowbytes,
                   &sm.mbar[b]);
        }
      }
    }
    if (i > 0) {
      // phase B for batch i-1 (its w, masks, clist, ccount were published before the last barrier)
      const int pb = b ^ 1;
      const int cnt = sm.ccount[pb];
      if (cnt > 0) {
        const unsigned ph = pb ? phase1 : phase0;
        mbar_wait(&sm.mbar[pb], ph);
        if (pb) phase1 ^= 1u; else phase0 ^= 1u;
        int myg = 0; bool act = false;
        if (lane < cnt) {
          myg = sm.clist[pb][lane];
          const unsigned mk = (sm.masks[pb][myg * 4 + (warp >> 1)] >> ((warp & 1) * 16)) & 0xffffu;
          act = mk != 0u;
        }
        unsigned todo = __ballot_sync(0xffffffffu, act);
        while (todo) {
          const int c = __ffs(todo) - 1;
          todo &= todo - 1;
          const int g = __shfl_sync(0xffffffffu, myg, c);
          const float4 wa = *reinterpret_cast<const float4 *>(&sm.wbuf[pb][g * HP + pg * 8]);
          const float4 wb = *reinterpret_cast<const float4 *>(&sm.wbuf[pb][g * HP + pg * 8 + 4]);
          const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const float4 f = *reinterpret_cast<const float4 *>(&sm.fbuf[pb][c][j * 64 + cg * 4]);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              acc[k][j * 4 + 0] = fmaf(w8[k], f.x, acc[k][j * 4 + 0]);
              acc[k][j * 4 + 1] = fmaf(w8[k], f.y, acc[k][j * 4 + 1]);
              acc[k][j * 4 + 2] = fmaf(w8[k], f.z, acc[k][j * 4 + 2]);
              acc[k][j * 4 + 3] = fmaf(w8[k], f.w, acc[k][j * 4 + 3]);
            }
          }
        }
      }
    }
    __syncthreads();
    if (!have_cur) break;
  }

  // epilogue
  if (tid < HP) {
    sm.Tfin[p] = st.T;
    if (inside && ch0 == 0) {
      const size_t pix = (size_t)pyi * W + pxi;
      alphas[pix] = 1.f - st.T;
      last_ids[pix] = st.last;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int pp = pg * 8 + k;
    int dx, dy;
    hp_pixel(pp, dx, dy);
    const int xx = x0 + dx, yy = y0 + dy;
    if (xx < W && yy < H) {
      const float T = sm.Tfin[pp];
      float *o = render + ((size_t)yy * W + xx) * D + ch0;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = j * 64 + cg * 4;
        if (c < nch) {
          float4 v = make_float4(acc[k][j * 4], acc[k][j * 4 + 1], acc[k][j * 4 + 2], acc[k][j * 4 + 3]);
          if (bg) {
            const float4 bgv = *reinterpret_cast<const float4 *>(bg + ch0 + c);
            v.x = fmaf(T, bgv.x, v.x); v.y = fmaf(T, bgv.y, v.y);
            v.z = fmaf(T, bgv.z, v.z); v.w = fmaf(T, bgv.w, v.w);
          }
          stg_cs4(reinterpret_cast<float4 *>(o + c), v);
        }
      }
    }
  }
}

template <int CDIM>
int launch_narrow(const float *geom, const float *colors, int D, const float *bg, int W, int H,
                  const int *offsets, const int *ids, float *render, float *alphas, int *last_ids,
                  cudaStream_t st) {
  const int tw = (W + GAGS_TILE - 1) / GAGS_TILE, th = (H + GAGS_TILE - 1) / GAGS_TILE;
  blend_fwd_narrow<CDIM><<<dim3(tw, th), dim3(16, 16), 0, st>>>(
      reinterpret_cast<const float4 *>(geom), colors, D, bg, W, H, tw, offsets, ids, render, alphas,
      last_ids);
  return (int)cudaGetLastError();
}

template <int NJ>
int launch_wide(const float *geom, const float *colors, int D, int ch0, const float *bg, int W,
                int H, const int *offsets, const int *ids, float *render, float *alphas,
                int *last_ids, cudaStream_t st) {
  const int tw = (W + GAGS_TILE - 1) / GAGS_TILE;
  const int hh = (H + HROWS - 1) / HROWS;
  const size_t smem = sizeof(WideSmem<NJ>);
  {   // per-device attribute: set on every launch (a process may drive several GPUs)
    cudaError_t e = cudaFuncSetAttribute(blend_fwd_wide<NJ>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  blend_fwd_wide<NJ><<<dim3(tw, hh), 256, smem, st>>>(reinterpret_cast<const float4 *>(geom), colors,
                                                     D, ch0, bg, W, H, tw, offsets, ids, render,
                                                     alphas, last_ids);
  return (int)cudaGetLastError();
}

}  // namespace

// defined in blend_fwd_tc.cu / api.cu
int gags_blend_fwd_tc(const float *geom, const float *colors, int32_t D, const float *background,
                      int32_t width, int32_t height, const int32_t *offsets,
                      const int32_t *flatten_ids, float *render, float *alphas, int32_t *last_ids,
                      unsigned char *wcache, int32_t *wmeta, int32_t *wlist, int32_t *wcount,
                      cudaStream_t st);
extern int g_gags_blend_impl;
extern "C" int gags_blend_last_ids_optional(int32_t D);

static int blend_fwd_impl(const float *geom, const float *colors, int32_t D, const float *background,
                          int32_t width, int32_t height, const int32_t *offsets,
                          const int32_t *flatten_ids, float *render, float *alphas,
                          int32_t *last_ids, unsigned char *wcache, int32_t *wmeta, int32_t *wlist,
                          int32_t *wcount, void *stream) {
  if (!geom || !colors || !offsets || !render || !alphas) return GAGS_EINVAL;
  if (D < 1 || width <= 0 || height <= 0) return GAGS_EINVAL;
  if (!gags_aligned16(geom)) return GAGS_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  // D = 32 also takes the tensor-core kernel (one 64-channel atom, half of it idle): measured
  // faster than the 16x16-tile SIMT kernel at 720p / 500 k Gaussians (BASELINE.json config 2)
  const bool tc = D >= 32 && D % 16 == 0 && g_gags_blend_impl != 1;
  // last_ids may be NULL only where the kernel has a path that does not track it
  if (!last_ids && !gags_blend_last_ids_optional(D)) return GAGS_EINVAL;
  if (wcache && !tc) return GAGS_EINVAL;
  if (D <= 32 && !tc) {
    if (D <= 4) return launch_narrow<4>(geom, colors, D, background, width, height, offsets, flatten_ids, render, alphas, last_ids, st);
    if (D <= 8) return launch_narrow<8>(geom, colors, D, background, width, height, offsets, flatten_ids, render, alphas, last_ids, st);
    if (D <= 16) return launch_narrow<16>(geom, colors, D, background, width, height, offsets, flatten_ids, render, alphas, last_ids, st);
    return launch_narrow<32>(geom, colors, D, background, width, height, offsets, flatten_ids, render, alphas, last_ids, st);
  }
  if (D % 4 != 0) return GAGS_EINVAL;
  if (!gags_aligned16(colors) || !gags_aligned16(render) || (background && !gags_aligned16(background)))
    return GAGS_EALIGN;
  if (tc)
    return gags_blend_fwd_tc(geom, colors, D, background, width, height, offsets, flatten_ids,
                             render, alphas, last_ids, wcache, wmeta, wlist, wcount, st);
  if (g_gags_blend_impl == 2) return GAGS_EINVAL;
  for (int ch0 = 0; ch0 < D; ch0 += 256) {
    const int nch = (D - ch0) < 256 ? (D - ch0) : 256;
    const int nj = (nch + 63) / 64;
    int rc;
    switch (nj) {
      case 1: rc = launch_wide<1>(geom, colors, D, ch0, background, width, height, offsets, flatten_ids, render, alphas, last_ids, st); break;
      case 2: rc = launch_wide<2>(geom, colors, D, ch0, background, width, height, offsets, flatten_ids, render, alphas, last_ids, st); break;
      case 3: rc = launch_wide<3>(geom, colors, D, ch0, background, width, height, offsets, flatten_ids, render, alphas, last_ids, st); break;
      default: rc = launch_wide<4>(geom, colors, D, ch0, background, width, height, offsets, flatten_ids, render, alphas, last_ids, st); break;
    }
    if (rc != 0) return rc;
  }
  return 0;
}

extern "C" int gags_blend_fwd(const float *geom, const float *colors, int32_t D,
                              const float *background, int32_t width, int32_t height,
                              const int32_t *offsets, const int32_t *flatten_ids, float *render,
                              float *alphas, int32_t *last_ids, void *stream) {
  return blend_fwd_impl(geom, colors, D, background, width, height, offsets, flatten_ids, render,
                        alphas, last_ids, nullptr, nullptr, nullptr, nullptr, stream);
}

extern int g_fwd_variant;
extern "C" int gags_blend_last_ids_optional(int32_t D) {
  return (D >= 32 && D % 16 == 0 && g_gags_blend_impl != 1 && g_fwd_variant == 3) ? 1 : 0;
}

extern "C" int gags_blend_cache_supported(int32_t D) {
  return (D > 32 && D % 16 == 0 && g_gags_blend_impl != 1) ? 1 : 0;
}

extern "C" int64_t gags_blend_cache_slots(int64_t n_isects, int32_t n_tiles) {
  // half tile (t, h) owns slots [2 b(t) + h span(t), + span(t)),  b(t) = (offsets[t] >> 5) + t,
  // span(t) = b(t+1) - b(t) >= ceil(len(t) / 32): a scan-free upper bound on its batch count
  return 2 * ((n_isects >> 5) + (int64_t)n_tiles) + 2;
}

extern "C" int gags_blend_fwd_cached(const float *geom, const float *colors, int32_t D,
                                     const float *background, int32_t width, int32_t height,
                                     const int32_t *offsets, const int32_t *flatten_ids,
                                     float *render, float *alphas, int32_t *last_ids,
                                     void *wcache, int32_t *wmeta, int32_t *wlist, int32_t *wcount,
                                     void *stream) {
  if (!wcache || !wmeta || !wlist || !wcount) return GAGS_EINVAL;
  if (!gags_aligned16(wcache)) return GAGS_EALIGN;
  return blend_fwd_impl(geom, colors, D, background, width, height, offsets, flatten_ids, render,
                        alphas, last_ids, reinterpret_cast<unsigned char *>(wcache), wmeta, wlist,
                        wcount, stream);
}
