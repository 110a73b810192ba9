// blend_bwd_tc.cu — K8a feature-only backward for wide features on the 5th-gen tensor cores.
// Replaces the v_colors half of gsplat rasterize_to_pixels_bwd<CDIM> (x ceil(D/32) chunk launches +
// the slice/cat autograd glue) reached from /root/reference/train.py:174 (loss.backward()) through
// /root/reference/gaussian_renderer/__init__.py:56-70; semantics = SURVEY.md Appendix A.6, feature-
// only case (frozen geometry, /root/reference/scene/gaussian_model.py:192-206):
//        v_colors[g, :] += sum_px w(g, px) * v_render[px, :],     w = the forward blend weight.
//
// Per 16x8 half tile this is the dense product (transposed so that the Gaussian count is the MMA N
// dimension, which has granularity 16 instead of 128):
//        Dt[ch, g] = Vt[ch, 128 px] * Wt[128 px, g]
// A = Vt is resident for the whole CTA: the half tile of v_render, split to bf16 hi/lo, MN-major
// SWIZZLE_128B.  B = Wt is produced 32 Gaussians at a time by the pixel warps as one 128-B row per
// pixel [hi(32) | lo(32)] — byte-identical to the forward kernel's A tile.  Per k-step two MMAs:
//        Vhi x [Whi | Wlo]  (N = 64, columns 0-31 and 32-63)   and   Vlo x Whi  (N = 32, columns 0-31)
// so that column g + column 32+g = hi*hi + lo*hi + hi*lo.  Accumulators are double-buffered in
// TMEM; four epilogue warps drain a batch (tcgen05.ld -> add halves -> per-warp transpose in smem ->
// 16-byte vector reductions red.global.add.v4.f32, 128 B contiguous per Gaussian row) while the
// next batch's MMAs run.
//
// Warps: 0-3 pixel (weights), 4-7 producers (scan + exact alpha >= 1/255 cull), 8-11 v_render
// staging then epilogue, 12 MMA issue.
//
// Roofline: HBM — H*W*4D (v_render, once) + N_contrib*4D*2 (reduction target; the reductions resolve
// in L2) + 12 B per list entry scanned.
#include "blend_tc_common.cuh"

namespace {

constexpr int KB = TC_KB;
constexpr int RING = TC_RING;
constexpr int BW_THREADS = 416;

struct BwCtl {
  uint64_t list[2], full[2], free_[2], accfull[2], accfree[2], vfull;
  uint32_t tmem_base;
  int gcount[2], gbase[2], skip[2];
  int done_warps, term;
  int wcnt[4];
  int bcnt[4], bskip[4];
  int bgid[4][KB];
  alignas(16) float4 rec0[2][KB];     // per-stage batch records read by the pixel threads
  alignas(16) float4 rec1[2][KB];
};

template <int MB>
struct BwLayout {
  static constexpr int VPART = MB * 32768;             // one bf16 part of the resident v_render tile
  static constexpr int V_OFF = 0;
  static constexpr int W_OFF = 2 * VPART;              // 2 stages x 16 KB
  static constexpr int STG_OFF = W_OFF + 32768;        // 4 epilogue warps x 4 KB
  static constexpr int RING_OFF = STG_OFF + 16384;
  static constexpr int CTL_OFF = RING_OFF + RING * 36;
  static constexpr int BYTES = CTL_OFF + (int)sizeof(BwCtl) + 1024;
  static constexpr int TCOLS = MB == 1 ? 128 : 256;    // 2 buffers x MB x 64 columns
};

template <int MB>
__global__ void __launch_bounds__(BW_THREADS, 1)
blend_bwd_tc(const float4 *__restrict__ geom, int D, int ch0, int nch, int W, int H, int tile_w,
             const int *__restrict__ offsets, const int *__restrict__ ids,
             const float *__restrict__ v_render, float *__restrict__ v_colors) {
  using L = BwLayout<MB>;
  const int tile = (blockIdx.y >> 1) * tile_w + blockIdx.x;
  const int s = offsets[tile], e = offsets[tile + 1];
  if (e <= s) return;

  extern __shared__ unsigned char smem_raw[];
  unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *sV = sm + L::V_OFF;
  unsigned char *sW = sm + L::W_OFF;
  float4 *rg0 = reinterpret_cast<float4 *>(sm + L::RING_OFF);
  float4 *rg1 = rg0 + RING;
  int *rgid = reinterpret_cast<int *>(rg1 + RING);
  BwCtl &ctl = *reinterpret_cast<BwCtl *>(sm + L::CTL_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int x0 = blockIdx.x * GAGS_TILE, y0 = blockIdx.y * 8;

  if (tid == 0) {
    for (int k = 0; k < 2; ++k) {
      mbar_init(&ctl.list[k], 4);
      mbar_init(&ctl.full[k], 4);
      mbar_init(&ctl.free_[k], 1);
      mbar_init(&ctl.accfull[k], 1);
      mbar_init(&ctl.accfree[k], 4);
      ctl.gcount[k] = 0; ctl.gbase[k] = 0; ctl.skip[k] = 0;
    }
    mbar_init(&ctl.vfull, 4);
    ctl.done_warps = 0;
    ctl.term = -1;
    mbar_fence_init();
  }
  if (warp == 12) tmem_alloc<L::TCOLS>(&ctl.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = ctl.tmem_base;

  if (warp < 4) {
    // ======================= pixel warps ============================================================
    const int pw = warp;
    const int dx = ((pw & 1) << 3) + (lane & 7), dy = ((pw >> 1) << 2) + (lane >> 3);
    const int pxi = x0 + dx, pyi = y0 + dy;
    const bool inside = (pxi < W) && (pyi < H);
    const float px = (float)pxi + 0.5f, py = (float)pyi + 0.5f;
    TcPixel ps;
    tc_pixel_init(ps, px, py, inside);
    bool counted = false;
    const uint32_t rowoff = (uint32_t)tid * 128u;
    for (int i = 0;; ++i) {
      const int st = i & 1;
      mbar_wait_bounded(&ctl.list[st], (i >> 1) & 1);
      const int nb = *reinterpret_cast<volatile int *>(&ctl.gcount[st]);
      if (nb == 0) break;
      if (i >= 2) mbar_wait_bounded(&ctl.free_[st], ((i >> 1) - 1) & 1);
      unsigned char *wrow = sW + st * 16384;
      const bool wdone = __all_sync(0xffffffffu, tc_pixel_done(ps));
      if (wdone) {
        if (lane == 0) atomicAdd(&ctl.skip[st], 1);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4 *>(wrow + sw128(rowoff + c * 16)) = z;
      } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint4 h, l;
          tc_weights8(&ctl.rec0[st][c * 8], &ctl.rec1[st][c * 8], ps, h, l);
          *reinterpret_cast<uint4 *>(wrow + sw128(rowoff + c * 16)) = h;
          *reinterpret_cast<uint4 *>(wrow + sw128(rowoff + (c + 4) * 16)) = l;
        }
      }
      fence_async_smem();
      mbar_arrive_warp(&ctl.full[st]);
      if (!counted && __all_sync(0xffffffffu, tc_pixel_done(ps))) {
        counted = true;
        if (lane == 0) atomicAdd(&ctl.done_warps, 1);
      }
    }
  } else if (warp < 8) {
    // ======================= producer warps: scan + cull + publish =================================
    const int p = tid - 128, pw = warp - 4;
    const float hx0 = (float)x0 + 0.5f, hy0 = (float)y0 + 0.5f;
    int scan = s, qtail = 0, qhead = 0;
    bool pending = false;
    int pend_idx = 0, pend_gid = 0;
    float4 pa0 = make_float4(0.f, 0.f, 0.f, 0.f), pa1 = pa0;
    int nxt_gid = (scan + p < e) ? __ldg(ids + scan + p) : -1;

    auto issue_scan = [&]() {
      pend_idx = scan + p;
      pend_gid = nxt_gid;
      if (pend_gid >= 0) {
        pa0 = __ldg(geom + pend_gid * 2);
        pa1 = __ldg(geom + pend_gid * 2 + 1);
      }
      scan += 128;
      nxt_gid = (scan + p < e) ? __ldg(ids + scan + p) : -1;
      pending = true;
    };
    auto finish_scan = [&]() {
      const unsigned mask = (pend_gid >= 0) ? tc_block_mask(pa0, pa1, hx0, hy0) : 0u;
      const bool keep = mask != 0u;
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) ctl.wcnt[pw] = __popc(bal);
      named_bar_sync(1, 128);
      int basec = qtail, total = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = ctl.wcnt[k];
        if (k < pw) basec += c;
        total += c;
      }
      if (keep) {
        const int slot = (basec + __popc(bal & ((1u << lane) - 1u))) & (RING - 1);
        rg0[slot] = pa0;
        rg1[slot] = make_float4(pa1.x, pa1.y, __int_as_float(pend_idx), __uint_as_float(mask));
        rgid[slot] = pend_gid;
      }
      qtail += total;
      pending = false;
      named_bar_sync(1, 128);
    };

    issue_scan();
    for (int i = 0;; ++i) {
      const int st = i & 1;
      const int dw = *reinterpret_cast<volatile int *>(&ctl.done_warps);
      const bool stop_all = named_bar_or(1, 128, dw == 4);
      while (!stop_all && (qtail - qhead) < KB && (pending || scan < e)) {
        if (!pending) issue_scan();
        finish_scan();
      }
      const int nb = stop_all ? 0 : min(KB, qtail - qhead);
      // next round's loads fly while we wait for the stage
      if (!pending && nb > 0 && (qtail - qhead - nb) < KB && scan < e) issue_scan();
      if (i >= 2) mbar_wait_bounded(&ctl.free_[st], ((i >> 1) - 1) & 1);
      if (p == 0) {
        ctl.gcount[st] = nb;
        ctl.gbase[st] = qhead & (RING - 1);
        ctl.bcnt[i & 3] = nb;
      }
      if (p < KB) {
        TcRec r = tc_null_rec();
        if (p < nb) {
          const int slot = (qhead + p) & (RING - 1);
          r = tc_make_rec(rg0[slot], rg1[slot]);
          ctl.bgid[i & 3][p] = rgid[slot];
        }
        ctl.rec0[st][p] = r.q0;
        ctl.rec1[st][p] = r.q1;
      }
      mbar_arrive_warp(&ctl.list[st]);
      if (nb == 0) break;
      qhead += nb;
    }
  } else if (warp < 12) {
    // ======================= v_render staging, then epilogue =======================================
    const int q = warp - 8;                         // TMEM lane quarter == warp % 4
    {
      const int n0 = lane * 8;
      const bool lane_on = n0 < MB * 128;
      const bool chan_ok = n0 < nch;
      const uint32_t coff = (uint32_t)(n0 >> 7) * 32768u + (uint32_t)((n0 >> 6) & 1) * 16384u +
                            (uint32_t)((n0 & 63) >> 3) * 16u;
      unsigned char *vhi = sV, *vlo = sV + L::VPART;
#pragma unroll 1
      for (int r0 = 0; r0 < 32; r0 += 8) {
        float4 v[8][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int ql = r0 + j;                    // pixel index inside this warp's 8x4 block
          const int xx = x0 + ((q & 1) << 3) + (ql & 7), yy = y0 + ((q >> 1) << 2) + (ql >> 3);
          v[j][0] = v[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (chan_ok && xx < W && yy < H) {
            const float4 *src = reinterpret_cast<const float4 *>(
                v_render + ((size_t)yy * W + xx) * D + ch0 + n0);
            v[j][0] = ldg_nc4(src);
            v[j][1] = ldg_nc4(src + 1);
          }
        }
        if (lane_on) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = q * 32 + r0 + j;          // row of the K = 128 px dimension
            uint4 h, l;
            split_pack2(v[j][0].x, v[j][0].y, h.x, l.x);
            split_pack2(v[j][0].z, v[j][0].w, h.y, l.y);
            split_pack2(v[j][1].x, v[j][1].y, h.z, l.z);
            split_pack2(v[j][1].z, v[j][1].w, h.w, l.w);
            const uint32_t off = (uint32_t)(r >> 3) * 1024u +
                                 sw128((uint32_t)(r & 7) * 128u + (coff & 127u)) + (coff & ~127u);
            *reinterpret_cast<uint4 *>(vhi + off) = h;
            *reinterpret_cast<uint4 *>(vlo + off) = l;
          }
        }
      }
      fence_async_smem();
      mbar_arrive_warp(&ctl.vfull);
    }
    float *stg = reinterpret_cast<float *>(sm + L::STG_OFF + q * 4096);
    for (int i = 0;; ++i) {
      const int buf = i & 1;
      mbar_wait_bounded(&ctl.accfull[buf], (i >> 1) & 1);
      if (*reinterpret_cast<volatile int *>(&ctl.term) == i) break;
      tc_fence_after();
      const int nb = *reinterpret_cast<volatile int *>(&ctl.bcnt[i & 3]);
      const bool skipped = *reinterpret_cast<volatile int *>(&ctl.bskip[i & 3]) != 0;
      float acc[MB][32];
      int gids[8];                                   // rows this lane reduces into (read before the
#pragma unroll                                       // buffer is handed back: bgid[] is recycled)
      for (int it = 0; it < 8; ++it) gids[it] = ctl.bgid[i & 3][it * 4 + (lane >> 3)];
      if (!skipped) {
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          uint32_t ra[32], rb[32];
          const uint32_t ta = tb + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * (MB * 64) + mb * 64);
          tmem_ld_32x32(ta, ra);
          tmem_ld_32x32(ta + 32, rb);
#pragma unroll
          for (int k = 0; k < 32; ++k) acc[mb][k] = __uint_as_float(ra[k]) + __uint_as_float(rb[k]);
        }
      }
      tc_fence_before();
      mbar_arrive_warp(&ctl.accfree[buf]);
      if (skipped) continue;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        // lane = channel (mb*128 + q*32 + lane); stage [g][ch] then read rows back
#pragma unroll
        for (int g = 0; g < 32; ++g) stg[g * 32 + lane] = acc[mb][g];
        __syncwarp();
        const int cbase = mb * 128 + q * 32 + (lane & 7) * 4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int g = it * 4 + (lane >> 3);
          const float4 v = *reinterpret_cast<const float4 *>(stg + g * 32 + (lane & 7) * 4);
          if (g < nb && cbase < nch && (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)) {
            red_add4(v_colors + (size_t)gids[it] * D + ch0 + cbase, v);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ======================= MMA issuer ============================================================
    if (lane == 0) {
      const uint32_t idesc64 = umma_idesc_bf16(64, true, true);
      const uint32_t idesc32 = umma_idesc_bf16(32, true, true);
      const uint64_t v_desc0 = umma_desc_sw128(smem_u32(sV), 16384, 1024);
      const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW), 16, 1024);
      int seen[2] = {0, 0};
      bool vready = false;
      for (int i = 0;; ++i) {
        const int st = i & 1, buf = i & 1;
        mbar_wait_bounded(&ctl.list[st], (i >> 1) & 1);
        const int nb = *reinterpret_cast<volatile int *>(&ctl.gcount[st]);
        if (i >= 2) mbar_wait_bounded(&ctl.accfree[buf], ((i >> 1) - 1) & 1);
        if (nb == 0) {
          ctl.term = i;
          __threadfence_block();
          mbar_arrive(&ctl.accfull[buf]);
          break;
        }
        mbar_wait_bounded(&ctl.full[st], (i >> 1) & 1);
        if (!vready) { mbar_wait_bounded(&ctl.vfull, 0); vready = true; }
        tc_fence_after();
        const int votes_now = *reinterpret_cast<volatile int *>(&ctl.skip[st]);
        const int votes = votes_now - seen[st];
        seen[st] = votes_now;
        const bool skip = votes >= 4;
        ctl.bskip[i & 3] = skip ? 1 : 0;
        __threadfence_block();
        if (!skip) {
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {
            const uint32_t d = tb + (uint32_t)(buf * (MB * 64) + mb * 64);
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t ahi = v_desc0 + (uint64_t)((mb * 32768 + ks * 2048) >> 4);
              const uint64_t alo = ahi + (uint64_t)(L::VPART >> 4);
              const uint64_t bw = w_desc0 + (uint64_t)((st * 16384 + ks * 2048) >> 4);
              umma_bf16_ss(d, ahi, bw, idesc64, ks > 0 ? 1u : 0u);
              umma_bf16_ss(d, alo, bw, idesc32, 1u);
            }
          }
        }
        umma_commit(&ctl.free_[st]);
        umma_commit(&ctl.accfull[buf]);
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc<L::TCOLS>(tb);
}

template <int MB>
int launch_bw(const float *geom, int D, int ch0, int nch, int W, int H, const int *offsets,
              const int *ids, const float *v_render, float *v_colors, cudaStream_t st) {
  using L = BwLayout<MB>;
  const int tw = (W + GAGS_TILE - 1) / GAGS_TILE;
  const int hh = (H + 7) / 8;
  {   // per-device attribute: set on every launch (a process may drive several GPUs)
    cudaError_t e = cudaFuncSetAttribute(blend_bwd_tc<MB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES);
    if (e != cudaSuccess) return (int)e;
  }
  blend_bwd_tc<MB><<<dim3(tw, hh), BW_THREADS, L::BYTES, st>>>(
      reinterpret_cast<const float4 *>(geom), D, ch0, nch, W, H, tw, offsets, ids, v_render,
      v_colors);
  return (int)cudaGetLastError();
}

}  // namespace

// Tensor-core feature backward: 32 < D, D % 16 == 0.  Channels are processed 256 per launch.
int gags_blend_bwd_features_tc(const float *geom, int32_t D, int32_t width, int32_t height,
                               const int32_t *offsets, const int32_t *flatten_ids,
                               const float *v_render, float *v_colors, cudaStream_t st) {
  for (int ch0 = 0; ch0 < D; ch0 += 256) {
    const int nch = (D - ch0) < 256 ? (D - ch0) : 256;
    const int rc = (nch <= 128)
                       ? launch_bw<1>(geom, D, ch0, nch, width, height, offsets, flatten_ids,
                                      v_render, v_colors, st)
                       : launch_bw<2>(geom, D, ch0, nch, width, height, offsets, flatten_ids,
                                      v_render, v_colors, st);
    if (rc != 0) return rc;
  }
  return 0;
}
