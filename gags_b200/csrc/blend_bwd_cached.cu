// blend_bwd_cached.cu — K8a feature-only backward from CACHED forward weights (tcgen05).
// Replaces the v_colors half of gsplat rasterize_to_pixels_bwd<CDIM> (x ceil(D/32) chunk launches +
// the slice/cat autograd glue) reached from /root/reference/train.py:174 (loss.backward()) through
// /root/reference/gaussian_renderer/__init__.py:56-70; semantics = SURVEY.md Appendix A.6, feature-
// only case (frozen geometry, /root/reference/scene/gaussian_model.py:192-206):
//        v_colors[g, :] += sum_px w(g, px) * v_render[px, :],     w = the forward blend weight.
//
// The training forward (gags_blend_fwd_cached, blend_fwd_tc.cu) saved, for every batch of 32
// Gaussians it blended into a 16x8 half tile, the weight tile exactly as it sat in shared memory
// (128 pixel rows x [hi(32) | lo(32)] bf16, SWIZZLE_128B, 16 KB) plus the 32 Gaussian ids.  The
// backward therefore recomputes nothing: it is a streaming GEMM per half tile,
//        Dt[ch, g] = Vt[ch, 128 px] * Wt[128 px, g]
// with A = Vt resident (128 channels of the half tile of v_render split to bf16 hi/lo, MN-major
// SWIZZLE_128B) and B = one cached tile per step landed by a 1-D bulk async copy (UBLKCP).  Per
// k-step  Vhi x [Whi|Wlo]  and  Vlo x [Whi|Wlo]  (N = 64 each), column g + column 32+g =
// (Vhi+Vlo)(Whi+Wlo).  One CTA = half tile x 128 channels with a 64 KB A tile, so TWO CTAs share an
// SM and one's v_render staging (pure HBM latency) hides behind the other's MMAs and reductions.
// Accumulators are double-buffered in TMEM; the epilogue warps drain a step (two tcgen05.ld in flight
// -> add halves -> per-warp transpose through 4 KB of shared memory -> 16-byte vector reductions,
// red.global.add.v4.f32, 128 contiguous bytes per Gaussian row and warp; exactly-zero rows are
// skipped) while the next step's MMAs run.  (Shared-memory staged bulk async reductions were tried
// first and were limited by the staging they pin; both forms meet the same L2 ceiling.)
// With the fused L1 loss and no pixel mask the staged operand is the exact SIGN (one bf16 part, see
// SGN below): half the MMAs and staging stores, and the operand buffer is double-buffered.
//
// Warps: 0-3 v_render staging, 4-7 epilogue (the four TMEM lane quarters), 8 bulk-copy producer,
// 9 MMA issue.  Persistent, two CTAs per SM (one is 40 % slower, a second epilogue set 2x slower).
//
// Roofline: HBM — H*W*4D (v_render, once) + 16.1 KB per cached batch + N_contrib*4D*2 (reduction
// target; the reductions resolve in L2).
#include <cstdlib>
#include "blend_tc_common.cuh"

#ifdef GAGS_TC_TIMING
__device__ long long g_cb_dbg[8 * 4 * 16 * 8];     // [cta slot][role][step][event]
extern "C" int gags_debug_timeline_bwd(long long *host_dst, int n) {
  return (int)cudaMemcpyFromSymbol(host_dst, g_cb_dbg, sizeof(long long) * (size_t)n);
}
#define CB_STAMP(role, step, ev)                                                               \
  do {                                                                                         \
    if (dbg_slot >= 0 && lane == 0 && (step) < 16)                                             \
      g_cb_dbg[((dbg_slot * 4 + (role)) * 16 + (step)) * 8 + (ev)] = clock64();               \
  } while (0)
#else
#define CB_STAMP(role, step, ev) do { } while (0)
#endif

namespace {

constexpr int CB_THREADS = 320;   // warps 0-3 v_render staging, 4-7 epilogue, 8 bulk-copy producer,
                                  // 9 MMA issue

constexpr int CB_JQ = 8;          // job-id ring (roles are never more than 4 jobs apart, see below)

struct CbCtl {
  uint64_t wfull[2], wfree[2], accfull[2], accfree[2], vfull[2], vfree[2];
  uint64_t jq_full[CB_JQ];
  uint32_t tmem_base;
  int jobq[CB_JQ];
};

// One job = one 16x8 half tile x one block of 128 channels.  Persistent CTAs (two per SM) pull job
// ids from a global counter (dynamic: a CTA that is dispatched late — e.g. behind a higher-priority
// side-stream kernel — simply finds less work left); inside a CTA the staging warps already load
// the NEXT job's v_render block (64 KB, pure HBM latency) while the epilogue warps are still
// reducing the current job's last steps — the reductions, not the MMAs, are what a job spends most
// of its time on.  Job ids travel through a ring of CB_JQ slots filled by staging warp 0: fetching
// job k needs vfull(k-1), hence vfree(k-2), hence every MMA of job k-2 issued, hence (accumulator
// double buffering) the epilogue at job >= k-4: no role is more than 4 jobs behind the fetcher.
struct CbLayout {
  static constexpr int VPART = 32768;                  // one bf16 part (hi or lo): 128 px x 128 ch
  static constexpr int V_OFF = 0;
  static constexpr int W_OFF = 2 * VPART;              // 2 stages x one 16 KB weight tile
  static constexpr int STG_OFF = W_OFF + 32768;        // 4 epilogue warps x [32 g][32 ch] fp32
  static constexpr int CTL_OFF = STG_OFF + 16384;
  static constexpr int BYTES = CTL_OFF + (int)sizeof(CbCtl);   // no slack: base must be 1 KB aligned
  static constexpr int TCOLS = 128;                    // 2 accumulator buffers x [hi(32) | lo(32)]
};
static_assert(CbLayout::BYTES <= (233472 / 2 - 1024), "cached backward must fit twice per SM");

// what every role needs to know about a job; all roles enumerate the jobs identically
struct CbJob {
  int nbat, hbase, x0, y0, cfirst, cvalid;
};
__device__ __forceinline__ CbJob cb_job(long long j, int nblk, int tile_w, int ch0, int nch,
                                        const int *__restrict__ offsets,
                                        const int *__restrict__ wcount) {
  CbJob jb;
  const int per_row = tile_w * nblk;
  const int by = (int)(j / per_row), bx = (int)(j - (long long)by * per_row);
  const int tx = bx / nblk, cblk = bx - tx * nblk;
  jb.nbat = __ldg(wcount + by * tile_w + tx);
  const int tile = (by >> 1) * tile_w + tx;
  const int s = __ldg(offsets + tile), e = __ldg(offsets + tile + 1);
  const int cbase = (s >> 5) + tile;
  jb.hbase = 2 * cbase + (by & 1) * ((e >> 5) + tile + 1 - cbase);
  jb.x0 = tx * GAGS_TILE;
  jb.y0 = by * 8;
  jb.cfirst = ch0 + cblk * 128;
  jb.cvalid = min(128, nch - cblk * 128);
  return jb;
}

// Fused L1 loss (L1 = true): `v_render` is then the RENDER itself and the staging warps turn it into
// the loss gradient on the fly, scale * m * sign(render - emb[seg]), while summing m * |render -
// emb[seg]| into *loss — the semantics of l1_loss_segmap_kernel (train_ops.cu), without the 2 GB
// gradient map ever existing in HBM.  Every (pixel, channel) is staged by exactly one job, half
// tiles without Gaussians included (they only contribute to the loss).
// LM = 3: the reference's full target (read_sam_clip_feature, scene/dataset_readers.py:54-121):
// `seg` holds three levels [3][H*W], `scale3` the per-pixel level weights [3][H*W], the target row is
// sum_l scale3[l] * emb[seg[l]], a pixel is valid when all three ids are, and (optionally) the
// gradient w.r.t. the level weights is accumulated into v_scale [3][H*W] (see l1_loss_sam_kernel).
struct CbL1 {
  const int *seg;
  const float *emb;
  const float *mask;
  float *loss;
  int n_seg;
  float scale;
  const float *scale3;
  float *v_scale;
  long long hw;
};

// SGN (LM = 1 without a pixel mask): the L1 gradient is scale * sign(render - target) with ONE scale
// for the whole image, so the staged operand is the sign itself — exactly representable in bf16:
// no lo part to compute, store or multiply (half the staging stores, half the MMAs), and the scale
// is applied once to the accumulator in the epilogue.  More accurate than the hi / lo split of
// scale * sign (whose 2^-17 representation error is the same for every pixel and does not average
// out), hence the 1e-5 agreement the tests ask between the fused and the two-kernel route.
template <int LM, bool SGN = false>
__global__ void __launch_bounds__(CB_THREADS, 2)
blend_bwd_cached(int D, int ch0, int nch, int nblk, int W, int H, int tile_w, long long njobs,
                 const int *__restrict__ offsets, const unsigned char *__restrict__ wcache,
                 const int *__restrict__ wmeta, const int *__restrict__ wlist,
                 const int *__restrict__ wcount, int *__restrict__ jobctr,
                 const float *__restrict__ v_render, float *__restrict__ v_colors, CbL1 l1) {
  using L = CbLayout;
  constexpr bool L1 = LM != 0;
  constexpr int NL = LM == 3 ? 3 : 1;               // target levels
  // The kernel has no static shared memory, so the dynamic window starts at the CTA's shared-memory
  // base (1 KB aligned, which SWIZZLE_128B needs); the layout uses every byte of the two-CTAs-per-SM
  // budget, so this is checked instead of padded.
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  unsigned char *sm = smem_raw;
  unsigned char *sV = sm + L::V_OFF;
  unsigned char *sW = sm + L::W_OFF;
  CbCtl &ctl = *reinterpret_cast<CbCtl *>(sm + L::CTL_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef GAGS_TC_TIMING
  int dbg_slot = (blockIdx.x % 37 == 5 && blockIdx.x / 37 < 8) ? (int)(blockIdx.x / 37) : -1;
  if (warp == 0) CB_STAMP(3, 0, 0);
#endif

  if (tid == 0) {
    for (int k = 0; k < 2; ++k) {
      mbar_init(&ctl.wfull[k], 1);
      mbar_init(&ctl.wfree[k], 1);
      mbar_init(&ctl.accfull[k], 1);
      mbar_init(&ctl.accfree[k], 4);               // epilogue warps
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&ctl.vfull[k], 4);                 // staging warps
      mbar_init(&ctl.vfree[k], 1);                 // tcgen05.commit after a job's last MMA
    }
    for (int k = 0; k < CB_JQ; ++k) mbar_init(&ctl.jq_full[k], 1);
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<L::TCOLS>(&ctl.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = ctl.tmem_base;
  // k-th job id of this CTA as seen by a non-staging role (-1 = no more work)
  auto next_job = [&](int k) -> long long {
    mbar_wait_bounded(&ctl.jq_full[k & (CB_JQ - 1)], (uint32_t)((k / CB_JQ) & 1));
    return (long long)*reinterpret_cast<volatile int *>(&ctl.jobq[k & (CB_JQ - 1)]);
  };

  if (warp < 4) {
    // ======================= v_render staging ======================================================
    const int q = warp;                             // pixel block (8x4) of the half tile
    const int n0 = (lane & 15) * 8;                 // lane l owns channels [8 (l & 15), +8) of pixel
    const uint32_t coff = (uint32_t)((n0 >> 6) & 1) * 16384u + (uint32_t)((n0 & 63) >> 3) * 16u;
    // SGN: one bf16 part per job, so the hi | lo space holds TWO jobs' operands — job k + 1 is staged
    // while the MMAs of job k still read theirs (the staging warps never wait for the job in flight)
    constexpr int NVB = SGN ? 2 : 1;
    int jn = 0;                                     // non-empty jobs staged so far
    float l1acc = 0.f;
    for (int k = 0;; ++k) {
      // fetch + publish the k-th job id (staging leads every other role)
      if (tid == 0) {
        const int got = atomicAdd(jobctr, 1);
        ctl.jobq[k & (CB_JQ - 1)] = (long long)got < njobs ? got : -1;
        __threadfence_block();
        mbar_arrive(&ctl.jq_full[k & (CB_JQ - 1)]);
      }
      named_bar_sync(3, 128);
      const long long j = (long long)*reinterpret_cast<volatile int *>(&ctl.jobq[k & (CB_JQ - 1)]);
      if (j < 0) break;
      const CbJob jb = cb_job(j, nblk, tile_w, ch0, nch, offsets, wcount);
      const bool live = jb.nbat > 0;
      if (!L1 && !live) continue;
      const bool chan_ok = n0 < jb.cvalid;
      const float *vbase = v_render + jb.cfirst + n0;
      // pixels per round: the fused-loss form also holds the target rows, so it takes half as many
      constexpr int PJ = LM == 3 ? 2 : (L1 ? 4 : 8);
      int sg_cur[PJ][NL];
      auto load_seg = [&](int round) {
#pragma unroll
        for (int jj = 0; jj < PJ; ++jj) {
          const int ql = 2 * (round * PJ + jj) + (lane >> 4);
          const int xx = jb.x0 + ((q & 1) << 3) + (ql & 7), yy = jb.y0 + ((q >> 1) << 2) + (ql >> 3);
#pragma unroll
          for (int l = 0; l < NL; ++l) {
            sg_cur[jj][l] = -1;
            if (L1 && chan_ok && xx < W && yy < H)
              sg_cur[jj][l] = __ldg(l1.seg + (size_t)l * l1.hw + (size_t)yy * W + xx);
          }
        }
      };
      if constexpr (L1) load_seg(0);
      // first half of the loads may fly before the previous job's MMAs have released the buffer
#pragma unroll 1
      for (int round = 0; round < 16 / PJ; ++round) {
        float4 v[PJ][2];
        if constexpr (L1) {
          // the segment ids of this round were loaded one round ahead (sg_cur): the target row's
          // address depends on them, and a load chain seg -> emb per round showed up as +0.1 ms
          int sg[PJ][NL];
          float mk[PJ], wl[PJ][NL];
#pragma unroll
          for (int jj = 0; jj < PJ; ++jj) {
            const int ql = 2 * (round * PJ + jj) + (lane >> 4);
            const int xx = jb.x0 + ((q & 1) << 3) + (ql & 7), yy = jb.y0 + ((q >> 1) << 2) + (ql >> 3);
            mk[jj] = 0.f;
            const bool in = chan_ok && xx < W && yy < H;
            if (in) mk[jj] = l1.mask ? fabsf(__ldg(l1.mask + (size_t)yy * W + xx)) : 1.f;
#pragma unroll
            for (int l = 0; l < NL; ++l) {
              sg[jj][l] = sg_cur[jj][l];
              wl[jj][l] = 1.f;
              if (LM == 3 && in) wl[jj][l] = __ldg(l1.scale3 + (size_t)l * l1.hw + (size_t)yy * W + xx);
              if (sg[jj][l] < 0 || sg[jj][l] >= l1.n_seg) mk[jj] = 0.f;   // no target: weight 0
            }
          }
          float4 t[PJ][NL][2];
#pragma unroll
          for (int jj = 0; jj < PJ; ++jj) {
            const int ql = 2 * (round * PJ + jj) + (lane >> 4);
            const int xx = jb.x0 + ((q & 1) << 3) + (ql & 7), yy = jb.y0 + ((q >> 1) << 2) + (ql >> 3);
            v[jj][0] = v[jj][1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int l = 0; l < NL; ++l) t[jj][l][0] = t[jj][l][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (chan_ok && xx < W && yy < H) {
              const float4 *src = reinterpret_cast<const float4 *>(vbase + ((size_t)yy * W + xx) * D);
              v[jj][0] = ldg_nc4(src);
              v[jj][1] = ldg_nc4(src + 1);
              if (mk[jj] != 0.f) {
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                  const float4 *te = reinterpret_cast<const float4 *>(
                      l1.emb + (size_t)sg[jj][l] * D + jb.cfirst + n0);
                  t[jj][l][0] = __ldg(te);
                  t[jj][l][1] = __ldg(te + 1);
                }
              }
            }
          }
          if (round + 1 < 16 / PJ) load_seg(round + 1);
#pragma unroll
          for (int jj = 0; jj < PJ; ++jj) {
            const float sc = SGN ? (mk[jj] != 0.f ? 1.f : 0.f) : l1.scale * mk[jj];
            float part = 0.f;
            float vs[NL];
#pragma unroll
            for (int l = 0; l < NL; ++l) vs[l] = 0.f;
#define GAGS_L1C(h, c)                                                                         \
            {                                                                                  \
              float tg = t[jj][0][h].c;                                                        \
              if (LM == 3)   /* same order as l1_loss_sam_kernel: ((w0 e0) + w1 e1) + w2 e2 */    \
                tg = fmaf(wl[jj][NL - 1], t[jj][NL - 1][h].c,                                   \
                          fmaf(wl[jj][1 % NL], t[jj][1 % NL][h].c, wl[jj][0] * tg));            \
              const float d = v[jj][h].c - tg;                                                 \
              part += fabsf(d);                                                                \
              const float gq = d > 0.f ? sc : (d < 0.f ? -sc : 0.f);   /* SGN: sc is 0 or 1 */   \
              v[jj][h].c = gq;                                                                 \
              if (LM == 3) {                                                                   \
                _Pragma("unroll") for (int l = 0; l < NL; ++l)                                 \
                    vs[l] = fmaf(-gq, t[jj][l][h].c, vs[l]);                                   \
              }                                                                                \
            }
            GAGS_L1C(0, x) GAGS_L1C(0, y) GAGS_L1C(0, z) GAGS_L1C(0, w)
            GAGS_L1C(1, x) GAGS_L1C(1, y) GAGS_L1C(1, z) GAGS_L1C(1, w)
#undef GAGS_L1C
            l1acc = fmaf(mk[jj], part, l1acc);
            if (LM == 3 && l1.v_scale != nullptr) {
              // this lane's 8 channels -> the 16 lanes that share the pixel -> one atomic per level
              const int ql = 2 * (round * PJ + jj) + (lane >> 4);
              const int xx = jb.x0 + ((q & 1) << 3) + (ql & 7), yy = jb.y0 + ((q >> 1) << 2) + (ql >> 3);
#pragma unroll
              for (int l = 0; l < NL; ++l) {
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) vs[l] += __shfl_xor_sync(0xffffffffu, vs[l], o);
                if ((lane & 15) == 0 && mk[jj] != 0.f && xx < W && yy < H)
                  atomicAdd(l1.v_scale + (size_t)l * l1.hw + (size_t)yy * W + xx,
                            SGN ? vs[l] * l1.scale : vs[l]);     // SGN: gq above is the bare sign
              }
            }
          }
          if (!live) continue;                         // empty half tile: loss only
        } else {
#pragma unroll
          for (int jj = 0; jj < PJ; ++jj) {
            const int ql = 2 * (round * PJ + jj) + (lane >> 4);   // pixel inside this 8x4 block
            const int xx = jb.x0 + ((q & 1) << 3) + (ql & 7), yy = jb.y0 + ((q >> 1) << 2) + (ql >> 3);
            v[jj][0] = v[jj][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (chan_ok && xx < W && yy < H) {
              const float4 *src = reinterpret_cast<const float4 *>(vbase + ((size_t)yy * W + xx) * D);
              v[jj][0] = ldg_nc4(src);
              v[jj][1] = ldg_nc4(src + 1);
            }
          }
        }
        const int vb = SGN ? (jn & 1) : 0;
        unsigned char *vhi = sV + vb * L::VPART, *vlo = sV + L::VPART;
        if (round == 0 && jn >= NVB)
          mbar_wait_bounded(&ctl.vfree[vb], (uint32_t)(((jn / NVB) - 1) & 1));
#pragma unroll
        for (int jj = 0; jj < PJ; ++jj) {
          const int r = q * 32 + 2 * (round * PJ + jj) + (lane >> 4);   // row of the K = 128 px dim
          const uint32_t off = (uint32_t)(r >> 3) * 1024u +
                               sw128((uint32_t)(r & 7) * 128u + (coff & 127u)) + (coff & ~127u);
          uint4 h, l;
          if constexpr (SGN) {                           // -1 / 0 / +1: exact in bf16, no lo part
            const __nv_bfloat162 b0 = __floats2bfloat162_rn(v[jj][0].x, v[jj][0].y);
            const __nv_bfloat162 b1 = __floats2bfloat162_rn(v[jj][0].z, v[jj][0].w);
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[jj][1].x, v[jj][1].y);
            const __nv_bfloat162 b3 = __floats2bfloat162_rn(v[jj][1].z, v[jj][1].w);
            h.x = *reinterpret_cast<const uint32_t *>(&b0);
            h.y = *reinterpret_cast<const uint32_t *>(&b1);
            h.z = *reinterpret_cast<const uint32_t *>(&b2);
            h.w = *reinterpret_cast<const uint32_t *>(&b3);
            *reinterpret_cast<uint4 *>(vhi + off) = h;
          } else {
            split_pack2(v[jj][0].x, v[jj][0].y, h.x, l.x);
            split_pack2(v[jj][0].z, v[jj][0].w, h.y, l.y);
            split_pack2(v[jj][1].x, v[jj][1].y, h.z, l.z);
            split_pack2(v[jj][1].z, v[jj][1].w, h.w, l.w);
            *reinterpret_cast<uint4 *>(vhi + off) = h;
            *reinterpret_cast<uint4 *>(vlo + off) = l;
          }
        }
      }
      if (!live) continue;
      fence_async_smem();
      mbar_arrive_warp(&ctl.vfull[SGN ? (jn & 1) : 0]);
      if (warp == 0 && jn < 2) CB_STAMP(3, 0, 1 + jn);
      ++jn;
    }
    if constexpr (L1) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) l1acc += __shfl_xor_sync(0xffffffffu, l1acc, o);
      if (lane == 0) atomicAdd(l1.loss, l1acc);
    }
  } else if (warp < 8) {
    // ======================= epilogue ==============================================================
    // The four warps are the four TMEM lane quarters = 128 channels.  Per step: TMEM -> registers ->
    // per-warp transpose -> 16-byte vector reds.
    const int q = warp - 4;                         // == warp % 4
    float *stg = reinterpret_cast<float *>(sm + L::STG_OFF + q * 4096);
    int gs = 0;                                     // global step counter of this CTA
    for (int k = 0;; ++k) {
      const long long j = next_job(k);
      if (j < 0) break;
      const CbJob jb = cb_job(j, nblk, tile_w, ch0, nch, offsets, wcount);
      for (int gi = 0; gi < jb.nbat; ++gi, ++gs) {
        const int buf = gs & 1;
        // Gaussian ids of the tile's 32 rows (lane g holds row g's id): straight from the cache,
        // issued before the wait so the latency hides behind the MMAs
        const int slot = jb.hbase + __ldg(wlist + jb.hbase + gi);
        const int gid_l = __ldg(wmeta + (size_t)slot * TC_KB + lane);
        if (q == 0 && gs < 16) CB_STAMP(0, gs, 0);
        mbar_wait_bounded(&ctl.accfull[buf], (uint32_t)((gs >> 1) & 1));
        tc_fence_after();
        if (q == 0 && gs < 16) CB_STAMP(0, gs, 1);
        float acc[32];
        {
          uint32_t ra[32], rb[32];
          const uint32_t ta = tb + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 64);
          tmem_ld_32x32_nowait(ta, ra);               // both halves in flight, one wait
          tmem_ld_32x32_nowait(ta + 32, rb);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            acc[k] = __uint_as_float(ra[k]) + __uint_as_float(rb[k]);
            if (SGN) acc[k] *= l1.scale;
          }
        }
        tc_fence_before();
        mbar_arrive_warp(&ctl.accfree[buf]);
        if (q == 0 && gs < 16) CB_STAMP(0, gs, 2);
        // Reductions: per-warp transpose through 4 KB of shared memory (lane = channel ->
        // lane = 4 channels of one of 4 rows), then 16-byte vector reds, 128 contiguous bytes per
        // Gaussian row and warp, fire-and-forget.  What bounds them (tools/red_rate.cu and the
        // timelines in profiles/): a warp sustains only so many red INSTRUCTIONS in flight, so
        // scalar reds from registers ran at 0.8 B/cycle/warp and v4 reds at ~3.5; shared-memory
        // staged TMA bulk reductions were limited by the staging buffers they pin (2 x 8 KB in
        // flight per CTA).  Chip-wide the rate tops out near 3 TB/s for distinct rows (each first
        // touch of a line is a DRAM read-modify-write) and 5+ TB/s when rows repeat in L2.
        // A survivor that reached no pixel of the half tile has an exactly-zero row: skipped.
#pragma unroll
        for (int g = 0; g < 32; ++g) stg[g * 32 + lane] = acc[g];
        __syncwarp();
        const int cb = q * 32 + (lane & 7) * 4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int g = it * 4 + (lane >> 3);
          const float4 v = *reinterpret_cast<const float4 *>(stg + g * 32 + (lane & 7) * 4);
          const int gg = __shfl_sync(0xffffffffu, gid_l, g);
          if (gg >= 0 && cb < jb.cvalid && (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f))
            red_add4(v_colors + (size_t)gg * D + jb.cfirst + cb, v);
        }
        __syncwarp();
        if (q == 0 && gs < 16) CB_STAMP(0, gs, 3);
      }
    }
  } else if (warp == 8) {
    // ======================= bulk-copy producer: one cached weight tile per step ====================
    if (lane == 0) {
      int gs = 0;
      for (int k = 0;; ++k) {
        const long long j = next_job(k);
        if (j < 0) break;
        const CbJob jb = cb_job(j, nblk, tile_w, ch0, nch, offsets, wcount);
        for (int gi = 0; gi < jb.nbat; ++gi, ++gs) {
          const int st = gs & 1;
          if (gs >= 2) mbar_wait_bounded(&ctl.wfree[st], (uint32_t)(((gs >> 1) - 1) & 1));
          if (gs < 16) CB_STAMP(1, gs, 0);
          mbar_expect_tx(&ctl.wfull[st], 16384u);
          const int slot = jb.hbase + __ldg(wlist + jb.hbase + gi);
          bulk_g2s(sW + st * 16384, wcache + (size_t)slot * 16384, 16384u, &ctl.wfull[st]);
        }
      }
    }
    __syncwarp();
  } else {
    // ======================= MMA issuer ============================================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(64, true, true);
      const uint64_t v_desc0 = umma_desc_sw128(smem_u32(sV), 16384, 1024);
      const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW), 16, 1024);
      int gs = 0, jn = 0;
      for (int k = 0;; ++k) {
        const long long j = next_job(k);
        if (j < 0) break;
        const CbJob jb = cb_job(j, nblk, tile_w, ch0, nch, offsets, wcount);
        if (jb.nbat <= 0) continue;
        constexpr int NVB = SGN ? 2 : 1;
        const int vb = SGN ? (jn & 1) : 0;
        mbar_wait_bounded(&ctl.vfull[vb], (uint32_t)((jn / NVB) & 1));
        for (int gi = 0; gi < jb.nbat; ++gi, ++gs) {
          const int st = gs & 1, buf = gs & 1;
          if (gs < 16) CB_STAMP(2, gs, 0);
          mbar_wait_bounded(&ctl.wfull[st], (uint32_t)((gs >> 1) & 1));
          if (gs < 16) CB_STAMP(2, gs, 1);
          if (gs >= 2) mbar_wait_bounded(&ctl.accfree[buf], (uint32_t)(((gs >> 1) - 1) & 1));
          tc_fence_after();
          if (gs < 16) CB_STAMP(2, gs, 2);
          const uint32_t d = tb + (uint32_t)(buf * 64);
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ahi = v_desc0 + (uint64_t)((vb * L::VPART + ks * 2048) >> 4);
            const uint64_t alo = ahi + (uint64_t)(L::VPART >> 4);
            const uint64_t bw = w_desc0 + (uint64_t)((st * 16384 + ks * 2048) >> 4);
            umma_bf16_ss(d, ahi, bw, idesc, ks > 0 ? 1u : 0u);
            if (!SGN) umma_bf16_ss(d, alo, bw, idesc, 1u);
          }
          umma_commit(&ctl.wfree[st]);
          umma_commit(&ctl.accfull[buf]);
          if (gs < 16) CB_STAMP(2, gs, 3);
        }
        umma_commit(&ctl.vfree[vb]);                 // the v_render block may be overwritten
        ++jn;
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) CB_STAMP(3, 0, 3);
  if (warp == 9) tmem_dealloc<L::TCOLS>(tb);
}

template <int LM, bool SGN = false>
int launch_cb(int D, int ch0, int nch, int W, int H, const int *offsets, const unsigned char *wcache,
              const int *wmeta, const int *wlist, int *wcount, const float *v_render,
              float *v_colors, CbL1 l1, cudaStream_t st) {
  using L = CbLayout;
  const int tw = (W + GAGS_TILE - 1) / GAGS_TILE;
  const int hh = (H + 7) / 8;
  const int nblk = (nch + 127) / 128;
  {   // per-device attribute: set on every launch (a process may drive several GPUs)
    cudaError_t e = cudaFuncSetAttribute(blend_bwd_cached<LM, SGN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES);
    if (e != cudaSuccess) return (int)e;
  }
  const long long njobs = (long long)tw * nblk * hh;
  if (njobs > 0x7fffffffLL - 4096) return GAGS_ERANGE;
  static int ctas_per_sm = 0;                                    // tuning: GAGS_B200_BWD_CTAS=1|2
  if (ctas_per_sm == 0) {
    const char *e = getenv("GAGS_B200_BWD_CTAS");
    ctas_per_sm = (e && atoi(e) == 1) ? 1 : 2;
  }
  const long long want = (long long)ctas_per_sm * gags_sm_count();   // persistent CTAs (two per SM)
  const unsigned grid = (unsigned)(njobs < want ? njobs : want);
  int *jobctr = wcount + (size_t)tw * hh;            // the caller's extra int behind the counts
  cudaError_t e = cudaMemsetAsync(jobctr, 0, sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  blend_bwd_cached<LM, SGN><<<grid, CB_THREADS, L::BYTES, st>>>(
      D, ch0, nch, nblk, W, H, tw, njobs, offsets, wcache, wmeta, wlist, wcount, jobctr, v_render,
      v_colors, l1);
  return (int)cudaGetLastError();
}

// One warp per half tile: every Gaussian of a batch the forward blended (and cached) gets its
// row flag set.  The feature backward reduces into exactly these rows (a superset of the rows whose
// gradient is non-zero), so a flag of 0 promises an all-zero gradient row — what the row-sparse
// optimiser pass (gags_adam_step_rows) and the sparse multi-GPU exchange skip.
__global__ void __launch_bounds__(256)
mark_rows_kernel(int tile_w, int n_half, const int *__restrict__ offsets,
                 const int *__restrict__ wmeta, const int *__restrict__ wlist,
                 const int *__restrict__ wcount, unsigned char *__restrict__ flags) {
  const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= n_half) return;
  const CbJob jb = cb_job(w, 1, tile_w, 0, 128, offsets, wcount);
  for (int gi = 0; gi < jb.nbat; ++gi) {
    const int slot = jb.hbase + __ldg(wlist + jb.hbase + gi);
    const int gid = __ldg(wmeta + (size_t)slot * TC_KB + lane);
    if (gid >= 0) flags[gid] = 1;
  }
}

}  // namespace

// 1 (default) = the fused single-level L1 backward without a mask stages sign(render - target)
// exactly (SGN above); 0 = always the hi / lo split of scale * sign (A/B switch for tests)
int g_bwd_sign_operand = 1;
extern "C" int gags_set_bwd_sign_operand(int32_t on) {
  if (on != 0 && on != 1) return GAGS_EINVAL;
  g_bwd_sign_operand = on;
  return 0;
}

extern "C" int gags_blend_bwd_features_cached(int32_t D, int32_t width, int32_t height,
                                              const int32_t *offsets, const void *wcache,
                                              const int32_t *wmeta, const int32_t *wlist,
                                              int32_t *wcount, const float *v_render,
                                              float *v_colors, void *stream) {
  if (!offsets || !wcache || !wmeta || !wlist || !wcount || !v_render || !v_colors) return GAGS_EINVAL;
  if (D <= 32 || D % 16 != 0 || width <= 0 || height <= 0) return GAGS_EINVAL;
  if (!gags_aligned16(wcache) || !gags_aligned16(v_render) || !gags_aligned16(v_colors))
    return GAGS_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned char *wc = reinterpret_cast<const unsigned char *>(wcache);
  for (int ch0 = 0; ch0 < D; ch0 += 256) {
    const int nch = (D - ch0) < 256 ? (D - ch0) : 256;
    const int rc = launch_cb<0>(D, ch0, nch, width, height, offsets, wc, wmeta, wlist, wcount,
                                    v_render, v_colors, CbL1{}, st);
    if (rc != 0) return rc;
  }
  return 0;
}

// Fused L1 loss + feature backward (see CbL1): *loss_out += sum m |render - emb[seg]|; the caller
// zeroes it and scales by 1 / (H W D).  grad_scale is d loss / d render's magnitude (1 / (H W D)
// for the mean).  Replaces gags_l1_loss_segmap + gags_blend_bwd_features_cached.
extern "C" int gags_blend_bwd_features_cached_l1(int32_t D, int32_t width, int32_t height,
                                                 const int32_t *offsets, const void *wcache,
                                                 const int32_t *wmeta, const int32_t *wlist,
                                                 int32_t *wcount, const float *render,
                                                 const int32_t *seg, const float *emb,
                                                 const float *mask, int32_t n_seg, float grad_scale,
                                                 float *loss_out, float *v_colors, void *stream) {
  if (!offsets || !wcache || !wmeta || !wlist || !wcount || !render || !v_colors || !seg || !emb ||
      !loss_out)
    return GAGS_EINVAL;
  if (D <= 32 || D % 16 != 0 || width <= 0 || height <= 0 || n_seg < 1) return GAGS_EINVAL;
  if (!gags_aligned16(wcache) || !gags_aligned16(render) || !gags_aligned16(v_colors) ||
      !gags_aligned16(emb))
    return GAGS_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned char *wc = reinterpret_cast<const unsigned char *>(wcache);
  const CbL1 l1{seg, emb, mask, loss_out, n_seg, grad_scale, nullptr, nullptr, 0};
  for (int ch0 = 0; ch0 < D; ch0 += 256) {
    const int nch = (D - ch0) < 256 ? (D - ch0) : 256;
    const int rc = (mask == nullptr && g_bwd_sign_operand)
                       ? launch_cb<1, true>(D, ch0, nch, width, height, offsets, wc, wmeta, wlist,
                                            wcount, render, v_colors, l1, st)
                       : launch_cb<1, false>(D, ch0, nch, width, height, offsets, wc, wmeta, wlist,
                                             wcount, render, v_colors, l1, st);
    if (rc != 0) return rc;
  }
  return 0;
}

// The same with the reference's full three-level target (CbL1, LM = 3; semantics of
// gags_l1_loss_sam): seg3 [3][H*W], scale_map3 [3][H*W], optional v_scale_map [3][H*W] (zeroed by the
// caller).
extern "C" int gags_blend_bwd_features_cached_sam(int32_t D, int32_t width, int32_t height,
                                                  const int32_t *offsets, const void *wcache,
                                                  const int32_t *wmeta, const int32_t *wlist,
                                                  int32_t *wcount, const float *render,
                                                  const int32_t *seg3, const float *emb,
                                                  const float *scale_map3, int32_t n_seg,
                                                  float grad_scale, float *loss_out,
                                                  float *v_scale_map, float *v_colors,
                                                  void *stream) {
  if (!offsets || !wcache || !wmeta || !wlist || !wcount || !render || !v_colors || !seg3 || !emb ||
      !scale_map3 || !loss_out)
    return GAGS_EINVAL;
  if (D <= 32 || D % 16 != 0 || width <= 0 || height <= 0 || n_seg < 1) return GAGS_EINVAL;
  if (!gags_aligned16(wcache) || !gags_aligned16(render) || !gags_aligned16(v_colors) ||
      !gags_aligned16(emb))
    return GAGS_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned char *wc = reinterpret_cast<const unsigned char *>(wcache);
  const CbL1 l1{seg3, emb, nullptr, loss_out, n_seg, grad_scale, scale_map3, v_scale_map,
                (long long)width * height};
  for (int ch0 = 0; ch0 < D; ch0 += 256) {
    const int nch = (D - ch0) < 256 ? (D - ch0) : 256;
    // the three-level target never has a pixel mask: the sign operand applies (see SGN)
    const int rc = g_bwd_sign_operand
                       ? launch_cb<3, true>(D, ch0, nch, width, height, offsets, wc, wmeta, wlist,
                                            wcount, render, v_colors, l1, st)
                       : launch_cb<3, false>(D, ch0, nch, width, height, offsets, wc, wmeta, wlist,
                                             wcount, render, v_colors, l1, st);
    if (rc != 0) return rc;
  }
  return 0;
}

// row_flags[g] = 1 for every Gaussian of every batch cached by gags_blend_fwd_cached for this view
// (flags are only ever set here: the caller owns clearing them).
extern "C" int gags_blend_cache_mark_rows(int32_t width, int32_t height, const int32_t *offsets,
                                          const int32_t *wmeta, const int32_t *wlist,
                                          const int32_t *wcount, uint8_t *row_flags, void *stream) {
  if (!offsets || !wmeta || !wlist || !wcount || !row_flags || width <= 0 || height <= 0)
    return GAGS_EINVAL;
  const int tw = (width + GAGS_TILE - 1) / GAGS_TILE;
  const int hh = (height + 7) / 8;
  const int n_half = tw * hh;
  mark_rows_kernel<<<(n_half + 7) / 8, 256, 0, (cudaStream_t)stream>>>(tw, n_half, offsets, wmeta,
                                                                      wlist, wcount, row_flags);
  return (int)cudaGetLastError();
}
