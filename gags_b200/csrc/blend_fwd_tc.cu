// blend_fwd_tc.cu — K7 alpha-blend forward for wide features on the 5th-gen tensor cores.
// Replaces gsplat rasterize_to_pixels_fwd<CDIM> + the channel_chunk=32 loop + torch.cat around it
// (reached from /root/reference/gaussian_renderer/__init__.py:56-70); semantics = SURVEY.md A.5.
//
// The D-wide blend of one 16x8 half tile is the dense product
//        render[128 px, D] = Wt[128 px, G] * F[G, D],      Wt[p, g] = alpha_g(p) * T_g(p)
// over the tile's depth-sorted Gaussians; it is issued transposed, acc[D, 128 px] = F^T * Wt^T
// (feature tile = MN-major A operand, 128 channels per instruction; weight tile = K-major B operand,
// N = 128 pixels), so that a TMEM lane is a channel and the epilogue needs no transpose.  fp32 parity (1e-4) is kept on bf16 tensor cores by
// splitting both operands x = hi + lo (bf16 each, |x - hi - lo| <= 2^-18 |x|) and issuing three
// products hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM (measured 4e-6 relative,
// tools/umma_probe.cu).
//
// One CTA = one half tile, 14 warps, warp-specialised, two smem stages of 32 Gaussians:
//   warps 4-7   scanner: walks the tile list 128 entries at a time, culls every Gaussian whose
//               alpha >= 1/255 ellipse misses the half tile (exact), keeps the survivors in a ring
//               and publishes one batch (32 blend records + Gaussian ids) per stage.
//   warps 8-11  converters: gather the batch's feature rows with coalesced 32-B loads, four rows
//               in flight per thread in a rolling register ring that runs across batch boundaries
//               (the next batch's rows are already loading while this one is split), split them to
//               bf16 hi/lo in registers and store them in the MN-major SWIZZLE_128B layout.
//   warps 0-3   one thread per pixel: evaluate alpha, transmittance product, write the weight row
//               [hi(32) | lo(32)] (128 B, K-major SWIZZLE_128B).
//   warp 13     (training) one thread stores each blended batch's weight tile to the cache with a
//               bulk async copy for the cached backward (blend_bwd_cached.cu).
//   warp 12     one thread issues tcgen05.mma (M=128 channels, N=128 pixels, K=16) x ceil(D/128)
//               channel blocks x 3 products x 2 k-steps per batch; tcgen05.commit frees the stage.
// Epilogue (warps 0-11): tcgen05.ld (lane = channel, column = pixel) -> + T*background (skipped
// for an all-zero background) -> 128-B warp-wide streaming stores of the channel-last raster.
//
// (The role list above is the round-1 layout, `gags_set_fwd_variant(2)`; the default V3 layout splits
// the pixel work into front warps 4-7 = scanner + alpha evaluation with lane = Gaussian and chain
// warps 0-3 = transmittance chain with lane = pixel, and its epilogue issues TMA tensor stores.)
//
// Kernels in this file:
//   blend_fwd_tc<NATOM, V3, 0>   the whole forward in one pass (inference; training without a lazily
//                                updated feature table), one CTA per half tile, 2 CTAs / SM
//   blend_fwd_tc<0, true, 1>     weights pass: geometry only, writes alphas + cached weight tiles
//   blend_fwd_tc<NATOM, true, 2> blend pass from the cache, one CTA per half tile (A/B form)
//   blend_fwd_pers<NATOM>        blend pass from the cache, ONE persistent CTA per SM (default)
//
// Roofline: HBM.  Algorithmic bytes per launch: N_contrib*4D (feature rows, once; re-reads are L2
// hits) + H*W*(4D+8) (render, alpha, last_ids) + 12 B per list entry scanned.
#include <cuda.h>
#include <cstring>
#include "blend_tc_common.cuh"

// Optional timeline instrumentation (tools/tc_timeline.py builds a second library with
// -DGAGS_TC_TIMING): selected CTAs stamp clock64() at role milestones into a global buffer.
#ifdef GAGS_TC_TIMING
__device__ long long g_tc_dbg[8 * 4 * 64 * 8];     // [cta slot][role][batch][event]
extern "C" int gags_debug_timeline(long long *host_dst, int n) {
  return (int)cudaMemcpyFromSymbol(host_dst, g_tc_dbg, sizeof(long long) * (size_t)n);
}
#define TC_STAMP(role, batch, ev)                                                              \
  do {                                                                                         \
    if (dbg_slot >= 0 && lane == 0 && (batch) < 64)                                            \
      g_tc_dbg[((dbg_slot * 4 + (role)) * 64 + (batch)) * 8 + (ev)] = clock64();              \
  } while (0)
#else
#define TC_STAMP(role, batch, ev) do { } while (0)
#endif

// 2 = one thread per pixel evaluates alpha and the chain (round-1 kernel); 3 = front / chain split.
int g_fwd_variant = 3;
extern int g_gags_blend_impl;   // api.cu: 0 auto, 1 SIMT only, 2 tensor cores required

namespace {

constexpr int KB = TC_KB;
constexpr int RING = TC_RING;
constexpr int TC_THREADS = 448;   // 13 role warps + the weight-tile store warp
// the weights pass (MODE 1) has no converters, no MMA issuer and no epilogue: 4 chain + 4 front warps
// + the store warp, three CTAs per SM
template <int MODE> struct TcCfg {
  static constexpr int THREADS = MODE == 1 ? 288 : TC_THREADS;
  static constexpr int MINB = MODE == 1 ? 4 : 2;
  static constexpr int NCW = MODE == 2 ? 8 : 4;      // converter warps (MODE 2: the chain warps too)
};

struct TcCtl {
  uint64_t list[2], full[2], free_[2], sdone[2], afull[2];
  uint32_t tmem_base;
  int term, bg_nonzero[8];
  int gcount[2];
  int skip[2];
  int done_warps, skip_from, any_mma;
  int wcnt[4];
  int gid[2][KB];
  // v3 roles (front = scanner + alpha evaluation, chain = transmittance chain)
  int lidx[2][4][KB];                 // list index of each batch entry (for last_ids), per front warp
  int nbw[2][4];                      // batch size as published by front warp w (== gcount)
  int wdone[4];                       // chain warp w: all 32 pixels finished
  alignas(16) float Tfin[128];
  alignas(16) float bgs[256];
  alignas(16) float4 rec0[2][KB];     // per-stage batch records read by the pixel threads
  alignas(16) float4 rec1[2][KB];
};

template <int NATOM>
struct TcLayout {
  static constexpr int BPART = NATOM * 4096;           // one bf16 part (hi or lo) of one stage
  static constexpr int A_OFF = 0;                      // 2 stages x 16 KB
  static constexpr int B_OFF = 32768;                  // [stage][part][BPART]
  static constexpr int RING_OFF = B_OFF + 4 * BPART;
  static constexpr int CTL_OFF = RING_OFF + RING * 36;
  static constexpr int BYTES = CTL_OFF + (int)sizeof(TcCtl) + 1024;   // + alignment slack
  static constexpr int MB = (NATOM + 1) / 2;           // 128-channel blocks (MMA M)
  static constexpr int TCOLS = MB * 128;               // accumulator: lane = channel, column = pixel
};
// two CTAs per SM: 228 KB of shared memory, 1 KB reserved per CTA
static_assert(TcLayout<4>::BYTES <= (233472 / 2 - 1024), "forward TC kernel must fit twice per SM");

// MODE 0: the whole forward in one pass.  The two-pass form splits it at the weight-tile cache:
// MODE 1 (weights pass): scan + cull + alpha + transmittance chain only — writes alphas / last_ids,
//         the cached weight tiles and the batch lists, touches no feature and no raster; what it
//         blends is therefore known (gags_blend_cache_mark_rows) BEFORE any feature row is read.
// MODE 2 (blend pass): render = cached weights x features — the front / chain warps are replaced
//         by one producer thread that lands each cached 16 KB tile in the A stage with a bulk copy.
template <int NATOM, bool V3, int MODE>
__global__ void __launch_bounds__(TcCfg<MODE>::THREADS, TcCfg<MODE>::MINB)
blend_fwd_tc(const float4 *__restrict__ geom, const float *__restrict__ colors, int D, int ch0,
             int nch, const float *__restrict__ bg, int W, int H, int tile_w,
             const int *__restrict__ offsets, const int *__restrict__ ids,
             float *__restrict__ render, float *__restrict__ alphas, int *__restrict__ last_ids,
             unsigned char *__restrict__ wcache, int *__restrict__ wmeta, int *__restrict__ wlist,
             int *__restrict__ wcount, const __grid_constant__ CUtensorMap tmap_render, int use_tma) {
  using L = TcLayout<NATOM>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *sA = sm + L::A_OFF;
  unsigned char *sB = sm + L::B_OFF;
  float4 *rg0 = reinterpret_cast<float4 *>(sm + L::RING_OFF);
  float4 *rg1 = rg0 + RING;
  int *rgid = reinterpret_cast<int *>(rg1 + RING);
  TcCtl &ctl = *reinterpret_cast<TcCtl *>(sm + L::CTL_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = (blockIdx.y >> 1) * tile_w + blockIdx.x;
  const int x0 = blockIdx.x * GAGS_TILE, y0 = blockIdx.y * 8;
  const int s = offsets[tile], e = offsets[tile + 1];
  // weight-tile cache (training only): batch i of this half tile lives in slot hbase + i, see
  // gags_blend_cache_slots() for the closed-form, scan-free slot bound
  const bool cache = MODE != 2 && wcache != nullptr;
  const int cbase = (s >> 5) + tile;
  const int hbase = 2 * cbase + (int)(blockIdx.y & 1) * ((e >> 5) + tile + 1 - cbase);
#ifdef GAGS_TC_TIMING
  // 8 sampled CTAs spread over the grid
  int dbg_slot = -1;
  {
    const int lin = blockIdx.y * gridDim.x + blockIdx.x, tot = gridDim.x * gridDim.y;
    for (int k = 0; k < 8; ++k) if (lin == (tot / 9) * (k + 1)) dbg_slot = k;
  }
  TC_STAMP(3, 0, 0);
#endif

  if (tid == 0) {
    for (int k = 0; k < 2; ++k) {
      mbar_init(&ctl.list[k], MODE == 2 ? 1 : 4);  // scanner warps (MODE 2: the tile producer)
      mbar_init(&ctl.full[k], MODE == 2 ? TcCfg<MODE>::NCW + 1 : 8);  // pixel + converter warps (MODE 2: producer + conv.)
      mbar_init(&ctl.free_[k], 1);                 // MMA commit: stage's B tile and records reusable
      mbar_init(&ctl.sdone[k], 1);                 // training: the tile store has read the A stage
      mbar_init(&ctl.afull[k], 4);                 // training: pixel warps -> store warp, A tile written
      ctl.gcount[k] = 0; ctl.skip[k] = 0;
    }
    for (int w = 0; w < 4; ++w) ctl.wdone[w] = 0;
    ctl.done_warps = 0; ctl.skip_from = 0; ctl.any_mma = 0; ctl.term = -1;
    mbar_fence_init();
  }
  // background: staged once; an all-zero background (the reference's default, black) is detected
  // here so that the epilogue skips its T * bg term and the shared-memory reads behind it
  if (tid < 256) {
    const float b = (bg != nullptr && tid < nch) ? __ldg(bg + ch0 + tid) : 0.f;
    ctl.bgs[tid] = b;
    const bool nzw = __any_sync(0xffffffffu, b != 0.f);
    if (lane == 0) ctl.bg_nonzero[warp] = nzw ? 1 : 0;
  }
  if constexpr (MODE != 1) {
    if (warp == 12) tmem_alloc<L::TCOLS>(&ctl.tmem_base);
  }
  // v3: the front warps start walking the tile list right away — the first scan round is two
  // dependent global loads (ids -> geometry) of pure latency, and nothing it touches (the survivor
  // ring, named barrier 1) depends on the barrier / TMEM set-up the other warps are doing
  TcScanner sc;
  if (V3 && MODE != 2 && warp >= 4 && warp < 8) {
    sc.init(geom, ids, s, e, (float)x0 + 0.5f, (float)y0 + 0.5f, rg0, rg1, rgid, ctl.wcnt,
            tid - 128);
    if (sc.scan < e) sc.issue();
    while (sc.queued() < KB && sc.more()) {
      if (!sc.pending) sc.issue();
      sc.finish();
    }
    if (!sc.pending && sc.queued() < 2 * KB && sc.scan < e) sc.issue();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = ctl.tmem_base;

  if (MODE == 2 && warp < 4) {
    // final transmittance of this half tile's pixels, for the background term of the epilogue
    const int dx = ((warp & 1) << 3) + (lane & 7), dy = ((warp >> 1) << 2) + (lane >> 3);
    const int pxi = x0 + dx, pyi = y0 + dy;
    ctl.Tfin[tid] = (pxi < W && pyi < H) ? 1.f - alphas[(size_t)pyi * W + pxi] : 0.f;
  }
  if (MODE == 2 && warp >= 4 && warp < 8) {
    // ======================= blend pass: cached-tile producer (warp 4) =============================
    if (warp == 4) {
      const int nbat = wcount[blockIdx.y * gridDim.x + blockIdx.x];
      // The batch metadata (list entry -> slot -> 32 Gaussian ids: two dependent global loads) is
      // fetched one batch AHEAD, before waiting for the stage: it is off the per-batch critical path.
      int wl = (lane < nbat) ? __ldg(wlist + hbase + lane) : 0;      // list entries of batches 0..31
      size_t slot_n = 0;
      int gid_n = -1;
      if (nbat > 0) {
        slot_n = (size_t)hbase + (size_t)__shfl_sync(0xffffffffu, wl, 0);
        gid_n = __ldg(wmeta + slot_n * KB + lane);
      }
      for (int i = 0; i <= nbat; ++i) {
        const int st = i & 1;
        const size_t slot = slot_n;
        const int gid = (i < nbat) ? gid_n : -1;
        if (i + 1 < nbat) {
          if (((i + 1) & 31) == 0) wl = (i + 1 + lane < nbat) ? __ldg(wlist + hbase + i + 1 + lane) : 0;
          slot_n = (size_t)hbase + (size_t)__shfl_sync(0xffffffffu, wl, (i + 1) & 31);
          gid_n = __ldg(wmeta + slot_n * KB + lane);
        }
        if (i >= 2) mbar_wait_bounded(&ctl.free_[st], ((i >> 1) - 1) & 1);
        ctl.gid[st][lane] = gid;
        const int nb = __popc(__ballot_sync(0xffffffffu, gid >= 0));
        if (lane == 0) ctl.gcount[st] = nb;
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&ctl.list[st]);
          if (i < nbat) {
            mbar_expect_tx(&ctl.full[st], 16384u);
            bulk_g2s(sA + st * 16384, wcache + slot * 16384, 16384u, &ctl.full[st]);
          }
        }
        __syncwarp();
      }
      // every MMA issued has completed once the last batch's commit has arrived
      if (nbat > 0) mbar_wait_bounded(&ctl.free_[(nbat - 1) & 1], ((nbat - 1) >> 1) & 1);
    }
  } else if (V3 && MODE != 2 && warp < 4) {
    // ======================= v3 chain warps: one thread per pixel ==================================
    // reads the 32 alphas the front warp of the same 8x4 block left in this pixel's A-tile row,
    // runs the transmittance chain and overwrites the row with the bf16 [hi | lo] weights
    const int pw = warp;
    const int dx = ((pw & 1) << 3) + (lane & 7), dy = ((pw >> 1) << 2) + (lane >> 3);
    const int pxi = x0 + dx, pyi = y0 + dy;
    const bool inside = (pxi < W) && (pyi < H);
    // last_ids == NULL selects the FAST chain (see tc3_chain8): T then accumulates sum w
    const bool want_last = last_ids != nullptr;
    TcChain ps;
    ps.P = inside ? 1.f : 0.f; ps.T = want_last ? 1.f : 0.f; ps.last = 0;
    bool counted = false;
    const uint32_t rowoff = (uint32_t)tid * 128u;
    int i = 0;
    for (;; ++i) {
      const int st = i & 1;
      if (warp == 0) TC_STAMP(0, i, 0);
      named_bar_sync(2 + st * 4 + pw, 64);          // front warp pw has left this batch's alphas
      if (warp == 0) TC_STAMP(0, i, 1);
      const int nb = *reinterpret_cast<volatile int *>(&ctl.nbw[st][pw]);
      if (nb == 0) {
        // batch i does not exist: tell the store warp (the front warps have already seen the store
        // of batch i-2 before publishing, so the store warp is never two phases behind `afull`)
        if (cache) {
          if (warp == 0 && lane == 0) {
            *reinterpret_cast<volatile int *>(&ctl.term) = i;
            __threadfence_block();
          }
          mbar_arrive_warp(&ctl.afull[st]);
        }
        break;
      }
      unsigned char *arow = sA + st * 16384;
      const bool wdone = __all_sync(0xffffffffu, ps.P <= GAGS_T_STOP);
      if (wdone) {
        if (lane == 0) atomicAdd(&ctl.skip[st], 1);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4 *>(arow + sw128(rowoff + c * 16)) = z;
      } else {
        int lastk = -1;
        // rolled on purpose (instruction-cache footprint, see tc3_front_alphas); round c reads the
        // alphas parked in chunks c and c + 4 and writes its hi / lo chunks back to the same two
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          unsigned char *ph = arow + sw128(rowoff + c * 16);
          unsigned char *pl = arow + sw128(rowoff + (c + 4) * 16);
          const float4 a0 = *reinterpret_cast<const float4 *>(ph);
          const float4 a1 = *reinterpret_cast<const float4 *>(pl);
          uint4 h, l;
          if (want_last) tc3_chain8<false>(a0, a1, c * 8, ps, lastk, h, l);
          else tc3_chain8<true>(a0, a1, c * 8, ps, lastk, h, l);
          *reinterpret_cast<uint4 *>(ph) = h;
          *reinterpret_cast<uint4 *>(pl) = l;
        }
        if (want_last && lastk >= 0) ps.last = ctl.lidx[st][pw][lastk];
      }
      if (warp == 0) TC_STAMP(0, i, 2);
      fence_async_smem();
      mbar_arrive_warp(&ctl.full[st]);
      if (cache) mbar_arrive_warp(&ctl.afull[st]);
      if (warp == 0) TC_STAMP(0, i, 3);
      if (!counted && __all_sync(0xffffffffu, ps.P <= GAGS_T_STOP)) {
        counted = true;
        if (lane == 0) {
          *reinterpret_cast<volatile int *>(&ctl.wdone[pw]) = 1;
          atomicMax(&ctl.skip_from, i + 1);
          __threadfence_block();
          atomicAdd(&ctl.done_warps, 1);
        }
      }
    }
    ctl.Tfin[tid] = want_last ? ps.T : 1.f - ps.T;
    if (inside && ch0 == 0) {
      const size_t pix = (size_t)pyi * W + pxi;
      alphas[pix] = want_last ? 1.f - ps.T : ps.T;
      if (want_last) last_ids[pix] = ps.last;
    }
    if (MODE != 1 && i > 0) mbar_wait_bounded(&ctl.free_[(i - 1) & 1], ((i - 1) >> 1) & 1);
  } else if (V3 && MODE != 2 && warp < 8) {
    // ======================= v3 front warps: scanner + alpha evaluation ============================
    // lane = Gaussian of the batch: the record stays in registers and the warp evaluates the 32
    // alphas of ITS 8x4 pixel block (front warp w <-> chain warp w), all independent
    const int fw = warp - 4;
    const float pxc = (float)(x0 + ((fw & 1) << 3)) + 0.5f;
    const float pyc = (float)(y0 + ((fw >> 1) << 2)) + 0.5f;
    for (int i = 0;; ++i) {
      const int st = i & 1;
      if (warp == 4) TC_STAMP(3, i + 1, 0);
      const int dw = *reinterpret_cast<volatile int *>(&ctl.done_warps);
      const bool stop_all = named_bar_or(1, 128, dw == 4);
      while (!stop_all && sc.queued() < KB && sc.more()) {
        if (!sc.pending) sc.issue();
        sc.finish();
      }
      if (warp == 4) TC_STAMP(3, i + 1, 1);
      const int nb = stop_all ? 0 : min(KB, sc.queued());
      if (!sc.pending && nb > 0 && (sc.queued() - nb) < KB && sc.scan < e) sc.issue();
      // stage reuse: the MMA of batch i-2 has read A / B / gid, and (training) its tile has left A
      if (i >= 2) {
        if (MODE != 1) mbar_wait_bounded(&ctl.free_[st], ((i >> 1) - 1) & 1);
        if (cache) mbar_wait_bounded(&ctl.sdone[st], ((i >> 1) - 1) & 1);
      }
      // this lane's Gaussian of the batch, straight from the survivor ring
      TcFrontRec fr;
      fr.mx = fr.my = fr.A = fr.B = fr.C = fr.op = 0.f;
      int gid = -1, lidx = 0;
      if (lane < nb) {
        const int slot = (sc.qhead + lane) & (RING - 1);
        const float4 g0 = rg0[slot], g1 = rg1[slot];
        const TcRec r = tc_make_rec(g0, g1);
        fr.mx = r.q0.x; fr.my = r.q0.y; fr.A = r.q0.z; fr.B = r.q0.w; fr.C = r.q1.x; fr.op = r.q1.y;
        lidx = __float_as_int(g1.z);
        gid = rgid[slot];
      }
      // chain warp w only synchronises with front warp w: each front warp leaves its own copy of
      // what its partner reads (nb and the list indices are uniform over the front group)
      ctl.lidx[st][fw][lane] = lidx;
      if (lane == 0) ctl.nbw[st][fw] = nb;
      if (fw == 0) {
        if (lane == 0) ctl.gcount[st] = nb;
        ctl.gid[st][lane] = gid;
        if (cache && nb > 0) wmeta[(size_t)(hbase + i) * KB + lane] = gid;
      }
      mbar_arrive_warp(&ctl.list[st]);
      if (warp == 4) TC_STAMP(3, i + 1, 2);
      if (nb > 0 && *reinterpret_cast<volatile int *>(&ctl.wdone[fw]) == 0)
        tc3_front_alphas(fr, pxc, pyc, sA + st * 16384 + fw * 4096, lane);
      named_bar_arrive(2 + st * 4 + fw, 64);        // hardware barrier: chain warp fw blocks, no polling
      if (nb == 0) break;
      sc.qhead += nb;
    }
  } else if (MODE != 2 && warp < 4) {
    // ======================= pixel warps: one thread per pixel =====================================
    const int pw = warp;
    const int dx = ((pw & 1) << 3) + (lane & 7), dy = ((pw >> 1) << 2) + (lane >> 3);
    const int pxi = x0 + dx, pyi = y0 + dy;
    const bool inside = (pxi < W) && (pyi < H);
    TcPixel ps;
    tc_pixel_init(ps, (float)pxi + 0.5f, (float)pyi + 0.5f, inside);
    bool counted = false;
    const uint32_t rowoff = (uint32_t)tid * 128u;
    int i = 0;
    for (;; ++i) {
      const int st = i & 1;
      if (warp == 0) TC_STAMP(0, i, 0);
      mbar_wait_bounded(&ctl.list[st], (i >> 1) & 1);
      if (warp == 0) TC_STAMP(0, i, 1);
      const int nb = *reinterpret_cast<volatile int *>(&ctl.gcount[st]);
      // (training) the weight tile of batch i-2 must have left this stage before it is overwritten;
      // only the writers of the A tile wait for that, not the scanner / converters
      const bool wait_store = cache && i >= 2;
      if (nb == 0) {
        // tell the store warp that batch i does not exist.  Its barrier (afull) is only ever
        // advanced by the pixel warps, and only after the store of batch i-2 has been seen
        // (sdone), so the store warp can never be two phases behind it.
        if (cache) {
          if (wait_store) mbar_wait_bounded(&ctl.sdone[st], ((i >> 1) - 1) & 1);
          if (warp == 0 && lane == 0) {
            *reinterpret_cast<volatile int *>(&ctl.term) = i;
            __threadfence_block();
          }
          mbar_arrive_warp(&ctl.afull[st]);
        }
        break;
      }
      unsigned char *arow = sA + st * 16384;
      const bool wdone = __all_sync(0xffffffffu, tc_pixel_done(ps));
      if (wait_store) mbar_wait_bounded(&ctl.sdone[st], ((i >> 1) - 1) & 1);
      if (wdone) {
        if (lane == 0) atomicAdd(&ctl.skip[st], 1);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4 *>(arow + sw128(rowoff + c * 16)) = z;
      } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint4 h, l;
          tc_weights8(&ctl.rec0[st][c * 8], &ctl.rec1[st][c * 8], ps, h, l);
          *reinterpret_cast<uint4 *>(arow + sw128(rowoff + c * 16)) = h;
          *reinterpret_cast<uint4 *>(arow + sw128(rowoff + (c + 4) * 16)) = l;
        }
      }
      if (warp == 0) TC_STAMP(0, i, 2);
      fence_async_smem();
      mbar_arrive_warp(&ctl.full[st]);
      if (cache) mbar_arrive_warp(&ctl.afull[st]);
      if (warp == 0) TC_STAMP(0, i, 3);
      if (!counted && __all_sync(0xffffffffu, tc_pixel_done(ps))) {
        counted = true;
        if (lane == 0) {
          // this warp votes "skip" for every batch after i
          atomicMax(&ctl.skip_from, i + 1);
          __threadfence_block();
          atomicAdd(&ctl.done_warps, 1);
        }
      }
    }
    ctl.Tfin[tid] = ps.T;
    if (inside && ch0 == 0) {
      const size_t pix = (size_t)pyi * W + pxi;
      alphas[pix] = 1.f - ps.T;
      last_ids[pix] = ps.last;
    }
    // every MMA issued has completed once the last batch's commit has arrived
    if (i > 0) mbar_wait_bounded(&ctl.free_[(i - 1) & 1], ((i - 1) >> 1) & 1);
  } else if (MODE != 2 && warp < 8) {
    // ======================= scanner warps =========================================================
    const int p = tid - 128;
    TcScanner sc;
    sc.init(geom, ids, s, e, (float)x0 + 0.5f, (float)y0 + 0.5f, rg0, rg1, rgid, ctl.wcnt, p);
    if (sc.scan < e) sc.issue();
    for (int i = 0;; ++i) {
      const int st = i & 1;
      if (warp == 4) TC_STAMP(3, i + 1, 0);
      // uniform stop decision: the pixel warps have all terminated
      const int dw = *reinterpret_cast<volatile int *>(&ctl.done_warps);
      const bool stop_all = named_bar_or(1, 128, dw == 4);
      while (!stop_all && sc.queued() < KB && sc.more()) {
        if (!sc.pending) sc.issue();
        sc.finish();
      }
      if (warp == 4) TC_STAMP(3, i + 1, 1);
      const int nb = stop_all ? 0 : min(KB, sc.queued());
      // the next round's loads fly while this thread waits for the stage
      if (!sc.pending && nb > 0 && (sc.queued() - nb) < KB && sc.scan < e) sc.issue();
      if (i >= 2) mbar_wait_bounded(&ctl.free_[st], ((i >> 1) - 1) & 1);
      if (p == 0) ctl.gcount[st] = nb;
      if (p < KB) {
        TcRec r = tc_null_rec();
        int gid = -1;
        if (p < nb) {
          const int slot = (sc.qhead + p) & (RING - 1);
          r = tc_make_rec(rg0[slot], rg1[slot]);
          gid = rgid[slot];
        }
        ctl.rec0[st][p] = r.q0;
        ctl.rec1[st][p] = r.q1;
        ctl.gid[st][p] = gid;
        if (cache && nb > 0) wmeta[(size_t)(hbase + i) * KB + p] = gid;
      }
      mbar_arrive_warp(&ctl.list[st]);
      if (warp == 4) TC_STAMP(3, i + 1, 2);
      if (nb == 0) break;
      sc.qhead += nb;
    }
  } else if (MODE != 1 && warp < 12) {
    // ======================= converter warps =======================================================
    // warp cw owns batch rows [ROWS cw, ROWS cw + ROWS), ROWS = 32 / NCW; lane owns channels
    // [8 lane, 8 lane + 8).  (MODE 2: the four chain warps convert as well, NCW = 8.)
    {
    constexpr int NCW = TcCfg<MODE>::NCW, ROWS = 32 / NCW, RND = ROWS / 4;
    const int cw = warp >= 8 ? warp - 8 : warp + 4;
    const int n0 = lane * 8;
    const bool chan_ok = n0 < nch;
    const uint32_t coff = (uint32_t)(n0 >> 6) * 4096u + (uint32_t)((n0 & 63) >> 3) * 16u;
    const float *cbase = colors + ch0 + n0;
    float4 v[4][2];

    auto header = [&](int i, int &nb, bool &skipb) {
      const int st = i & 1;
      mbar_wait_bounded(&ctl.list[st], (i >> 1) & 1);
      nb = *reinterpret_cast<volatile int *>(&ctl.gcount[st]);
      const int dw = *reinterpret_cast<volatile int *>(&ctl.done_warps);
      const int sf = *reinterpret_cast<volatile int *>(&ctl.skip_from);
      skipb = (dw == 4) && (i >= sf);               // all four pixel warps vote skip for batch i
    };
    auto load_row = [&](int i, int nb, bool skipb, int row, float4 (&dst)[2]) {
      dst[0] = dst[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < nb && chan_ok && !skipb) {
        const int gid = ctl.gid[i & 1][row];
        const float4 *src = reinterpret_cast<const float4 *>(cbase + (size_t)gid * D);
        dst[0] = __ldg(src);
        dst[1] = __ldg(src + 1);
      }
    };
    auto store_row = [&](unsigned char *bhi, unsigned char *blo, int row, const float4 (&src)[2]) {
      uint4 h, l;
      split_pack2(src[0].x, src[0].y, h.x, l.x);
      split_pack2(src[0].z, src[0].w, h.y, l.y);
      split_pack2(src[1].x, src[1].y, h.z, l.z);
      split_pack2(src[1].z, src[1].w, h.w, l.w);
      const uint32_t off = (uint32_t)(row >> 3) * 1024u +
                           sw128((uint32_t)(row & 7) * 128u + (coff & 127u)) + (coff & ~127u);
      *reinterpret_cast<uint4 *>(bhi + off) = h;
      *reinterpret_cast<uint4 *>(blo + off) = l;
    };

    int nb;
    bool skipb;
    header(0, nb, skipb);
    if (nb > 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) load_row(0, nb, skipb, cw * ROWS + j, v[j]);
    }
    for (int i = 0;; ++i) {
      if (nb == 0) break;
      const int st = i & 1;
      if (warp == 8) TC_STAMP(1, i, 0);
      unsigned char *bhi = sB + (st * 2 + 0) * L::BPART;
      unsigned char *blo = sB + (st * 2 + 1) * L::BPART;
      const int nbr = (nb + 15) & ~15;               // rows the MMAs will read
      const bool do_store = chan_ok && !skipb;
      if (i >= 2) mbar_wait_bounded(&ctl.free_[st], ((i >> 1) - 1) & 1);
      if (warp == 8) TC_STAMP(1, i, 2);
      int nb2 = -1;
      bool skip2 = false;
      bool ahead = false;
      // ONE copy of the 4-row store + refill body, run for rows [0,4) and [4,8) of this warp
      // (rolled: instruction-cache footprint).  The refill of the second round already belongs to
      // batch i+1 and is issued only if that batch's list is out: never block on it here (its
      // publication may itself be waiting for this batch to be consumed).
#pragma unroll 1
      for (int h = 0; h < RND; ++h) {
        int li = i, ln = nb, lrow = cw * ROWS + 4 * (h + 1);
        bool ls = skipb;
        if (h == RND - 1) {
          ahead = __all_sync(0xffffffffu,
                             mbar_test_wait(&ctl.list[(i + 1) & 1], ((i + 1) >> 1) & 1));
          if (ahead) header(i + 1, nb2, skip2);
          li = i + 1; ln = nb2; lrow = cw * ROWS; ls = skip2;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int row = cw * ROWS + 4 * h + j;
          if (do_store && row < nbr) store_row(bhi, blo, row, v[j]);
          load_row(li, ln, ls, lrow + j, v[j]);
        }
        if (warp == 8 && h == 0) TC_STAMP(1, i, 3);
      }
      fence_async_smem();
      mbar_arrive_warp(&ctl.full[st]);
      if (warp == 8) TC_STAMP(1, i, 4);
      if (!ahead) {
        header(i + 1, nb2, skip2);
#pragma unroll
        for (int j = 0; j < 4; ++j) load_row(i + 1, nb2, skip2, cw * ROWS + j, v[j]);
      }
      nb = nb2;
      skipb = skip2;
    }
    }
  } else if (MODE != 1 && warp == 12) {
    // ======================= MMA issuer ============================================================
    if (lane == 0) {
      // transposed product: D[channel, pixel] += F^T[channel, g] * W^T[g, pixel].  The feature tile
      // is the MN-major A operand (one 128-channel block per instruction), the weight tile the
      // K-major B operand (N = 128 pixels), so that a TMEM lane is a channel and the epilogue's
      // warp-wide stores are contiguous in the channel-last output without a transpose.
      const uint32_t idesc = umma_idesc_bf16(128, true, false);
      const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sA), 16, 1024);
      const uint64_t f_desc0 = umma_desc_sw128(smem_u32(sB), 4096, 1024);
      uint32_t acc = 0;
      int seen[2] = {0, 0};
      for (int i = 0;; ++i) {
        const int st = i & 1;
        mbar_wait_bounded(&ctl.list[st], (i >> 1) & 1);
        const int nb = *reinterpret_cast<volatile int *>(&ctl.gcount[st]);
        if (nb == 0) break;
        TC_STAMP(2, i, 0);
        mbar_wait_bounded(&ctl.full[st], (i >> 1) & 1);
        TC_STAMP(2, i, 1);
        tc_fence_after();
        const int votes_now = *reinterpret_cast<volatile int *>(&ctl.skip[st]);
        const int votes = votes_now - seen[st];
        seen[st] = votes_now;
        if (MODE != 1 && votes < 4) {
          const int nk = (nb + 15) >> 4;
#pragma unroll 1
          for (int ks = 0; ks < nk; ++ks) {
            // only the 14-bit start-address field (16-B units) changes between descriptors
            const uint64_t whi = w_desc0 + (uint64_t)((st * 16384 + ks * 32) >> 4);
            const uint64_t wlo = whi + (uint64_t)(64 >> 4);
#pragma unroll
            for (int mb = 0; mb < L::MB; ++mb) {
              const uint64_t fhi =
                  f_desc0 + (uint64_t)(((st * 2) * L::BPART + mb * 8192 + ks * 2048) >> 4);
              const uint64_t flo = fhi + (uint64_t)(L::BPART >> 4);
              const uint32_t d = tb + (uint32_t)(mb * 128);
              umma_bf16_ss(d, fhi, whi, idesc, acc);
              umma_bf16_ss(d, flo, whi, idesc, 1);
              umma_bf16_ss(d, fhi, wlo, idesc, 1);
            }
            acc = 1;
          }
        }
        umma_commit(&ctl.free_[st]);
        TC_STAMP(2, i, 2);
      }
      ctl.any_mma = (int)acc;
    }
    __syncwarp();
  } else {
    // ======================= weight-tile store warp (training forward only) =========================
    // One bulk async copy (TMA, shared -> global) per blended batch: the 16 KB tile leaves exactly
    // as the MMA reads it.  The stage is handed back (second arrival on free_) once the copy has
    // finished reading shared memory.
    if (cache && lane == 0) {
      int seen[2] = {0, 0};
      int stored = 0;
      for (int i = 0;; ++i) {
        const int st = i & 1;
        mbar_wait_bounded(&ctl.afull[st], (i >> 1) & 1);
        if (*reinterpret_cast<volatile int *>(&ctl.term) == i) break;
        const int votes_now = *reinterpret_cast<volatile int *>(&ctl.skip[st]);
        const int votes = votes_now - seen[st];
        seen[st] = votes_now;
        if (votes < 4) {
          wlist[hbase + stored++] = i;                // batches the backward has to visit
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                           wcache + (size_t)(hbase + i) * 16384),
                       "r"(smem_u32(sA + st * 16384)), "r"(16384u)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        mbar_arrive(&ctl.sdone[st]);
      }
      wcount[blockIdx.y * gridDim.x + blockIdx.x] = stored;
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  }

  // ========================= epilogue: TMEM -> registers -> smem transpose -> HBM ==================
  if (warp == 0) TC_STAMP(3, 0, 1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) TC_STAMP(3, 0, 2);
  if (MODE != 1 && warp < 12) {
    const bool any = ctl.any_mma != 0;
    bool use_bg = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) use_bg = use_bg || (ctl.bg_nonzero[k] != 0);
    const int q = warp & 3, third = warp >> 2;
    // one item = one 128-channel block x this warp's 32 channels x one 8x4 pixel block: a single
    // tcgen05.ld (32 lanes x 32 columns), 32 conflict-free shared-memory stores into the warp's
    // [4 rows][8 px][32 ch] staging box and ONE tensor store (UTMASTG; the hardware clips the ragged
    // right / bottom edges and channels past D).  The main loop's A / B stages are dead by now and
    // hold the boxes; with 256 channels there is room for two per warp, so the store of item k
    // reads shared memory while item k+1 is staged.  (Per-lane streaming stores — 32 per item —
    // kept the LSU queue full for ~10 k cycles per CTA; so did 8-column items, by sheer count.)
    constexpr int NBUF = (L::B_OFF + 4 * L::BPART >= 12 * 8192) ? 2 : 1;
    unsigned char *boxes = sm + warp * (4096 * NBUF);
    int nbox = 0;
#pragma unroll 1
    for (int idx = third; idx < L::MB * 4; idx += 3) {
      const int mb = idx >> 2, pc = idx & 3;
      const int ch = mb * 128 + q * 32 + lane;
      if (mb * 128 + q * 32 >= nch) continue;                      // warp-uniform
      const int xb = x0 + ((pc & 1) << 3), yb = y0 + ((pc >> 1) << 2);
      if (yb >= H || xb >= W) continue;                            // warp-uniform
      uint32_t r[32];
      if (any) {
        tmem_ld_32x32(tb + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * 128 + pc * 32), r);
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) r[k] = 0u;
      }
      if (use_bg) {
        const float b = ch < nch ? ctl.bgs[ch] : 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = __float_as_uint(fmaf(ctl.Tfin[pc * 32 + j], b, __uint_as_float(r[j])));
      }
      if (use_tma) {
        unsigned char *box = boxes + (NBUF == 2 ? (nbox & 1) * 4096 : 0);
        if (lane == 0) bulk_wait_group_read<NBUF - 1>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          *reinterpret_cast<uint32_t *>(box + j * 128 + lane * 4) = r[j];
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmap_render, box, ch0 + mb * 128 + q * 32, xb, yb);
          bulk_commit_group();
        }
        ++nbox;
        continue;
      }
      if (ch >= nch) continue;
      float *dst = render + ((size_t)yb * W + xb) * D + ch0 + ch;
      const size_t rowstride = (size_t)W * D;
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        if (yb + y < H) {
#pragma unroll
          for (int x = 0; x < 8; ++x)
            if (xb + x < W) stg_cs1(dst + (unsigned)(x * D), __uint_as_float(r[y * 8 + x]));
        }
        dst += rowstride;
      }
    }
    if (use_tma && lane == 0) bulk_wait_group<0>();
  }
  if (warp == 0) TC_STAMP(3, 0, 3);
  tc_fence_before();
  __syncthreads();
  if constexpr (MODE != 1) {
    if (warp == 12) tmem_dealloc<L::TCOLS>(tb);
  }
}

// epilogue: 1 = TMA tensor stores from a staged box (UTMASTG), 0 = per-lane streaming stores
int g_fwd_tma_epilogue = 1;

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// tensor map of the channel-last raster [H][W][D] fp32 with a (32 channels x 8 pixels x 4 rows) box
bool make_render_map(CUtensorMap *m, float *render, int D, int W, int H) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || (D & 3) || !gags_aligned16(render)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)W, (cuuint64_t)H};
  const cuuint64_t strides[2] = {(cuuint64_t)D * 4, (cuuint64_t)W * D * 4};
  const cuuint32_t box[3] = {32, 8, 4};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, render, dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the same raster with a (256 channels x 8 pixels x 4 rows) box: a pixel's whole 1 KB row leaves in
// one piece (persistent blend pass, D % 256 == 0 channel blocks)
bool make_render_map_wide(CUtensorMap *m, float *render, int D, int W, int H) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || (D & 3) || D < 256 || !gags_aligned16(render)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)W, (cuuint64_t)H};
  const cuuint64_t strides[2] = {(cuuint64_t)D * 4, (cuuint64_t)W * D * 4};
  const cuuint32_t box[3] = {256, 8, 4};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, render, dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NATOM, bool V3, int MODE = 0>
int launch_tc(const float *geom, const float *colors, int D, int ch0, int nch, const float *bg, int W,
              int H, const int *offsets, const int *ids, float *render, float *alphas,
              int *last_ids, unsigned char *wcache, int *wmeta, int *wlist, int *wcount,
              cudaStream_t st) {
  using L = TcLayout<NATOM>;
  const int tw = (W + GAGS_TILE - 1) / GAGS_TILE;
  const int hh = (H + 7) / 8;
  {   // per-device attribute: set on every launch (a process may drive several GPUs)
    cudaError_t e = cudaFuncSetAttribute(blend_fwd_tc<NATOM, V3, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES);
    if (e != cudaSuccess) return (int)e;
  }
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  const int use_tma =
      (MODE != 1 && g_fwd_tma_epilogue && make_render_map(&tmap, render, D, W, H)) ? 1 : 0;
  blend_fwd_tc<NATOM, V3, MODE><<<dim3(tw, hh), TcCfg<MODE>::THREADS, L::BYTES, st>>>(
      reinterpret_cast<const float4 *>(geom), colors, D, ch0, nch, bg, W, H, tw, offsets, ids,
      render, alphas, last_ids, wcache, wmeta, wlist, wcount, tmap, use_tma);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int gags_set_fwd_variant(int32_t v) {
  // 2 / 3 = role layout (see the header); + 10 = per-lane streaming stores instead of the TMA
  // tensor-store epilogue
  if (v != 2 && v != 3 && v != 12 && v != 13) return GAGS_EINVAL;
  g_fwd_variant = v % 10;
  g_fwd_tma_epilogue = v < 10 ? 1 : 0;
  return 0;
}

// Tensor-core wide forward: 32 < D, D % 16 == 0.  Channels are processed 256 per launch.  When
// `wcache` is non-NULL the first launch also saves every batch's weight tile (+ Gaussian ids) for
// gags_blend_bwd_features_cached (blend_bwd_cached.cu).
int gags_blend_fwd_tc(const float *geom, const float *colors, int32_t D, const float *background,
                      int32_t width, int32_t height, const int32_t *offsets,
                      const int32_t *flatten_ids, float *render, float *alphas, int32_t *last_ids,
                      unsigned char *wcache, int32_t *wmeta, int32_t *wlist, int32_t *wcount,
                      cudaStream_t st) {
  for (int ch0 = 0; ch0 < D; ch0 += 256) {
    const int nch = (D - ch0) < 256 ? (D - ch0) : 256;
    const int natom = (nch + 63) / 64;
    unsigned char *wc = ch0 == 0 ? wcache : nullptr;
    int rc;
#define GAGS_TC_ARGS geom, colors, D, ch0, nch, background, width, height, offsets, flatten_ids, \
                     render, alphas, last_ids, wc, wmeta, wlist, wcount, st
    if (g_fwd_variant == 2) {
      switch (natom) {
        case 1: rc = launch_tc<1, false>(GAGS_TC_ARGS); break;
        case 2: rc = launch_tc<2, false>(GAGS_TC_ARGS); break;
        case 3: rc = launch_tc<3, false>(GAGS_TC_ARGS); break;
        default: rc = launch_tc<4, false>(GAGS_TC_ARGS); break;
      }
    } else {
      switch (natom) {
        case 1: rc = launch_tc<1, true>(GAGS_TC_ARGS); break;
        case 2: rc = launch_tc<2, true>(GAGS_TC_ARGS); break;
        case 3: rc = launch_tc<3, true>(GAGS_TC_ARGS); break;
        default: rc = launch_tc<4, true>(GAGS_TC_ARGS); break;
      }
    }
#undef GAGS_TC_ARGS
    if (rc != 0) return rc;
  }
  return 0;
}

extern int g_fwd_blend_persistent;
int gags_blend_fwd_from_cache_persistent(const float *colors, int32_t D, const float *background,
                                         int32_t width, int32_t height, const int32_t *offsets,
                                         const unsigned char *wc, const int32_t *wmeta,
                                         const int32_t *wlist, int32_t *wcount, const float *alphas,
                                         float *render, cudaStream_t st);

// ---- the forward in two passes, split at the weight-tile cache -------------------------------------
// Pass 1 (weights): everything that depends on the geometry only — tile walk, exact cull, alpha,
// transmittance chain — writes alphas (+ last_ids), the blended batches' weight tiles and id lists.
// Pass 2 (blend): render = cached weights x features on the tensor cores.  Between the two the
// caller knows exactly which feature rows the view reads (gags_blend_cache_mark_rows), which is
// what lets a lazily-updated feature table be brought up to date for those rows only.
extern "C" int gags_blend_fwd_weights(const float *geom, int32_t width, int32_t height,
                                      const int32_t *offsets, const int32_t *flatten_ids,
                                      float *alphas, int32_t *last_ids, void *wcache,
                                      int32_t *wmeta, int32_t *wlist, int32_t *wcount,
                                      void *stream) {
  if (!geom || !offsets || !alphas || !wcache || !wmeta || !wlist || !wcount || width <= 0 ||
      height <= 0)
    return GAGS_EINVAL;
  if (g_gags_blend_impl == 1) return GAGS_EINVAL;
  if (!gags_aligned16(geom) || !gags_aligned16(wcache)) return GAGS_EALIGN;
  return launch_tc<0, true, 1>(geom, nullptr, 64, 0, 64, nullptr, width, height, offsets,
                               flatten_ids, nullptr, alphas, last_ids,
                               reinterpret_cast<unsigned char *>(wcache), wmeta, wlist, wcount,
                               (cudaStream_t)stream);
}

extern "C" int gags_blend_fwd_from_cache(const float *colors, int32_t D, const float *background,
                                         int32_t width, int32_t height, const int32_t *offsets,
                                         const void *wcache, const int32_t *wmeta,
                                         const int32_t *wlist, const int32_t *wcount,
                                         const float *alphas, float *render, void *stream) {
  if (!colors || !offsets || !wcache || !wmeta || !wlist || !wcount || !alphas || !render ||
      width <= 0 || height <= 0)
    return GAGS_EINVAL;
  if (D <= 32 || D % 16 != 0 || g_gags_blend_impl == 1) return GAGS_EINVAL;
  if (!gags_aligned16(colors) || !gags_aligned16(wcache) || !gags_aligned16(render)) return GAGS_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char *wc = const_cast<unsigned char *>(reinterpret_cast<const unsigned char *>(wcache));
  if (g_fwd_blend_persistent && (D % 4) == 0 && g_fwd_tma_epilogue)
    return gags_blend_fwd_from_cache_persistent(colors, D, background, width, height, offsets, wc,
                                                wmeta, wlist, const_cast<int32_t *>(wcount), alphas,
                                                render, st);
  for (int ch0 = 0; ch0 < D; ch0 += 256) {
    const int nch = (D - ch0) < 256 ? (D - ch0) : 256;
    const int natom = (nch + 63) / 64;
    int rc;
#define GAGS_TC2_ARGS nullptr, colors, D, ch0, nch, background, width, height, offsets, nullptr,      \
                      render, const_cast<float *>(alphas), nullptr, wc, const_cast<int *>(wmeta),   \
                      const_cast<int *>(wlist), const_cast<int *>(wcount), st
    switch (natom) {
      case 1: rc = launch_tc<1, true, 2>(GAGS_TC2_ARGS); break;
      case 2: rc = launch_tc<2, true, 2>(GAGS_TC2_ARGS); break;
      case 3: rc = launch_tc<3, true, 2>(GAGS_TC2_ARGS); break;
      default: rc = launch_tc<4, true, 2>(GAGS_TC2_ARGS); break;
    }
#undef GAGS_TC2_ARGS
    if (rc != 0) return rc;
  }
  return 0;
}

// ---- the blend pass as a persistent kernel ---------------------------------------------------------
// A half tile has only ~4 blended batches at BASELINE config 3, so a one-CTA-per-half-tile blend pass
// spends about a third of each CTA's life in its prologue (TMEM allocation, barrier set-up, launch)
// and in the TMEM -> shared -> TMA-store epilogue that nothing of ITS OWN overlaps.  Here ONE CTA per
// SM owns the whole TMEM (two 256-column accumulators) and 3 A/B stages and pulls half tiles from a
// global counter:
//   warp 0      producer: fetches the tile, reads its batch metadata a batch ahead, publishes the
//               Gaussian ids of each batch and lands its cached weight tile in the A stage (bulk copy)
//   warp 1      MMA issuer: accumulator (tile k) & 1; commit -> stage free, last batch -> accfull
//   warps 2-9   converters: 4 feature rows each per batch (gather, bf16 hi/lo split, swizzled store)
//   warps 10-17 epilogue: drain accumulator (tile k) & 1 while the MMAs of tile k+1 fill the other —
//               tcgen05.ld -> (+ T bg) -> staging box -> one TMA tensor store per 32 ch x 32 px item
// The stage ring and every barrier phase run on GLOBAL counters (batches / tiles seen by this CTA),
// as in the persistent cached backward.  Bit-identical to the one-CTA-per-half-tile blend pass: the
// same tiles, the same conversions, the same MMA order per accumulator element.
namespace {

constexpr int PB_THREADS = 576;
constexpr int PB_NST = 3;          // A/B stages
constexpr int PB_TQ = 4;           // tile-info ring

struct PbCtl {
  uint64_t list[PB_NST], full[PB_NST], free_[PB_NST], accfull[2], accfree[2];
  uint64_t tq_full[PB_TQ], tq_free[PB_TQ];
  uint32_t tmem_base;
  int tq_tile[PB_TQ], tq_nbat[PB_TQ];
  int gcount[PB_NST];
  int gid[PB_NST][TC_KB];
  int bg_nonzero[8];
  alignas(16) float bgs[256];
};

template <int NATOM>
struct PbLayout {
  static constexpr int BPART = NATOM * 4096;
  static constexpr int A_OFF = 0;                                   // PB_NST x 16 KB
  static constexpr int B_OFF = PB_NST * 16384;                      // [stage][part][BPART]
  static constexpr int BOX_OFF = B_OFF + PB_NST * 2 * BPART;        // 8 epilogue warps x 2 x 4 KB
  static constexpr int CTL_OFF = BOX_OFF + 8 * 2 * 4096;
  static constexpr int BYTES = CTL_OFF + (int)sizeof(PbCtl) + 1024;
  static constexpr int MB = (NATOM + 1) / 2;
};
static_assert(PbLayout<4>::BYTES <= 232448, "persistent blend pass must fit one SM");

template <int NATOM>
__global__ void __launch_bounds__(PB_THREADS, 1)
blend_fwd_pers(const float *__restrict__ colors, int D, int ch0, int nch,
               const float *__restrict__ bg, int W, int H, int tile_w, int ntiles,
               const int *__restrict__ offsets, const float *__restrict__ alphas,
               const unsigned char *__restrict__ wcache, const int *__restrict__ wmeta,
               const int *__restrict__ wlist, const int *__restrict__ wcount,
               int *__restrict__ tilectr, const __grid_constant__ CUtensorMap tmap_render,
               const __grid_constant__ CUtensorMap tmap_wide, int use_wide) {
  using L = PbLayout<NATOM>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *sA = sm + L::A_OFF;
  unsigned char *sB = sm + L::B_OFF;
  PbCtl &ctl = *reinterpret_cast<PbCtl *>(sm + L::CTL_OFF);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int k = 0; k < PB_NST; ++k) {
      mbar_init(&ctl.list[k], 1);
      mbar_init(&ctl.full[k], 9);                  // 8 converter warps + the producer's expect_tx
      mbar_init(&ctl.free_[k], 1);
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&ctl.accfull[k], 1);
      mbar_init(&ctl.accfree[k], 8);               // epilogue warps
    }
    for (int k = 0; k < PB_TQ; ++k) {
      mbar_init(&ctl.tq_full[k], 1);
      mbar_init(&ctl.tq_free[k], 8);               // epilogue warps (the last readers of a tile)
    }
    mbar_fence_init();
  }
  if (tid < 256) {
    const float b = (bg != nullptr && tid < nch) ? __ldg(bg + ch0 + tid) : 0.f;
    ctl.bgs[tid] = b;
    const bool nzw = __any_sync(0xffffffffu, b != 0.f);
    if (lane == 0) ctl.bg_nonzero[warp] = nzw ? 1 : 0;
  }
  if (warp == 1) tmem_alloc<512>(&ctl.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = ctl.tmem_base;

  // tile k of this CTA, as published by the producer: (linear half-tile index or -1, batch count)
  auto tile_info = [&](int k, int &j, int &nbat) {
    const int q = k & (PB_TQ - 1);
    mbar_wait_bounded(&ctl.tq_full[q], (uint32_t)((k / PB_TQ) & 1));
    j = *reinterpret_cast<volatile int *>(&ctl.tq_tile[q]);
    nbat = *reinterpret_cast<volatile int *>(&ctl.tq_nbat[q]);
  };

  if (warp == 0) {
    // ======================= producer ==============================================================
    int gs = 0;
    for (int k = 0;; ++k) {
      const int q = k & (PB_TQ - 1);
      if (k >= PB_TQ) mbar_wait_bounded(&ctl.tq_free[q], (uint32_t)(((k / PB_TQ) - 1) & 1));
      int j = 0;
      if (lane == 0) j = atomicAdd(tilectr, 1);
      j = __shfl_sync(0xffffffffu, j, 0);
      if (j >= ntiles) j = -1;
      int nbat = 0, hbase = 0;
      if (j >= 0) {
        const int by = j / tile_w, bx = j - by * tile_w;
        const int tile = (by >> 1) * tile_w + bx;
        const int s = __ldg(offsets + tile), e = __ldg(offsets + tile + 1);
        const int cbase = (s >> 5) + tile;
        hbase = 2 * cbase + (by & 1) * ((e >> 5) + tile + 1 - cbase);
        nbat = __ldg(wcount + j);
      }
      if (lane == 0) {
        ctl.tq_tile[q] = j;
        ctl.tq_nbat[q] = nbat;
        __threadfence_block();
        mbar_arrive(&ctl.tq_full[q]);
      }
      if (j < 0) break;
      int wl = (lane < nbat) ? __ldg(wlist + hbase + lane) : 0;
      size_t slot_n = 0;
      int gid_n = -1;
      if (nbat > 0) {
        slot_n = (size_t)hbase + (size_t)__shfl_sync(0xffffffffu, wl, 0);
        gid_n = __ldg(wmeta + slot_n * TC_KB + lane);
      }
      for (int i = 0; i < nbat; ++i, ++gs) {
        const int st = gs % PB_NST;
        const size_t slot = slot_n;
        const int gid = gid_n;
        if (i + 1 < nbat) {
          if (((i + 1) & 31) == 0) wl = (i + 1 + lane < nbat) ? __ldg(wlist + hbase + i + 1 + lane) : 0;
          slot_n = (size_t)hbase + (size_t)__shfl_sync(0xffffffffu, wl, (i + 1) & 31);
          gid_n = __ldg(wmeta + slot_n * TC_KB + lane);
        }
        if (gs >= PB_NST) mbar_wait_bounded(&ctl.free_[st], (uint32_t)(((gs / PB_NST) - 1) & 1));
        ctl.gid[st][lane] = gid;
        const int nb = __popc(__ballot_sync(0xffffffffu, gid >= 0));
        if (lane == 0) ctl.gcount[st] = nb;
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          mbar_arrive(&ctl.list[st]);
          mbar_expect_tx(&ctl.full[st], 16384u);
          bulk_g2s(sA + st * 16384, wcache + slot * 16384, 16384u, &ctl.full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer ============================================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, true, false);
      const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sA), 16, 1024);
      const uint64_t f_desc0 = umma_desc_sw128(smem_u32(sB), 4096, 1024);
      int gs = 0;
      for (int k = 0;; ++k) {
        int j, nbat;
        tile_info(k, j, nbat);
        if (j < 0) break;
        const int buf = k & 1;
        if (k >= 2) mbar_wait_bounded(&ctl.accfree[buf], (uint32_t)(((k >> 1) - 1) & 1));
        tc_fence_after();
        uint32_t acc = 0;
        for (int i = 0; i < nbat; ++i, ++gs) {
          const int st = gs % PB_NST;
          mbar_wait_bounded(&ctl.list[st], (uint32_t)((gs / PB_NST) & 1));
          const int nb = *reinterpret_cast<volatile int *>(&ctl.gcount[st]);
          mbar_wait_bounded(&ctl.full[st], (uint32_t)((gs / PB_NST) & 1));
          tc_fence_after();
          const int nk = (nb + 15) >> 4;
#pragma unroll 1
          for (int ks = 0; ks < nk; ++ks) {
            const uint64_t whi = w_desc0 + (uint64_t)((st * 16384 + ks * 32) >> 4);
            const uint64_t wlo = whi + (uint64_t)(64 >> 4);
#pragma unroll
            for (int mb = 0; mb < L::MB; ++mb) {
              const uint64_t fhi =
                  f_desc0 + (uint64_t)(((st * 2) * L::BPART + mb * 8192 + ks * 2048) >> 4);
              const uint64_t flo = fhi + (uint64_t)(L::BPART >> 4);
              const uint32_t d = tb + (uint32_t)(buf * 256 + mb * 128);
              umma_bf16_ss(d, fhi, whi, idesc, acc);
              umma_bf16_ss(d, flo, whi, idesc, 1);
              umma_bf16_ss(d, fhi, wlo, idesc, 1);
            }
            acc = 1;
          }
          umma_commit(&ctl.free_[st]);
        }
        if (nbat > 0) umma_commit(&ctl.accfull[buf]);      // every MMA of the tile has completed
        else mbar_arrive(&ctl.accfull[buf]);
      }
    }
    __syncwarp();
  } else if (warp < 10) {
    // ======================= converters ============================================================
    const int cw = warp - 2;
    const int n0 = lane * 8;
    const bool chan_ok = n0 < nch;
    const uint32_t coff = (uint32_t)(n0 >> 6) * 4096u + (uint32_t)((n0 & 63) >> 3) * 16u;
    const float *cbase = colors + ch0 + n0;
    int gs = 0;
    for (int k = 0;; ++k) {
      int j, nbat;
      tile_info(k, j, nbat);
      if (j < 0) break;
      for (int i = 0; i < nbat; ++i, ++gs) {
        const int st = gs % PB_NST;
        // the list of batch gs is only published after the MMAs of batch gs - PB_NST have released
        // the stage: seeing it also means the B stage may be overwritten
        mbar_wait_bounded(&ctl.list[st], (uint32_t)((gs / PB_NST) & 1));
        const int nb = *reinterpret_cast<volatile int *>(&ctl.gcount[st]);
        const int nbr = (nb + 15) & ~15;
        float4 v[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = cw * 4 + r;
          v[r][0] = v[r][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < nb && chan_ok) {
            const int gid = ctl.gid[st][row];
            const float4 *src = reinterpret_cast<const float4 *>(cbase + (size_t)gid * D);
            v[r][0] = __ldg(src);
            v[r][1] = __ldg(src + 1);
          }
        }
        unsigned char *bhi = sB + (st * 2 + 0) * L::BPART;
        unsigned char *blo = sB + (st * 2 + 1) * L::BPART;
        if (chan_ok) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int row = cw * 4 + r;
            if (row < nbr) {
              uint4 h, l;
              split_pack2(v[r][0].x, v[r][0].y, h.x, l.x);
              split_pack2(v[r][0].z, v[r][0].w, h.y, l.y);
              split_pack2(v[r][1].x, v[r][1].y, h.z, l.z);
              split_pack2(v[r][1].z, v[r][1].w, h.w, l.w);
              const uint32_t off = (uint32_t)(row >> 3) * 1024u +
                                   sw128((uint32_t)(row & 7) * 128u + (coff & 127u)) + (coff & ~127u);
              *reinterpret_cast<uint4 *>(bhi + off) = h;
              *reinterpret_cast<uint4 *>(blo + off) = l;
            }
          }
        }
        fence_async_smem();
        mbar_arrive_warp(&ctl.full[st]);
      }
    }
  } else {
    // ======================= epilogue ==============================================================
    const int ew = warp - 10;                      // 0..7
    const int q = warp & 3;                        // the TMEM lane quarter this warp may read
    const int sub = (ew >> 2);                     // the two warps of a quarter split the items
    bool use_bg = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) use_bg = use_bg || (ctl.bg_nonzero[k] != 0);
    unsigned char *boxes = sm + L::BOX_OFF + ew * 8192;
    int nbox = 0;
    for (int k = 0;; ++k) {
      int j, nbat;
      tile_info(k, j, nbat);
      if (j < 0) break;
      const int buf = k & 1;
      const int by = j / tile_w, bx = j - by * tile_w;
      const int x0 = bx * GAGS_TILE, y0 = by * 8;
      mbar_wait_bounded(&ctl.accfull[buf], (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      if (NATOM == 4 && use_wide) {
        // 256 channels: the eight warps are the eight (128-channel block, lane quarter) pairs, so
        // together they hold every channel of a 8x4 pixel block — staged as [32 px][256 ch] and
        // stored by ONE tensor store per block: each pixel's 1 KB row reaches DRAM in one piece
        const int mb = sub;
        const int ch = mb * 128 + q * 32 + lane;
        unsigned char *wide = sm + L::BOX_OFF;
#pragma unroll 1
        for (int pc = 0; pc < 4; ++pc) {
          const int xb = x0 + ((pc & 1) << 3), yb = y0 + ((pc >> 1) << 2);
          if (yb >= H || xb >= W) continue;                          // uniform over the 8 warps
          uint32_t r[32];
          if (nbat > 0) {
            tmem_ld_32x32(tb + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + mb * 128 + pc * 32), r);
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t) r[t] = 0u;
          }
          if (use_bg) {
            const int px = xb + (lane & 7), py = yb + (lane >> 3);
            const float Tl = (px < W && py < H) ? 1.f - __ldg(alphas + (size_t)py * W + px) : 0.f;
            const float b = ctl.bgs[ch];
#pragma unroll
            for (int t = 0; t < 32; ++t)
              r[t] = __float_as_uint(fmaf(__shfl_sync(0xffffffffu, Tl, t), b, __uint_as_float(r[t])));
          }
          unsigned char *box = wide + (nbox & 1) * 32768;
          if (ew == 0 && lane == 0) bulk_wait_group_read<1>();      // the box's previous store has read it
          named_bar_sync(1, 256);
#pragma unroll
          for (int t = 0; t < 32; ++t) *reinterpret_cast<uint32_t *>(box + t * 1024 + ch * 4) = r[t];
          fence_async_smem();
          named_bar_sync(2, 256);
          if (ew == 0 && lane == 0) {
            tma_store_3d(&tmap_wide, box, ch0, xb, yb);
            bulk_commit_group();
          }
          ++nbox;
        }
      } else {
#pragma unroll 1
      for (int idx = sub; idx < L::MB * 4; idx += 2) {
        const int mb = idx >> 2, pc = idx & 3;
        const int ch = mb * 128 + q * 32 + lane;
        if (mb * 128 + q * 32 >= nch) continue;                      // warp-uniform
        const int xb = x0 + ((pc & 1) << 3), yb = y0 + ((pc >> 1) << 2);
        if (yb >= H || xb >= W) continue;                            // warp-uniform
        uint32_t r[32];
        if (nbat > 0) {
          tmem_ld_32x32(tb + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + mb * 128 + pc * 32), r);
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) r[t] = 0u;
        }
        if (use_bg) {
          // final transmittance of the block's 32 pixels (lane t <-> pixel t), broadcast per pixel
          const int px = xb + (lane & 7), py = yb + (lane >> 3);
          const float Tl = (px < W && py < H) ? 1.f - __ldg(alphas + (size_t)py * W + px) : 0.f;
          const float b = ch < nch ? ctl.bgs[ch] : 0.f;
#pragma unroll
          for (int t = 0; t < 32; ++t)
            r[t] = __float_as_uint(fmaf(__shfl_sync(0xffffffffu, Tl, t), b, __uint_as_float(r[t])));
        }
        unsigned char *box = boxes + (nbox & 1) * 4096;
        if (lane == 0) bulk_wait_group_read<1>();
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 32; ++t) *reinterpret_cast<uint32_t *>(box + t * 128 + lane * 4) = r[t];
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmap_render, box, ch0 + mb * 128 + q * 32, xb, yb);
          bulk_commit_group();
        }
        ++nbox;
      }
      }
      // the accumulator has been read (tcgen05.ld waited for): the MMAs of tile k + 2 may refill it
      tc_fence_before();
      mbar_arrive_warp(&ctl.accfree[buf]);
      mbar_arrive_warp(&ctl.tq_free[k & (PB_TQ - 1)]);
    }
    if (lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tb);
}

template <int NATOM>
int launch_pers(const float *colors, int D, int ch0, int nch, const float *bg, int W, int H,
                const int *offsets, const float *alphas, const unsigned char *wcache,
                const int *wmeta, const int *wlist, int *wcount, float *render, cudaStream_t st) {
  using L = PbLayout<NATOM>;
  const int tw = (W + GAGS_TILE - 1) / GAGS_TILE;
  const int hh = (H + 7) / 8;
  const int ntiles = tw * hh;
  cudaError_t e = cudaFuncSetAttribute(blend_fwd_pers<NATOM>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES);
  if (e != cudaSuccess) return (int)e;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (!make_render_map(&tmap, render, D, W, H)) return GAGS_EINVAL;
  CUtensorMap tmapw;
  memset(&tmapw, 0, sizeof(tmapw));
  const int use_wide = (NATOM == 4 && nch == 256 && make_render_map_wide(&tmapw, render, D, W, H)) ? 1 : 0;
  int *tilectr = wcount + ntiles;                    // the caller's extra int behind the counts
  e = cudaMemsetAsync(tilectr, 0, sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  const int grid = ntiles < gags_sm_count() ? ntiles : gags_sm_count();
  blend_fwd_pers<NATOM><<<grid, PB_THREADS, L::BYTES, st>>>(colors, D, ch0, nch, bg, W, H, tw, ntiles,
                                                            offsets, alphas, wcache, wmeta, wlist,
                                                            wcount, tilectr, tmap, tmapw, use_wide);
  return (int)cudaGetLastError();
}

}  // namespace

// 1 = the persistent blend pass (default), 0 = one CTA per half tile
int g_fwd_blend_persistent = 1;
extern "C" int gags_set_blend_pass(int32_t persistent) {
  if (persistent != 0 && persistent != 1) return GAGS_EINVAL;
  g_fwd_blend_persistent = persistent;
  return 0;
}

int gags_blend_fwd_from_cache_persistent(const float *colors, int32_t D, const float *background,
                                         int32_t width, int32_t height, const int32_t *offsets,
                                         const unsigned char *wc, const int32_t *wmeta,
                                         const int32_t *wlist, int32_t *wcount, const float *alphas,
                                         float *render, cudaStream_t st) {
  for (int ch0 = 0; ch0 < D; ch0 += 256) {
    const int nch = (D - ch0) < 256 ? (D - ch0) : 256;
    const int natom = (nch + 63) / 64;
    int rc;
#define GAGS_PB_ARGS colors, D, ch0, nch, background, width, height, offsets, alphas, wc, wmeta,   \
                     wlist, wcount, render, st
    switch (natom) {
      case 1: rc = launch_pers<1>(GAGS_PB_ARGS); break;
      case 2: rc = launch_pers<2>(GAGS_PB_ARGS); break;
      case 3: rc = launch_pers<3>(GAGS_PB_ARGS); break;
      default: rc = launch_pers<4>(GAGS_PB_ARGS); break;
    }
#undef GAGS_PB_ARGS
    if (rc != 0) return rc;
  }
  return 0;
}
