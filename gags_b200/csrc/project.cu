// project.cu — K1 projection + frustum cull (+ fused activations, tile count, packed blend record)
// and K2 its VJP.  Replaces gsplat fully_fused_projection_{fwd,bwd} reached from
// /root/reference/gaussian_renderer/__init__.py:56-70; maths = SURVEY.md Appendix A.1 / A.7.
//
// Roofline: HBM.  Algorithmic bytes per Gaussian: 44 in (xyz 12, quat 16, scale 12, opacity 4),
// 32 (geom) + 4 (tiles) + 32 (API arrays radii/means2d/depths/conics/opac) out.
// [N,3] AoS inputs are staged through shared memory with coalesced 16-byte loads; [N,4] quats and
// the 32-byte geom record are already 16-byte-per-thread coalesced.
#include "common.cuh"

namespace {

constexpr int PB = 256;  // Gaussians per block

// cooperative coalesced copy of `nfloat` floats global->shared (16-B vectors when aligned)
__device__ __forceinline__ void stage_in(float *s, const float *g, int nfloat, bool vec_ok) {
  if (vec_ok) {
    const int nv = nfloat >> 2;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    float4 *s4 = reinterpret_cast<float4 *>(s);
    for (int i = threadIdx.x; i < nv; i += PB) s4[i] = ldg_nc4(g4 + i);
    for (int i = (nv << 2) + threadIdx.x; i < nfloat; i += PB) s[i] = __ldg(g + i);
  } else {
    for (int i = threadIdx.x; i < nfloat; i += PB) s[i] = __ldg(g + i);
  }
}
__device__ __forceinline__ void stage_out(float *g, const float *s, int nfloat, bool vec_ok) {
  if (vec_ok) {
    const int nv = nfloat >> 2;
    float4 *g4 = reinterpret_cast<float4 *>(g);
    const float4 *s4 = reinterpret_cast<const float4 *>(s);
    for (int i = threadIdx.x; i < nv; i += PB) g4[i] = s4[i];
    for (int i = (nv << 2) + threadIdx.x; i < nfloat; i += PB) g[i] = s[i];
  } else {
    for (int i = threadIdx.x; i < nfloat; i += PB) g[i] = s[i];
  }
}

struct Proj {
  float x, y, z;          // camera-space mean
  float Rq[9];            // rotation from the normalised quaternion
  float s[3];             // activated scales
  float S[6];             // camera-space covariance (00,01,02,11,12,22)
  float tx, ty;           // clamped x,y used in the Jacobian
  bool clampx, clampy;
  float c00, c01, c11, det;
};

__device__ __forceinline__ void quat_to_R(float w, float x, float y, float z, float *R) {
  R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
  R[3] = 2.f * (x * y + w * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
  R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

// shared forward maths; returns false when culled by the near/far planes
__device__ __forceinline__ bool project_core(const CamDev &cam, float mx, float my, float mz,
                                             float4 q, float s0, float s1, float s2, Proj &p) {
  const float *R = cam.R;
  p.x = R[0] * mx + R[1] * my + R[2] * mz + cam.t[0];
  p.y = R[3] * mx + R[4] * my + R[5] * mz + cam.t[1];
  p.z = R[6] * mx + R[7] * my + R[8] * mz + cam.t[2];
  if (!(p.z >= cam.near_plane && p.z <= cam.far_plane)) return false;
  const float inv = 1.0f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  quat_to_R(q.x * inv, q.y * inv, q.z * inv, q.w * inv, p.Rq);
  if (cam.flags & GAGS_F_LOG_SCALES) { s0 = expf(s0); s1 = expf(s1); s2 = expf(s2); }
  p.s[0] = s0 * cam.smod; p.s[1] = s1 * cam.smod; p.s[2] = s2 * cam.smod;
  // M = Rq diag(s); Sigma = M M^T
  float M[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) M[i * 3 + j] = p.Rq[i * 3 + j] * p.s[j];
  float Sg[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Sg[i * 3 + j] = M[i * 3] * M[j * 3] + M[i * 3 + 1] * M[j * 3 + 1] + M[i * 3 + 2] * M[j * 3 + 2];
  // Sigma_c = R Sigma R^T
  float RS[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      RS[i * 3 + j] = R[i * 3] * Sg[j] + R[i * 3 + 1] * Sg[3 + j] + R[i * 3 + 2] * Sg[6 + j];
  float Sc[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Sc[i * 3 + j] = RS[i * 3] * R[j * 3] + RS[i * 3 + 1] * R[j * 3 + 1] + RS[i * 3 + 2] * R[j * 3 + 2];
  p.S[0] = Sc[0]; p.S[1] = Sc[1]; p.S[2] = Sc[2]; p.S[3] = Sc[4]; p.S[4] = Sc[5]; p.S[5] = Sc[8];
  // perspective Jacobian with the 0.3-slack frustum clamp
  const float tanx = 0.5f * cam.W / cam.fx, tany = 0.5f * cam.H / cam.fy;
  const float lxp = (cam.W - cam.cx) / cam.fx + 0.3f * tanx, lxn = cam.cx / cam.fx + 0.3f * tanx;
  const float lyp = (cam.H - cam.cy) / cam.fy + 0.3f * tany, lyn = cam.cy / cam.fy + 0.3f * tany;
  const float rz = 1.0f / p.z;
  const float xr = p.x * rz, yr = p.y * rz;
  p.clampx = (xr < -lxn) || (xr > lxp);
  p.clampy = (yr < -lyn) || (yr > lyp);
  p.tx = p.z * fminf(lxp, fmaxf(-lxn, xr));
  p.ty = p.z * fminf(lyp, fmaxf(-lyn, yr));
  const float a = cam.fx * rz, b = -cam.fx * p.tx * rz * rz;
  const float c = cam.fy * rz, d = -cam.fy * p.ty * rz * rz;
  p.c00 = a * a * p.S[0] + 2.f * a * b * p.S[2] + b * b * p.S[5] + cam.eps2d;
  p.c01 = a * c * p.S[1] + a * d * p.S[2] + b * c * p.S[4] + b * d * p.S[5];
  p.c11 = c * c * p.S[3] + 2.f * c * d * p.S[4] + d * d * p.S[5] + cam.eps2d;
  p.det = p.c00 * p.c11 - p.c01 * p.c01;
  return true;
}

__global__ void __launch_bounds__(PB)
project_fwd_kernel(const float *__restrict__ means, const float *__restrict__ quats,
                   const float *__restrict__ scales, const float *__restrict__ opacities,
                   long long N, CamDev cam, int vec_ok, int *__restrict__ radii,
                   float *__restrict__ means2d, float *__restrict__ depths,
                   float *__restrict__ conics, float *__restrict__ opac_out,
                   int *__restrict__ tiles_touched, float *__restrict__ geom) {
  __shared__ __align__(16) float s_mean[PB * 3];
  __shared__ __align__(16) float s_scale[PB * 3];
  cam_resolve(cam);
  const long long base = (long long)blockIdx.x * PB;
  const int cnt = (int)min((long long)PB, N - base);
  stage_in(s_mean, means + base * 3, cnt * 3, vec_ok);
  stage_in(s_scale, scales + base * 3, cnt * 3, vec_ok);
  __syncthreads();
  const int t = threadIdx.x;
  const long long i = base + t;
  int radius = 0;
  float m2x = 0.f, m2y = 0.f, depth = 0.f, ca = 0.f, cb = 0.f, cc = 0.f, op = 0.f;
  int ntile = 0;
  if (t < cnt) {
    const float4 q = __ldg(reinterpret_cast<const float4 *>(quats) + i);
    op = __ldg(opacities + i);
    if (cam.flags & GAGS_F_LOGIT_OPACITY) op = 1.0f / (1.0f + expf(-op));
    Proj p;
    if (project_core(cam, s_mean[t * 3], s_mean[t * 3 + 1], s_mean[t * 3 + 2], q, s_scale[t * 3],
                     s_scale[t * 3 + 1], s_scale[t * 3 + 2], p) && p.det > 0.f) {
      const float rz = 1.0f / p.z;
      const float ux = cam.fx * p.x * rz + cam.cx;
      const float uy = cam.fy * p.y * rz + cam.cy;
      const float bb = 0.5f * (p.c00 + p.c11);
      const float v1 = bb + sqrtf(fmaxf(0.01f, bb * bb - p.det));
      const float rad = ceilf(3.f * sqrtf(v1));
      const bool out = (rad <= cam.radius_clip) || (ux + rad <= 0.f) || (ux - rad >= (float)cam.W) ||
                       (uy + rad <= 0.f) || (uy - rad >= (float)cam.H);
      if (!out) {
        radius = (int)fminf(rad, 2.0e9f);
        m2x = ux; m2y = uy; depth = p.z;
        const float idet = 1.0f / p.det;
        ca = p.c11 * idet; cb = -p.c01 * idet; cc = p.c00 * idet;
        int x0, x1, y0, y1;
        tile_bounds(m2x, m2y, radius, cam.tile_w, cam.tile_h, x0, x1, y0, y1);
        ntile = (x1 - x0) * (y1 - y0);
      }
    }
    radii[i] = radius;
    reinterpret_cast<float2 *>(means2d)[i] = make_float2(m2x, m2y);
    depths[i] = depth;
    if (opac_out) opac_out[i] = op;
    if (tiles_touched) tiles_touched[i] = ntile;
    if (geom) {
      float4 *g4 = reinterpret_cast<float4 *>(geom) + i * 2;
      g4[0] = make_float4(m2x, m2y, ca, cb);
      g4[1] = make_float4(cc, op, depth, (float)radius);
    }
  }
  // conics [N,3]: stage through smem for coalesced stores (reuse s_mean)
  __syncthreads();
  if (t < cnt) { s_mean[t * 3] = ca; s_mean[t * 3 + 1] = cb; s_mean[t * 3 + 2] = cc; }
  __syncthreads();
  stage_out(conics + base * 3, s_mean, cnt * 3, vec_ok);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PB)
project_bwd_kernel(const float *__restrict__ means, const float *__restrict__ quats,
                   const float *__restrict__ scales, long long N, CamDev cam,
                   const int *__restrict__ radii, const float *__restrict__ v_means2d,
                   const float *__restrict__ v_depths, const float *__restrict__ v_conics,
                   float *__restrict__ v_means, float *__restrict__ v_quats,
                   float *__restrict__ v_scales) {
  const long long i = (long long)blockIdx.x * PB + threadIdx.x;
  if (i >= N) return;
  cam_resolve(cam);
  float gm[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f};
  if (radii[i] > 0) {
    const float4 qraw = __ldg(reinterpret_cast<const float4 *>(quats) + i);
    const float sr0 = scales[i * 3], sr1 = scales[i * 3 + 1], sr2 = scales[i * 3 + 2];
    Proj p;
    project_core(cam, means[i * 3], means[i * 3 + 1], means[i * 3 + 2], qraw, sr0, sr1, sr2, p);
    const float idet = 1.0f / p.det;
    const float qa = p.c11 * idet, qb = -p.c01 * idet, qc = p.c00 * idet;   // conic Q
    const float va = v_conics[i * 3], vb = 0.5f * v_conics[i * 3 + 1], vc = v_conics[i * 3 + 2];
    // G = -Q V Q  (full symmetric-matrix gradient w.r.t. cov2d)
    const float t00 = qa * va + qb * vb, t01 = qa * vb + qb * vc;
    const float t10 = qb * va + qc * vb, t11 = qb * vb + qc * vc;
    const float G00 = -(t00 * qa + t01 * qb), G01 = -(t00 * qb + t01 * qc);
    const float G11 = -(t10 * qb + t11 * qc);
    const float rz = 1.0f / p.z, rz2 = rz * rz;
    const float J00 = cam.fx * rz, J02 = -cam.fx * p.tx * rz2;
    const float J11 = cam.fy * rz, J12 = -cam.fy * p.ty * rz2;
    // v_Sigma_c = J^T G J  (3x3 symmetric)
    const float GJ[6] = {G00 * J00, G01 * J11, G00 * J02 + G01 * J12,
                         G01 * J00, G11 * J11, G01 * J02 + G11 * J12};   // G J (2x3)
    float vSc[9];
    vSc[0] = J00 * GJ[0]; vSc[1] = J00 * GJ[1]; vSc[2] = J00 * GJ[2];
    vSc[3] = J11 * GJ[3]; vSc[4] = J11 * GJ[4]; vSc[5] = J11 * GJ[5];
    vSc[6] = J02 * GJ[0] + J12 * GJ[3]; vSc[7] = J02 * GJ[1] + J12 * GJ[4];
    vSc[8] = J02 * GJ[2] + J12 * GJ[5];
    // v_J = 2 G J Sigma_c  (2x3)
    const float Sc[9] = {p.S[0], p.S[1], p.S[2], p.S[1], p.S[3], p.S[4], p.S[2], p.S[4], p.S[5]};
    float vJ[6];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        vJ[r * 3 + c] = 2.f * (GJ[r * 3] * Sc[c] + GJ[r * 3 + 1] * Sc[3 + c] + GJ[r * 3 + 2] * Sc[6 + c]);
    // camera-space mean gradient
    const float vmx = v_means2d[i * 2], vmy = v_means2d[i * 2 + 1];
    float vx = cam.fx * rz * vmx, vy = cam.fy * rz * vmy;
    float vz = -(cam.fx * p.x * vmx + cam.fy * p.y * vmy) * rz2 + (v_depths ? v_depths[i] : 0.f);
    vz += -cam.fx * rz2 * vJ[0] - cam.fy * rz2 * vJ[4];
    vz += 2.f * cam.fx * p.tx * rz2 * rz * vJ[2] + 2.f * cam.fy * p.ty * rz2 * rz * vJ[5];
    const float vtx = -cam.fx * rz2 * vJ[2], vty = -cam.fy * rz2 * vJ[5];
    if (p.clampx) vz += vtx * p.tx * rz; else vx += vtx;
    if (p.clampy) vz += vty * p.ty * rz; else vy += vty;
    const float *R = cam.R;
    gm[0] = R[0] * vx + R[3] * vy + R[6] * vz;
    gm[1] = R[1] * vx + R[4] * vy + R[7] * vz;
    gm[2] = R[2] * vx + R[5] * vy + R[8] * vz;
    // v_Sigma = R^T vSc R
    float A[9], vS[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        A[r * 3 + c] = R[r] * vSc[c] + R[3 + r] * vSc[3 + c] + R[6 + r] * vSc[6 + c];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        vS[r * 3 + c] = A[r * 3] * R[c] + A[r * 3 + 1] * R[3 + c] + A[r * 3 + 2] * R[6 + c];
    // v_M = (vS + vS^T) M ; M = Rq diag(s)
    float vM[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc += (vS[r * 3 + k] + vS[k * 3 + r]) * p.Rq[k * 3 + c] * p.s[c];
        vM[r * 3 + c] = acc;
      }
    float vR[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float acc = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r) { acc += p.Rq[r * 3 + c] * vM[r * 3 + c]; vR[r * 3 + c] = vM[r * 3 + c] * p.s[c]; }
      gs[c] = (cam.flags & GAGS_F_LOG_SCALES) ? acc * p.s[c] : acc * cam.smod;
    }
    const float n = sqrtf(qraw.x * qraw.x + qraw.y * qraw.y + qraw.z * qraw.z + qraw.w * qraw.w);
    const float inv = 1.0f / n;
    const float w = qraw.x * inv, x = qraw.y * inv, y = qraw.z * inv, z = qraw.w * inv;
    float vq[4];
    vq[0] = 2.f * (-z * vR[1] + y * vR[2] + z * vR[3] - x * vR[5] - y * vR[6] + x * vR[7]);
    vq[1] = 2.f * (y * vR[1] + z * vR[2] + y * vR[3] - 2.f * x * vR[4] - w * vR[5] + z * vR[6] + w * vR[7] - 2.f * x * vR[8]);
    vq[2] = 2.f * (-2.f * y * vR[0] + x * vR[1] + w * vR[2] + x * vR[3] + z * vR[5] - w * vR[6] + z * vR[7] - 2.f * y * vR[8]);
    vq[3] = 2.f * (-2.f * z * vR[0] - w * vR[1] + x * vR[2] + w * vR[3] - 2.f * z * vR[4] + y * vR[5] + x * vR[6] + y * vR[7]);
    const float dotq = w * vq[0] + x * vq[1] + y * vq[2] + z * vq[3];
    gq[0] = (vq[0] - w * dotq) * inv; gq[1] = (vq[1] - x * dotq) * inv;
    gq[2] = (vq[2] - y * dotq) * inv; gq[3] = (vq[3] - z * dotq) * inv;
  }
  if (v_means) { v_means[i * 3] = gm[0]; v_means[i * 3 + 1] = gm[1]; v_means[i * 3 + 2] = gm[2]; }
  if (v_quats) reinterpret_cast<float4 *>(v_quats)[i] = make_float4(gq[0], gq[1], gq[2], gq[3]);
  if (v_scales) { v_scales[i * 3] = gs[0]; v_scales[i * 3 + 1] = gs[1]; v_scales[i * 3 + 2] = gs[2]; }
}

__global__ void opacity_bwd_kernel(const float *__restrict__ o, const float *__restrict__ v,
                                   long long N, float *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { const float a = o[i]; out[i] = v[i] * a * (1.f - a); }
}

}  // namespace

extern "C" int gags_project_fwd(const float *means, const float *quats, const float *scales,
                                const float *opacities, int64_t N, const gags_camera_t *cam,
                                int32_t tile_w, int32_t tile_h, int32_t *radii, float *means2d,
                                float *depths, float *conics, float *opac_out,
                                int32_t *tiles_touched, float *geom, void *stream) {
  if (!means || !quats || !scales || !opacities || !cam || !radii || !means2d || !depths || !conics)
    return GAGS_EINVAL;
  if (N < 0 || cam->width <= 0 || cam->height <= 0) return GAGS_EINVAL;
  if (N == 0) return 0;
  if (!gags_aligned16(quats) || (geom && !gags_aligned16(geom)) || (((uintptr_t)means2d) & 7u))
    return GAGS_EALIGN;
  const int vec_ok = gags_aligned16(means) && gags_aligned16(scales) && gags_aligned16(conics);
  const CamDev cd = make_camdev(cam, tile_w, tile_h);
  const unsigned grid = (unsigned)((N + PB - 1) / PB);
  project_fwd_kernel<<<grid, PB, 0, (cudaStream_t)stream>>>(means, quats, scales, opacities,
      (long long)N, cd, vec_ok, radii, means2d, depths, conics, opac_out, tiles_touched, geom);
  GAGS_CHECK_LAUNCH();
  return 0;
}

extern "C" int gags_project_bwd(const float *means, const float *quats, const float *scales,
                                int64_t N, const gags_camera_t *cam, const int32_t *radii,
                                const float *conics, const float *v_means2d, const float *v_depths,
                                const float *v_conics, float *v_means, float *v_quats,
                                float *v_scales, void *stream) {
  (void)conics;
  if (!means || !quats || !scales || !cam || !radii || !v_means2d || !v_conics) return GAGS_EINVAL;
  if (N < 0) return GAGS_EINVAL;
  if (N == 0) return 0;
  if (!gags_aligned16(quats) || (v_quats && !gags_aligned16(v_quats))) return GAGS_EALIGN;
  const CamDev cd = make_camdev(cam, 1, 1);
  const unsigned grid = (unsigned)((N + PB - 1) / PB);
  project_bwd_kernel<<<grid, PB, 0, (cudaStream_t)stream>>>(means, quats, scales, (long long)N, cd,
      radii, v_means2d, v_depths, v_conics, v_means, v_quats, v_scales);
  GAGS_CHECK_LAUNCH();
  return 0;
}

extern "C" int gags_opacity_bwd(const float *opac_act, const float *v_opac, int64_t N,
                                float *v_logit, void *stream) {
  if (!opac_act || !v_opac || !v_logit || N < 0) return GAGS_EINVAL;
  if (N == 0) return 0;
  opacity_bwd_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      opac_act, v_opac, (long long)N, v_logit);
  GAGS_CHECK_LAUNCH();
  return 0;
}
