// sh.cu — K3 spherical-harmonics colours and their VJP.  Replaces gsplat spherical_harmonics
// fwd/bwd reached from /root/reference/gaussian_renderer/__init__.py:51-53,56-70.  The basis and its
// sign convention are those of /root/reference/utils/sh_utils.py:57-112 (degrees 0..4); the
// "+0.5, clamp at 0" step follows gsplat.rasterization.
// Roofline: HBM — 12*K + 12 bytes in, 12 out per Gaussian.
#include "common.cuh"

namespace {

__constant__ float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                             -1.0925484305920792f, 0.5462742152960396f};
__constant__ float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                             0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                             -0.5900435899266435f};
__constant__ float kC4[9] = {2.5033429417967046f, -1.7701307697799304f, 0.9461746957575601f,
                             -0.6690465435572892f, 0.10578554691520431f, -0.6690465435572892f,
                             0.47308734787878004f, -1.7701307697799304f, 0.6258357354491761f};
#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f

// basis values B[0..nb) at unit direction (x,y,z)
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float *B) {
  B[0] = SH_C0;
  if (deg < 1) return;
  B[1] = -SH_C1 * y; B[2] = SH_C1 * z; B[3] = -SH_C1 * x;
  if (deg < 2) return;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  B[4] = kC2[0] * xy; B[5] = kC2[1] * yz; B[6] = kC2[2] * (2.f * zz - xx - yy);
  B[7] = kC2[3] * xz; B[8] = kC2[4] * (xx - yy);
  if (deg < 3) return;
  B[9] = kC3[0] * y * (3.f * xx - yy); B[10] = kC3[1] * xy * z;
  B[11] = kC3[2] * y * (4.f * zz - xx - yy); B[12] = kC3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
  B[13] = kC3[4] * x * (4.f * zz - xx - yy); B[14] = kC3[5] * z * (xx - yy);
  B[15] = kC3[6] * x * (xx - 3.f * yy);
  if (deg < 4) return;
  B[16] = kC4[0] * xy * (xx - yy); B[17] = kC4[1] * yz * (3.f * xx - yy);
  B[18] = kC4[2] * xy * (7.f * zz - 1.f); B[19] = kC4[3] * yz * (7.f * zz - 3.f);
  B[20] = kC4[4] * (zz * (35.f * zz - 30.f) + 3.f); B[21] = kC4[5] * xz * (7.f * zz - 3.f);
  B[22] = kC4[6] * (xx - yy) * (7.f * zz - 1.f); B[23] = kC4[7] * xz * (xx - 3.f * yy);
  B[24] = kC4[8] * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
}

// gradient of sum_k B_k * g_k w.r.t. (x,y,z), g_k = sum_c coeff[k][c] * v_color[c]
__device__ __forceinline__ void sh_basis_vjp(int deg, float x, float y, float z, const float *g,
                                             float &vx, float &vy, float &vz) {
  vx = vy = vz = 0.f;
  if (deg < 1) return;
  vy += -SH_C1 * g[1]; vz += SH_C1 * g[2]; vx += -SH_C1 * g[3];
  if (deg < 2) return;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  vx += kC2[0] * y * g[4];              vy += kC2[0] * x * g[4];
  vy += kC2[1] * z * g[5];              vz += kC2[1] * y * g[5];
  vx += kC2[2] * (-2.f * x) * g[6];     vy += kC2[2] * (-2.f * y) * g[6];  vz += kC2[2] * 4.f * z * g[6];
  vx += kC2[3] * z * g[7];              vz += kC2[3] * x * g[7];
  vx += kC2[4] * 2.f * x * g[8];        vy += kC2[4] * (-2.f * y) * g[8];
  if (deg < 3) return;
  vx += kC3[0] * 6.f * xy * g[9];       vy += kC3[0] * (3.f * xx - 3.f * yy) * g[9];
  vx += kC3[1] * yz * g[10];            vy += kC3[1] * xz * g[10];         vz += kC3[1] * xy * g[10];
  vx += kC3[2] * (-2.f * xy) * g[11];   vy += kC3[2] * (4.f * zz - xx - 3.f * yy) * g[11];
  vz += kC3[2] * 8.f * yz * g[11];
  vx += kC3[3] * (-6.f * xz) * g[12];   vy += kC3[3] * (-6.f * yz) * g[12];
  vz += kC3[3] * (6.f * zz - 3.f * xx - 3.f * yy) * g[12];
  vx += kC3[4] * (4.f * zz - 3.f * xx - yy) * g[13]; vy += kC3[4] * (-2.f * xy) * g[13];
  vz += kC3[4] * 8.f * xz * g[13];
  vx += kC3[5] * 2.f * xz * g[14];      vy += kC3[5] * (-2.f * yz) * g[14]; vz += kC3[5] * (xx - yy) * g[14];
  vx += kC3[6] * (3.f * xx - 3.f * yy) * g[15]; vy += kC3[6] * (-6.f * xy) * g[15];
  if (deg < 4) return;
  vx += kC4[0] * (3.f * xx * y - yy * y) * g[16]; vy += kC4[0] * (xx * x - 3.f * x * yy) * g[16];
  vx += kC4[1] * 6.f * xy * z * g[17];  vy += kC4[1] * z * (3.f * xx - 3.f * yy) * g[17];
  vz += kC4[1] * y * (3.f * xx - yy) * g[17];
  vx += kC4[2] * y * (7.f * zz - 1.f) * g[18]; vy += kC4[2] * x * (7.f * zz - 1.f) * g[18];
  vz += kC4[2] * 14.f * xy * z * g[18];
  vy += kC4[3] * z * (7.f * zz - 3.f) * g[19]; vz += kC4[3] * y * (21.f * zz - 3.f) * g[19];
  vz += kC4[4] * (140.f * zz * z - 60.f * z) * g[20];
  vx += kC4[5] * z * (7.f * zz - 3.f) * g[21]; vz += kC4[5] * x * (21.f * zz - 3.f) * g[21];
  vx += kC4[6] * 2.f * x * (7.f * zz - 1.f) * g[22]; vy += kC4[6] * (-2.f * y) * (7.f * zz - 1.f) * g[22];
  vz += kC4[6] * (xx - yy) * 14.f * z * g[22];
  vx += kC4[7] * z * (3.f * xx - 3.f * yy) * g[23]; vy += kC4[7] * (-6.f * xy * z) * g[23];
  vz += kC4[7] * x * (xx - 3.f * yy) * g[23];
  vx += kC4[8] * (4.f * xx * x - 12.f * x * yy) * g[24]; vy += kC4[8] * (-12.f * xx * y + 4.f * yy * y) * g[24];
}

__global__ void __launch_bounds__(256)
sh_fwd_kernel(int deg, const float *__restrict__ means, const float *__restrict__ campos,
              const float *__restrict__ coeffs, int K, const int *__restrict__ radii, long long N,
              float *__restrict__ colors, int stride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float r = 0.f, g = 0.f, b = 0.f;
  const float cxp = __ldg(campos), cyp = __ldg(campos + 1), czp = __ldg(campos + 2);
  if (!radii || radii[i] > 0) {
    float dx = means[i * 3] - cxp, dy = means[i * 3 + 1] - cyp, dz = means[i * 3 + 2] - czp;
    const float inv = 1.0f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-20f);
    dx *= inv; dy *= inv; dz *= inv;
    float B[25];
    sh_basis(deg, dx, dy, dz, B);
    const int nb = (deg + 1) * (deg + 1);
    const float *c = coeffs + (size_t)i * K * 3;
    for (int k = 0; k < nb; ++k) {
      r = fmaf(B[k], c[k * 3], r); g = fmaf(B[k], c[k * 3 + 1], g); b = fmaf(B[k], c[k * 3 + 2], b);
    }
  }
  float *o = colors + (size_t)i * stride;
  o[0] = fmaxf(r + 0.5f, 0.f); o[1] = fmaxf(g + 0.5f, 0.f); o[2] = fmaxf(b + 0.5f, 0.f);
}

__global__ void __launch_bounds__(256)
sh_bwd_kernel(int deg, const float *__restrict__ means, const float *__restrict__ campos,
              const float *__restrict__ coeffs, int K, const int *__restrict__ radii, long long N,
              const float *__restrict__ v_colors, int vstride, float *__restrict__ v_coeffs,
              float *__restrict__ v_means) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float *vc = v_coeffs + (size_t)i * K * 3;
  const bool vis = !radii || radii[i] > 0;
  const int nb = (deg + 1) * (deg + 1);
  if (!vis) {
    for (int k = 0; k < K * 3; ++k) vc[k] = 0.f;
    return;
  }
  const float cxp = __ldg(campos), cyp = __ldg(campos + 1), czp = __ldg(campos + 2);
  float dx = means[i * 3] - cxp, dy = means[i * 3 + 1] - cyp, dz = means[i * 3 + 2] - czp;
  const float len = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-20f);
  const float inv = 1.0f / len;
  const float x = dx * inv, y = dy * inv, z = dz * inv;
  float B[25];
  sh_basis(deg, x, y, z, B);
  const float *c = coeffs + (size_t)i * K * 3;
  // clamp_min(sh + 0.5, 0): gradient passes where the forward value was positive
  float r = 0.f, g = 0.f, b = 0.f;
  for (int k = 0; k < nb; ++k) {
    r = fmaf(B[k], c[k * 3], r); g = fmaf(B[k], c[k * 3 + 1], g); b = fmaf(B[k], c[k * 3 + 2], b);
  }
  const float *vin = v_colors + (size_t)i * vstride;
  const float vr = (r + 0.5f > 0.f) ? vin[0] : 0.f;
  const float vg = (g + 0.5f > 0.f) ? vin[1] : 0.f;
  const float vb = (b + 0.5f > 0.f) ? vin[2] : 0.f;
  float gk[25];
  for (int k = 0; k < K; ++k) {
    const float bk = (k < nb) ? B[k] : 0.f;
    vc[k * 3] = bk * vr; vc[k * 3 + 1] = bk * vg; vc[k * 3 + 2] = bk * vb;
    if (k < nb) gk[k] = c[k * 3] * vr + c[k * 3 + 1] * vg + c[k * 3 + 2] * vb;
  }
  if (v_means) {
    float vx, vy, vz;
    sh_basis_vjp(deg, x, y, z, gk, vx, vy, vz);
    // through the normalisation d = v / |v|
    const float dot = x * vx + y * vy + z * vz;
    v_means[i * 3] += (vx - x * dot) * inv;
    v_means[i * 3 + 1] += (vy - y * dot) * inv;
    v_means[i * 3 + 2] += (vz - z * dot) * inv;
  }
}

}  // namespace

extern "C" int gags_sh_fwd(int32_t degree, const float *means, const float *campos,
                           const float *coeffs, int32_t K, const int32_t *radii, int64_t N,
                           float *colors, int32_t out_stride, void *stream) {
  if (!means || !campos || !coeffs || !colors || N < 0) return GAGS_EINVAL;
  if (degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) || out_stride < 3) return GAGS_EINVAL;
  if (N == 0) return 0;
  sh_fwd_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      degree, means, campos, coeffs, K, radii, (long long)N, colors,
      out_stride);
  GAGS_CHECK_LAUNCH();
  return 0;
}

extern "C" int gags_sh_bwd(int32_t degree, const float *means, const float *campos,
                           const float *coeffs, int32_t K, const int32_t *radii, int64_t N,
                           const float *v_colors, int32_t v_stride, float *v_coeffs, float *v_means,
                           void *stream) {
  if (!means || !campos || !coeffs || !v_colors || !v_coeffs || N < 0) return GAGS_EINVAL;
  if (degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) || K > 25 || v_stride < 3)
    return GAGS_EINVAL;
  if (N == 0) return 0;
  sh_bwd_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      degree, means, campos, coeffs, K, radii, (long long)N, v_colors,
      v_stride, v_coeffs, v_means);
  GAGS_CHECK_LAUNCH();
  return 0;
}
