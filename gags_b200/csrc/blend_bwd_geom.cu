// blend_bwd_geom.cu — wide-D (32 < D <= 256) geometry/opacity gradients of the blend (K8b).
#include "blend_common.cuh"

int gags_blend_bwd_geom_wide(const float *geom, const float *colors, int32_t D,
                             const float *background, int32_t width, int32_t height,
                             const int32_t *offsets, const int32_t *flatten_ids,
                             const float *render_alphas, const int32_t *last_ids,
                             const float *v_render, const float *v_alphas, float *v_means2d,
                             float *v_conics, float *v_opacities, cudaStream_t st) {
  return GAGS_ERANGE;  // TODO(round 1): replaced below once the narrow path is parity-green
}
