// api.cu — library identification and error strings for the C-ABI in include/gags_b200.h.
#include "common.cuh"

extern "C" const char *gags_version(void) { return "gags_b200 0.1.0"; }
extern "C" const char *gags_build_arch(void) { return "sm_100a"; }
extern "C" const char *gags_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case GAGS_EINVAL: return "GAGS_EINVAL: invalid argument";
    case GAGS_EALIGN: return "GAGS_EALIGN: pointer or row stride not 16-byte aligned";
    case GAGS_ESMALL: return "GAGS_ESMALL: workspace too small";
    case GAGS_ERANGE: return "GAGS_ERANGE: value outside the supported range";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown gags error";
}

// 0 = auto (tensor-core path whenever the shape allows), 1 = SIMT kernels only, 2 = tensor-core
// kernels required (GAGS_EINVAL when the shape does not fit).  Process-wide; used by the parity
// tests to pin one implementation against the other.
int g_gags_blend_impl = 0;
extern "C" int gags_set_blend_impl(int32_t impl) {
  if (impl < 0 || impl > 2) return GAGS_EINVAL;
  g_gags_blend_impl = impl;
  return 0;
}
extern "C" int gags_get_blend_impl(void) { return g_gags_blend_impl; }
