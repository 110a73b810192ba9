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
