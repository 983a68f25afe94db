// Version / error-string entry points of libparc_b200.
#include "parc_common.cuh"

extern "C" int parc_abi_version(void) { return PARC_ABI_VERSION; }

extern "C" const char* parc_error_string(int code) {
  switch (code) {
    case PARC_OK: return "ok";
    case PARC_E_NULL: return "a required pointer is NULL";
    case PARC_E_SIZE: return "negative or inconsistent size";
    case PARC_E_MODEL: return "unsupported character model";
    case PARC_E_ALIGN: return "pointer not aligned for 16-byte vector access";
    case PARC_E_LAYOUT: return "row_floats does not match the model's packed-row layout";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown parc error";
}
