// Library-level entry points of libapr_b200: version, status strings, device info.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace apr {

static thread_local char g_cuda_err[512] = "";

void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

int sm_count() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached = n;
  return n;
}

}  // namespace apr

extern "C" {

int apr_abi_version(void) { return APR_ABI_VERSION; }

const char* apr_status_string(int status) {
  switch (status) {
    case APR_OK: return "ok";
    case APR_E_ARG: return "invalid argument (null pointer, bad shape, id out of range, or d not a multiple of 4 in [4,512])";
    case APR_E_ALIGN: return "pointer not 16-byte aligned";
    case APR_E_WORKSPACE: return "workspace too small";
    case APR_E_CUDA: return "CUDA runtime error";
    case APR_E_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

const char* apr_last_cuda_error(void) { return apr::g_cuda_err; }

int apr_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  APR_CUDA_CHECK(cudaGetDevice(&dev));
  int n = 0, maj = 0, min = 0;
  APR_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  APR_CUDA_CHECK(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  APR_CUDA_CHECK(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return APR_OK;
}

}  // extern "C"
