// Library-level entry points of libapr_b200: version, status strings, device info.
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "context.cuh"

namespace apr {

static thread_local char g_cuda_err[512] = "";

// ------------------------------------------------------------------------------------------------
// per-device contexts (context.cuh)
// ------------------------------------------------------------------------------------------------
constexpr int kMaxDevices = 64;
static std::mutex g_ctx_mu;
static DeviceContext* g_ctx[kMaxDevices] = {nullptr};

static bool context_init(DeviceContext& a) {
  // index preparation runs at the LOWEST priority (it only has to stay one sub-chunk ahead: let it fill the gaps the
  // step kernels leave), the fast-path kernels at the highest
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  return cudaStreamCreateWithPriority(&a.fast_stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
         cudaStreamCreateWithPriority(&a.prep_stream, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
         cudaStreamCreateWithPriority(&a.pair_stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
         cudaStreamCreateWithFlags(&a.capture_stream, cudaStreamNonBlocking) == cudaSuccess &&
         cudaStreamCreateWithPriority(&a.gen_stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
         cudaStreamCreateWithPriority(&a.fast_lo_stream, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
         cudaStreamCreateWithPriority(&a.pair_lo_stream, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
         cudaEventCreateWithFlags(&a.join3, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&a.join2, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&a.entry, cudaEventDisableTiming) == cudaSuccess;
}

static void context_free(DeviceContext* a) {
  for (auto& g : a->graphs) {
    if (g.done) { cudaEventSynchronize(g.done); cudaEventDestroy(g.done); }
    if (g.exec) cudaGraphExecDestroy(g.exec);
  }
  for (auto e : a->prep_done) cudaEventDestroy(e);
  for (cudaStream_t s : {a->capture_stream, a->fast_stream, a->pair_stream, a->prep_stream, a->gen_stream, a->fast_lo_stream,
                         a->pair_lo_stream})
    if (s) cudaStreamDestroy(s);
  for (cudaEvent_t e : {a->fork, a->join, a->join2, a->join3, a->entry, a->tc_ev[0], a->tc_ev[1]}) if (e) cudaEventDestroy(e);
  delete a;
}

DeviceContext* device_context() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  DeviceContext* c = g_ctx[dev];
  if (!c) {
    c = new DeviceContext();
    c->device = dev;
    c->ok = context_init(*c);
    if (!c->ok) { set_cuda_error(cudaGetLastError(), "apr device context"); context_free(c); return nullptr; }
    g_ctx[dev] = c;
  }
  return c;
}

void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

int sm_count() {
  static int cached[kMaxDevices] = {0};
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  if (cached[dev]) return cached[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached[dev] = n;
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

int apr_context_create(int32_t device) {
  int prev = -1;
  APR_CUDA_CHECK(cudaGetDevice(&prev));
  if (device < 0 || device >= apr::kMaxDevices) return APR_E_ARG;
  APR_CUDA_CHECK(cudaSetDevice(device));
  apr::DeviceContext* c = apr::device_context();
  cudaSetDevice(prev);
  return c ? APR_OK : APR_E_CUDA;
}

int apr_context_destroy(int32_t device) {
  if (device < 0 || device >= apr::kMaxDevices) return APR_E_ARG;
  apr::DeviceContext* c = nullptr;
  {
    std::lock_guard<std::mutex> lk(apr::g_ctx_mu);
    c = apr::g_ctx[device];
    apr::g_ctx[device] = nullptr;
  }
  if (!c) return APR_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  { std::lock_guard<std::recursive_mutex> lk(c->mu); }   // wait for a call in flight on another thread
  apr::context_free(c);
  if (prev >= 0) cudaSetDevice(prev);
  return APR_OK;
}

int apr_context_stats(int64_t* out4_host) {
  if (!out4_host) return APR_E_ARG;
  apr::DeviceContext* c = apr::device_context();
  if (!c) return APR_E_CUDA;
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  out4_host[0] = int64_t(c->graph_instantiations);
  out4_host[1] = int64_t(c->graph_updates);
  out4_host[2] = int64_t(c->graph_launches);
  out4_host[3] = int64_t(c->graphs.size());
  return APR_OK;
}

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
