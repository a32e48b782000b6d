// Per-device host state of libapr_b200: side streams, fork/join events, the cache of executable CUDA graphs of the
// step, the epoch counter of the cross-rank barrier, the event pair of the evaluation timing hook and the cache of the
// item-operand images of the tensor-core evaluation.
//
// One DeviceContext exists per CUDA device, created on the first call made while that device is current
// (apr_context_create does it explicitly) and keyed by the device ordinal, so that a process driving several devices --
// or several host threads -- never shares streams or graphs across devices.  `mu` serialises the entry points that
// mutate the context (train / sharded-train / tensor-core evaluation): one call at a time per device, as the
// reference's single sess.run thread (SURVEY 8b "single host thread per GPU").
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <vector>

namespace apr {

struct GraphEntry {
  std::vector<unsigned char> key;   // bytes of the static StepCtx this executable graph was built / updated for
  const void* fn = nullptr;         // fast-kernel instantiation (identifies the template = the graph's kernels)
  int n_steps = 0;                  // steps per launch of this graph
  int topo = 0;                     // adver | pairs << 1: node set of the captured step
  cudaGraphExec_t exec = nullptr;
  cudaEvent_t done = nullptr;       // recorded after every launch: destroy / update only when it has completed
  unsigned long long last_use = 0;
};

struct DeviceContext {
  int device = -1;
  std::recursive_mutex mu;
  bool ok = false;
  // mode 0 of the training step
  cudaStream_t capture_stream = nullptr;  // origin stream of the stream captures
  cudaStream_t fast_stream = nullptr;     // fast-path kernels
  cudaStream_t pair_stream = nullptr;     // pair work units
  cudaStream_t prep_stream = nullptr;     // index preparation, pipelined one sub-chunk ahead of the step kernels
  // "general path first" variant (narrow rows: the chain of general stages, not the fast kernel, is the critical path)
  cudaStream_t gen_stream = nullptr;      // general stages at the HIGHEST priority
  cudaStream_t fast_lo_stream = nullptr, pair_lo_stream = nullptr;   // fast / pair kernels at the lowest
  cudaEvent_t fork = nullptr, join = nullptr, join2 = nullptr, join3 = nullptr, entry = nullptr;
  std::vector<cudaEvent_t> prep_done;
  std::vector<GraphEntry> graphs;
  unsigned long long graph_clock = 0;
  unsigned long long graph_instantiations = 0, graph_updates = 0, graph_launches = 0;   // apr_context_stats
  // row-sharded training
  int xbarrier_epoch = 0;
  // tensor-core evaluation
  bool tc_timing = false, tc_ev_valid = false;
  cudaEvent_t tc_ev[2] = {nullptr, nullptr};
  // item-operand image cache of apr_eval_fullrank_tc: valid while (Q pointer, range, d, version) match
  const void* qimg_Q = nullptr;
  void* qimg_ws = nullptr;
  int qimg_lo = 0, qimg_hi = 0, qimg_d = 0;
  unsigned long long qimg_version = 0;
  bool qimg_valid = false;

  cudaEvent_t prep_event(size_t k) {
    while (prep_done.size() <= k) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      prep_done.push_back(e);
    }
    return prep_done[k];
  }
};

// context of the CURRENT device (created on first use); nullptr if its streams / events cannot be created
DeviceContext* device_context();

}  // namespace apr
