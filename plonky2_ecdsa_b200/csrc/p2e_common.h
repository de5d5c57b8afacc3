// p2e_common.h -- per-device contexts (device, stream, scratch, per-subsystem state) shared by the C-ABI translation
// units.  One process may drive several GPUs: p2e_init(device) creates a context per device and p2e_set_device()
// selects the calling THREAD's current one (like cudaSetDevice); every entry point works on that context.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include <string>

namespace p2e {

constexpr int MAX_DEVICES = 16;

struct Context {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  // grow-only device scratch used by the host-pointer entry points; `host_mu` serialises those entry points per
  // device (they share the scratch slots and the context's stream).  The *_dev entry points do not take it.
  void* scratch[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t scratch_bytes[4] = {0, 0, 0, 0};
  std::recursive_mutex host_mu;
  // per-device state of the subsystems (owned by their translation units, released through the hooks below)
  void* chain = nullptr;
  void* ntt = nullptr;
  void* genstream = nullptr;
};

Context& ctx();                       // context of the calling thread's current device (a never-ready dummy if none)
int set_error(const char* where, cudaError_t e);
int set_error_msg(const char* msg);
int ensure_ready();                   // also makes the context's device the CUDA current device of this thread
// returns a device buffer of at least `bytes` in scratch slot `slot` (kept until shutdown)
int get_scratch(int slot, size_t bytes, void** out);
// tunables that used to be environment variables (p2e_set_option / p2e_get_option)
long option(int id);
enum OptionId : int { OPT_CHAIN_CHUNK = 0, OPT_HOST_CHUNKS = 1, OPT_GENSTREAM_FUSED = 2, OPT_GENSTREAM_GRAPH = 3, OPT_NTT_LOG_TILE = 4, OPT_CHAIN_TRACE = 5, OPT_CHAIN_SPLIT = 6, OPT_CHAIN_PHASES = 7, OPT_GENSTREAM_WALK = 8, OPT_GENSTREAM_WALK_SHAPE = 9, OPT_CHAIN_REPLAY_LANES = 10, OPT_CHAIN_GRAPH = 11, N_OPTIONS = 12 };

#define P2E_CUDA(call)                                  \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return p2e::set_error(#call, e__); \
  } while (0)

// host-pointer entry points: one at a time per device
#define P2E_HOST_LOCK() std::lock_guard<std::recursive_mutex> host_lock__(p2e::ctx().host_mu)

void chain_release_device(Context& c);
void ntt_release_device(Context& c);
void genstream_release_device(Context& c);

static inline cudaStream_t pick_stream(void* s) { return s ? (cudaStream_t)s : ctx().stream; }

}  // namespace p2e
