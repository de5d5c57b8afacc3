// p2e_common.h -- process-wide context (device, stream, scratch) shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace p2e {

struct Context {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  // grow-only device scratch used by the host-pointer entry points
  void* scratch[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t scratch_bytes[4] = {0, 0, 0, 0};
  std::string last_error;
};

Context& ctx();
int set_error(const char* where, cudaError_t e);
int set_error_msg(const char* msg);
int ensure_ready();
// returns a device buffer of at least `bytes` in scratch slot `slot` (kept until shutdown)
int get_scratch(int slot, size_t bytes, void** out);

#define P2E_CUDA(call)                                  \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return p2e::set_error(#call, e__); \
  } while (0)

void chain_release_device_tapes();
void ntt_release_tables();
void genstream_release_device();

static inline cudaStream_t pick_stream(void* s) { return s ? (cudaStream_t)s : ctx().stream; }

}  // namespace p2e
