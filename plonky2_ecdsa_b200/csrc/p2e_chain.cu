// p2e_chain.cu -- ECDSA / glv_mul witness tapes (placeholder until the chain kernels land).
#include "../../include/p2e.h"
#include "p2e_common.h"

using namespace p2e;
extern "C" {
size_t p2e_ecdsa_tape_slots(void) { return 0; }
size_t p2e_ecdsa_tape_layout(uint32_t*, uint32_t*) { return 0; }
int p2e_ecdsa_witness_dev(size_t, const uint64_t*, const uint64_t*, const uint64_t*, const uint64_t*, const uint64_t*,
                          size_t, uint64_t*, size_t, uint32_t*, void*) {
  return set_error_msg("p2e_ecdsa_witness: chain kernels not built in this revision");
}
int p2e_ecdsa_witness(size_t, const uint64_t*, const uint64_t*, const uint64_t*, const uint64_t*, const uint64_t*,
                      uint64_t*, uint32_t*) {
  return set_error_msg("p2e_ecdsa_witness: chain kernels not built in this revision");
}
size_t p2e_glv_mul_tape_slots(void) { return 0; }
size_t p2e_glv_mul_tape_layout(uint32_t*, uint32_t*) { return 0; }
int p2e_glv_mul_witness_dev(size_t, const uint64_t*, const uint64_t*, const uint64_t*, size_t, uint64_t*, size_t,
                            uint32_t*, void*) {
  return set_error_msg("p2e_glv_mul_witness: chain kernels not built in this revision");
}
int p2e_glv_mul_witness(size_t, const uint64_t*, const uint64_t*, const uint64_t*, uint64_t*, uint32_t*) {
  return set_error_msg("p2e_glv_mul_witness: chain kernels not built in this revision");
}
}
