/* A plain C host of the engine (no Python, no torch): what a cgo / FFI binding does, in 60 lines.
 *   gcc -std=c11 -I include examples/host_demo.c -L plonky2_ecdsa_b200 -lp2e -Wl,-rpath,$PWD/plonky2_ecdsa_b200 -o host_demo
 * Without a CUDA device it reports the library's error and exits 2 (there is no CPU fallback); with one it runs a batch
 * of MulNonnativeGenerator witnesses (gates/mul_nonnative.rs:249-324) and checks x*y = q*p + r limb-wise mod 2^29.
 * It also exercises the two host-only parts of the ABI: tape layouts and the serialized-generator loader. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "p2e.h"

static void put_u32(uint8_t** p, uint32_t v) { memcpy(*p, &v, 4); *p += 4; }
static void put_u64(uint8_t** p, uint64_t v) { memcpy(*p, &v, 8); *p += 8; }

int main(void) {
  /* host-only: the ECDSA tape layout (gadgets/ecdsa.rs:30-53) */
  size_t slots = p2e_ecdsa_tape_slots();
  size_t gens = p2e_ecdsa_tape_layout(NULL, NULL);
  printf("verify_secp256k1_message_circuit: %zu generator instances, %zu witness slots\n", gens, slots);
  if (slots != 82615 || gens != 4642) return 1;

  /* host-only: a serialized MulNonnativeGenerator + CheckSumGenerator pair with their copy constraints */
  uint8_t buf[4096], *p = buf;
  put_u32(&p, P2E_GEN_MUL_NONNATIVE_BASE); put_u64(&p, 0); put_u64(&p, 9);      /* row 0, num_limbs 9 */
  put_u32(&p, P2E_GEN_CHECK_SUM); put_u64(&p, 1); put_u64(&p, 9);               /* row 1 */
  for (uint64_t k = 0; k < 17; k++) {                                           /* connect(wire(0, 36 + k), wire(1, k)) */
    put_u32(&p, P2E_GEN_COPY);
    *p++ = 1; put_u64(&p, 0); put_u64(&p, 36 + k);
    *p++ = 1; put_u64(&p, 1); put_u64(&p, k);
  }
  int h = p2e_genstream_load(buf, (size_t)(p - buf));
  if (h < 0) { fprintf(stderr, "genstream: %s\n", p2e_last_error()); return 1; }
  printf("generator stream: %ld generators, %ld levels, %ld slots, %ld input slots\n", p2e_genstream_info(h, 1),
         p2e_genstream_info(h, 2), p2e_genstream_info(h, 0), p2e_genstream_info(h, 4));
  if (p2e_genstream_info(h, 2) != 2 || p2e_genstream_info(h, 4) != 18) return 1;
  p2e_genstream_free(h);

  if (p2e_init(0) != 0) {
    fprintf(stderr, "no device: %s\n", p2e_last_error());
    return 2;
  }
  enum { N = 1000 };
  uint64_t *x = calloc(9 * N, 8), *y = calloc(9 * N, 8), *out = calloc((size_t)P2E_MUL_OUT_ROWS * N, 8);
  uint32_t* err = calloc(N, 4);
  uint64_t s = 0x5EC9256B1ull;
  for (int l = 0; l < 9; l++)
    for (int i = 0; i < N; i++) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      x[l * N + i] = (s >> 35) & (l == 8 ? 0xFFFFFF : 0x1FFFFFFF);
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      y[l * N + i] = (s >> 35) & (l == 8 ? 0xFFFFFF : 0x1FFFFFFF);
    }
  if (p2e_mul_nonnative_witness(P2E_SECP256K1_BASE, N, x, y, out, err) != 0) {
    fprintf(stderr, "%s\n", p2e_last_error());
    return 1;
  }
  /* limb 0 of x*y - q*p - r is a multiple of 2^29 (p mod 2^29 = 0x1ffffc2f) */
  for (int i = 0; i < N; i++) {
    uint64_t v = x[i] * y[i] - out[(P2E_MUL_OUT_Q)*N + i] * 0x1ffffc2full - out[(P2E_MUL_OUT_R)*N + i];
    if (err[i] || (v & 0x1FFFFFFF)) { fprintf(stderr, "row %d inconsistent\n", i); return 1; }
  }
  printf("%d mulmod witnesses ok\n", N);
  p2e_shutdown();
  return 0;
}
