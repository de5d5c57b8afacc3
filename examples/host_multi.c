/* One host process, every GPU of the box, one host thread per GPU -- the shape a single Rust process takes
 * (the reference is one process: gadgets/ecdsa.rs:91-125 builds and proves in place).
 *   gcc -std=c11 -pthread -I include examples/host_multi.c -L plonky2_ecdsa_b200 -lp2e -Wl,-rpath,$PWD/plonky2_ecdsa_b200 -o host_multi
 * Each thread selects its device (p2e_set_device), runs its contiguous shard [g N/G, (g+1) N/G) of ONE batch of ECDSA
 * circuit witnesses through the host-pointer entry point into its region of the shared result matrix, and the main
 * thread checks that every shard equals what device 0 computes for the same signatures.  A second phase has two
 * threads hammer the SAME device with host-pointer calls (they are serialised by the library, never corrupted).
 * Exit code 2: no device.  With one GPU the shards all run on device 0 (still from G threads). */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "p2e.h"

enum { N = 96, MAXG = 8 };
static uint64_t *in[5], *out_all, *out_ref;
static uint32_t *err_all, *err_ref;
static size_t slots;
static int n_dev, n_shards;

struct job { int g; int rc; };

static void* shard_thread(void* arg) {
  struct job* j = (struct job*)arg;
  const int dev = j->g % n_dev;
  if ((j->rc = p2e_set_device(dev)) != 0) return NULL;
  const size_t lo = (size_t)N * j->g / n_shards, hi = (size_t)N * (j->g + 1) / n_shards, m = hi - lo;
  /* the entry point takes [9][m] matrices: copy the shard's columns out of the [9][N] batch, run, copy back */
  uint64_t* sin[5];
  for (int k = 0; k < 5; k++) {
    sin[k] = malloc(9 * m * 8);
    for (int l = 0; l < 9; l++) memcpy(sin[k] + l * m, in[k] + (size_t)l * N + lo, m * 8);
  }
  uint64_t* so = malloc(slots * m * 8);
  uint32_t* se = malloc(m * 4);
  j->rc = p2e_ecdsa_witness(m, sin[0], sin[1], sin[2], sin[3], sin[4], so, se);
  if (j->rc == 0) {
    for (size_t s = 0; s < slots; s++) memcpy(out_all + s * N + lo, so + s * m, m * 8);
    memcpy(err_all + lo, se, m * 4);
  } else {
    fprintf(stderr, "shard %d on device %d: %s\n", j->g, dev, p2e_last_error());
  }
  for (int k = 0; k < 5; k++) free(sin[k]);
  free(so);
  free(se);
  return NULL;
}

static void* hammer_thread(void* arg) {
  struct job* j = (struct job*)arg;
  j->rc = p2e_set_device(0);
  uint64_t* o = malloc(slots * N * 8);
  uint32_t* e = malloc(N * 4);
  for (int rep = 0; rep < 3 && j->rc == 0; rep++) {
    j->rc = p2e_ecdsa_witness(N, in[0], in[1], in[2], in[3], in[4], o, e);
    if (j->rc == 0 && (memcmp(o, out_ref, slots * N * 8) || memcmp(e, err_ref, N * 4))) j->rc = -99;
  }
  free(o);
  free(e);
  return NULL;
}

int main(void) {
  n_dev = p2e_device_count();
  if (n_dev <= 0 || p2e_init_devices(n_dev > MAXG ? MAXG : n_dev) != 0) {
    fprintf(stderr, "no device: %s\n", p2e_last_error());
    return 2;
  }
  if (n_dev > MAXG) n_dev = MAXG;
  n_shards = n_dev > 1 ? n_dev : 3;
  slots = p2e_ecdsa_tape_slots();
  /* inputs: any canonical limbs do (a wrong signature still has a witness for every generator; the final
   * connect fails -> P2E_ERR_CONNECT, identically on every device) */
  uint64_t s = 0x5EC9256B1ull;
  for (int k = 0; k < 5; k++) {
    in[k] = calloc(9 * N, 8);
    for (int l = 0; l < 9; l++)
      for (int i = 0; i < N; i++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        in[k][l * N + i] = (s >> 35) & (l == 8 ? 0x7FFFFF : 0x1FFFFFFF);
      }
  }
  /* pk must be on the curve for curve_assert_valid; use the generator for every row */
  static const uint64_t GX[9] = {0x16f81798, 0x0f940ad8, 0x138a3656, 0x17f9b65b, 0x10b07029, 0x114ae743, 0x0eb15681, 0x0fdf3b97, 0x79be66};
  static const uint64_t GY[9] = {0x1b10d4b8, 0x023e847f, 0x01550667, 0x0f68914d, 0x108a8fd1, 0x1dfe0708, 0x11957693, 0x0ee4d478, 0x483ada};
  for (int l = 0; l < 9; l++)
    for (int i = 0; i < N; i++) { in[3][l * N + i] = GX[l]; in[4][l * N + i] = GY[l]; }
  out_all = calloc(slots * N, 8);
  out_ref = calloc(slots * N, 8);
  err_all = calloc(N, 4);
  err_ref = calloc(N, 4);
  if (p2e_set_device(0) != 0 || p2e_ecdsa_witness(N, in[0], in[1], in[2], in[3], in[4], out_ref, err_ref) != 0) {
    fprintf(stderr, "%s\n", p2e_last_error());
    return 1;
  }
  pthread_t th[MAXG];
  struct job jobs[MAXG];
  for (int g = 0; g < n_shards; g++) { jobs[g].g = g; jobs[g].rc = -1; pthread_create(&th[g], NULL, shard_thread, &jobs[g]); }
  for (int g = 0; g < n_shards; g++) pthread_join(th[g], NULL);
  for (int g = 0; g < n_shards; g++) if (jobs[g].rc != 0) return 1;
  if (memcmp(out_all, out_ref, slots * N * 8) || memcmp(err_all, err_ref, N * 4)) {
    fprintf(stderr, "sharded result differs from the single-device result\n");
    return 1;
  }
  printf("%d shards on %d device(s) from one process: %d witnesses identical to device 0\n", n_shards, n_dev, N);
  for (int g = 0; g < 2; g++) { jobs[g].g = g; jobs[g].rc = -1; pthread_create(&th[g], NULL, hammer_thread, &jobs[g]); }
  for (int g = 0; g < 2; g++) pthread_join(th[g], NULL);
  for (int g = 0; g < 2; g++) if (jobs[g].rc != 0) { fprintf(stderr, "concurrent host calls: rc %d\n", jobs[g].rc); return 1; }
  printf("2 threads x 3 concurrent host-pointer calls on device 0: all identical\n");
  p2e_shutdown();
  return 0;
}
