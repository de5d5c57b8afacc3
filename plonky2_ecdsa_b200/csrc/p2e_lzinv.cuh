// p2e_lzinv.cuh -- base-field inversion x^(p-2) in lazy limbs with out-of-line multiply / square bodies, for kernels
// that are bound by the latency of ONE ladder per thread (standalone inverse witness, generator-stream inverse groups).
// Every translation unit that includes this header gets its own copy of the two bodies.
#pragma once
#include "p2e_field.cuh"

namespace p2e {

// Base-field Fermat ladder in lazy limbs through out-of-line multiply / square bodies: inlined, the ladder is ~160 KB
// of straight-line code and the (latency-bound) kernel waits on instruction fetch (ncu: no_instruction 1.2 per issue).
struct L9 {
  i32 v[NL];
};
static __device__ __noinline__ L9 gs_lzm(L9 a, L9 b) {
  L9 r;
  lz_mul(a.v, b.v, r.v);
  return r;
}
static __device__ __noinline__ L9 gs_lzs(L9 a) {
  L9 r;
  lz_sqr(a.v, r.v);
  return r;
}
__device__ __forceinline__ L9 gs_lzs_n(L9 t, int n) {
  for (int k = 0; k < n; k++) t = gs_lzs(t);
  return t;
}
// x^(p-2), addition chain of p2e_field.cuh::lz_inv
__device__ __forceinline__ L9 gs_lz_inv(L9 a) {
  const L9 x2 = gs_lzm(gs_lzs(a), a);
  const L9 x3 = gs_lzm(gs_lzs(x2), a);
  L9 u = gs_lzm(gs_lzs_n(x3, 3), x3);                 // x6
  u = gs_lzm(gs_lzs_n(u, 3), x3);                     // x9
  u = gs_lzm(gs_lzs_n(u, 2), x2);                     // x11
  const L9 x22 = gs_lzm(gs_lzs_n(u, 11), u);
  const L9 x44 = gs_lzm(gs_lzs_n(x22, 22), x22);
  const L9 x88 = gs_lzm(gs_lzs_n(x44, 44), x44);
  u = gs_lzm(gs_lzs_n(x88, 88), x88);                 // x176
  u = gs_lzm(gs_lzs_n(u, 44), x44);                   // x220
  u = gs_lzm(gs_lzs_n(u, 3), x3);                     // x223
  u = gs_lzm(gs_lzs_n(u, 23), x22);
  u = gs_lzm(gs_lzs_n(u, 5), a);
  u = gs_lzm(gs_lzs_n(u, 3), x2);
  return gs_lzm(gs_lzs_n(u, 2), a);
}

// canonical in, canonical out (the inverse is unique, so this is bit-identical to any other way of computing it)
__device__ __forceinline__ void inv_p_lazy(const u32 (&a)[NL], u32 (&r)[NL]) {
  L9 lx;
  lz_from_canon(a, lx.v);
  const L9 li = gs_lz_inv(lx);
  lz_canon(li.v, r);
}

}  // namespace p2e
