// p2e_f8inv.cuh -- base-field inversion x^(p-2) on the saturated 8 x 32-bit field (p2e_f8.cuh) with out-of-line multiply
// / square bodies, for kernels that are bound by the latency of ONE ladder per thread (chain K2, standalone inverse
// witness, generator-stream inverse groups).  Every translation unit that includes this header gets its own copy of
// the two bodies.  The inverse is unique, so this equals the reference's Fermat exp_biguint(p - 2) bit for bit.
#pragma once
#include "p2e_f8.cuh"

namespace p2e {

// Operands and results cross the call boundary BY VALUE (structs of 8 words): ptxas keeps them in registers.  Inlined,
// the ladder is ~100 KB of straight-line code and a latency-bound kernel waits on instruction fetch.
static __device__ __noinline__ F8 f8m_v(F8 a, F8 b) {
  F8 r;
  f8_mul(a.v, b.v, r.v);
  return r;
}
static __device__ __noinline__ F8 f8s_v(F8 a) {
  F8 r;
  f8_sqr(a.v, r.v);
  return r;
}
__device__ __forceinline__ F8 f8s_n(F8 t, int n) {
  for (int k = 0; k < n; k++) t = f8s_v(t);
  return t;
}
// x^(p-2): 255 squarings + 15 multiplications (addition chain over the run structure of p - 2 = 2^256 - 2^32 - 979)
static __device__ F8 f8_inv_calls(F8 a) {
  const F8 x2 = f8m_v(f8s_v(a), a);
  const F8 x3 = f8m_v(f8s_v(x2), a);
  F8 u = f8m_v(f8s_n(x3, 3), x3);                  // x6
  u = f8m_v(f8s_n(u, 3), x3);                      // x9
  u = f8m_v(f8s_n(u, 2), x2);                      // x11
  const F8 x22 = f8m_v(f8s_n(u, 11), u);
  const F8 x44 = f8m_v(f8s_n(x22, 22), x22);
  const F8 x88 = f8m_v(f8s_n(x44, 44), x44);
  u = f8m_v(f8s_n(x88, 88), x88);                  // x176
  u = f8m_v(f8s_n(u, 44), x44);                    // x220
  u = f8m_v(f8s_n(u, 3), x3);                      // x223
  u = f8m_v(f8s_n(u, 23), x22);
  u = f8m_v(f8s_n(u, 5), a);
  u = f8m_v(f8s_n(u, 3), x2);
  return f8m_v(f8s_n(u, 2), a);
}

// canonical 29-bit limbs in, canonical 29-bit limbs out
__device__ __forceinline__ void inv_p_f8(const u32 (&a)[NL], u32 (&r)[NL]) {
  F8 x;
  f8_from_l29(a, x.v);
  const F8 y = f8_inv_calls(x);
  f8_to_canon_l29(y.v, r);
}

}  // namespace p2e
