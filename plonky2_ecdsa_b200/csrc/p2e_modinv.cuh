// p2e_modinv.cuh -- modular inverse by the Bernstein-Yang division steps ("safegcd"), constant trip count: 20 rounds
// of 30 division steps on the low words, each round applied to the full 256-bit state as one 2 x 2 transition matrix
// (signed 30-bit limbs, 64-bit accumulators -> IMAD.WIDE).  ~20 k instructions per inverse against ~54 k for the
// Fermat ladder on the saturated base field and ~110 k for the Montgomery ladder modulo the group order, no data-
// dependent branch (all lanes of a warp stay together), and a third of the dependent-latency of the ladder -- the
// one-thread-per-signature kernels of the ECDSA chain (K0's s^-1 mod n, K2's batched inversion) are latency bound.
//
// The reference inverts by exp_biguint(x, p - 2) (nonnative.rs:863); the inverse of x != 0 modulo a prime is unique and
// 0 maps to 0 in both algorithms, so the witness bits are the same (tests/test_hostcheck.py compares the two bodies,
// tests/test_gpu_* check it on the device).  Included by p2e_field.cuh (needs ModP / ModN and the integer typedefs).
#pragma once

namespace p2e {

typedef int32_t mi32;

template <class MOD>
struct ModInv30;
template <>
struct ModInv30<ModP> {   // p in 30-bit limbs, p^-1 mod 2^30
  P2E_HD static constexpr mi32 m(int i) {
    constexpr mi32 t[9] = {0x3ffffc2f, 0x3ffffffb, 0x3fffffff, 0x3fffffff, 0x3fffffff, 0x3fffffff, 0x3fffffff, 0x3fffffff, 0xffff};
    return t[i];
  }
  static constexpr u32 INV30 = 0x2ddacacfu;
};
template <>
struct ModInv30<ModN> {
  P2E_HD static constexpr mi32 m(int i) {
    constexpr mi32 t[9] = {0x10364141, 0x3f497a33, 0x348a03bb, 0x2bb739ab, 0x3ffffeba, 0x3fffffff, 0x3fffffff, 0x3fffffff, 0xffff};
    return t[i];
  }
  static constexpr u32 INV30 = 0x2a774ec1u;
};

constexpr mi32 MI_M30 = 0x3fffffff;

// 30 division steps on the low 30 bits of f (odd) and g; zeta = -(delta + 1/2).  Returns the new zeta and the matrix
// t = [u v; q r] with t * [f, g] = 2^30 * [f', g'].
P2E_HD mi32 mi_divsteps30(mi32 zeta, u32 f, u32 g, mi32& tu, mi32& tv, mi32& tq, mi32& tr) {
  u32 u = 1, v = 0, q = 0, r = 1;
#ifndef MI_UNROLL
#define MI_UNROLL 6
#endif
  constexpr int mi_unroll = MI_UNROLL;
#pragma unroll mi_unroll
  for (int i = 0; i < 30; i++) {
    u32 c1 = (u32)(zeta >> 31);            // zeta < 0
    const u32 c2 = 0u - (g & 1u);          // g odd
    const u32 x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
    g += x & c2;
    q += y & c2;
    r += z & c2;
    c1 &= c2;
    zeta = (mi32)((u32)zeta ^ c1) - 1;
    f += g & c1;
    u += q & c1;
    v += r & c1;
    g >>= 1;
    u <<= 1;
    v <<= 1;
  }
  tu = (mi32)u;
  tv = (mi32)v;
  tq = (mi32)q;
  tr = (mi32)r;
  return zeta;
}

// [f, g] <- t * [f, g] / 2^30 (exact)
P2E_HD void mi_update_fg(mi32 (&f)[9], mi32 (&g)[9], mi32 u, mi32 v, mi32 q, mi32 r) {
  i64 cf = (i64)u * f[0] + (i64)v * g[0];
  i64 cg = (i64)q * f[0] + (i64)r * g[0];
  cf >>= 30;
  cg >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    cf += (i64)u * f[i] + (i64)v * g[i];
    cg += (i64)q * f[i] + (i64)r * g[i];
    f[i - 1] = (mi32)cf & MI_M30;
    g[i - 1] = (mi32)cg & MI_M30;
    cf >>= 30;
    cg >>= 30;
  }
  f[8] = (mi32)cf;
  g[8] = (mi32)cg;
}

// [d, e] <- t * [d, e] / 2^30 (mod m), both kept in (-2m, m)
template <class MOD>
P2E_HD void mi_update_de(mi32 (&d)[9], mi32 (&e)[9], mi32 u, mi32 v, mi32 q, mi32 r) {
  typedef ModInv30<MOD> MI;
  const mi32 sd = d[8] >> 31, se = e[8] >> 31;
  mi32 md = (u & sd) + (v & se), me = (q & sd) + (r & se);
  i64 cd = (i64)u * d[0] + (i64)v * e[0];
  i64 ce = (i64)q * d[0] + (i64)r * e[0];
  // the multiples of m that clear the low 30 bits
  md -= (mi32)((MI::INV30 * (u32)cd + (u32)md) & (u32)MI_M30);
  me -= (mi32)((MI::INV30 * (u32)ce + (u32)me) & (u32)MI_M30);
  cd += (i64)MI::m(0) * md;
  ce += (i64)MI::m(0) * me;
  cd >>= 30;
  ce >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    cd += (i64)u * d[i] + (i64)v * e[i];
    ce += (i64)q * d[i] + (i64)r * e[i];
    cd += (i64)MI::m(i) * md;
    ce += (i64)MI::m(i) * me;
    d[i - 1] = (mi32)cd & MI_M30;
    e[i - 1] = (mi32)ce & MI_M30;
    cd >>= 30;
    ce >>= 30;
  }
  d[8] = (mi32)cd;
  e[8] = (mi32)ce;
}

// d in (-2m, m), negated when sign < 0, brought to [0, m)
template <class MOD>
P2E_HD void mi_normalize(mi32 (&d)[9], mi32 sign) {
  typedef ModInv30<MOD> MI;
  mi32 add = d[8] >> 31;
  const mi32 neg = sign >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    d[i] += MI::m(i) & add;
    d[i] = (d[i] ^ neg) - neg;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    d[i + 1] += d[i] >> 30;
    d[i] &= MI_M30;
  }
  add = d[8] >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) d[i] += MI::m(i) & add;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    d[i + 1] += d[i] >> 30;
    d[i] &= MI_M30;
  }
}

// x: 8 words, x < m.  r = x^-1 mod m in [0, m) (0 for x = 0).
template <class MOD>
P2E_HD void modinv_words(const u32 (&x)[8], u32 (&out)[8]) {
  typedef ModInv30<MOD> MI;
  mi32 f[9], g[9], d[9], e[9];
#pragma unroll
  for (int i = 0; i < 9; i++) {
    f[i] = MI::m(i);
    d[i] = 0;
    e[i] = (i == 0);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int k = (30 * i) >> 5, sh = (30 * i) & 31;
    u32 w = x[k] >> sh;
    if (sh > 2) w |= x[k + 1] << (32 - sh);
    g[i] = (mi32)(w & (u32)MI_M30);
  }
  g[8] = (mi32)(x[7] >> 16);
  mi32 zeta = -1;
#pragma unroll 1
  for (int it = 0; it < 20; it++) {
    mi32 u, v, q, r;
    zeta = mi_divsteps30(zeta, (u32)f[0], (u32)g[0], u, v, q, r);
    mi_update_de<MOD>(d, e, u, v, q, r);
    mi_update_fg(f, g, u, v, q, r);
  }
  // g = 0, f = +-gcd = +-1 (x != 0): x^-1 = d * f
  mi_normalize<MOD>(d, f[8]);
#pragma unroll
  for (int k = 0; k < 8; k++) out[k] = ((u32)d[k] >> (2 * k)) | ((u32)d[k + 1] << (30 - 2 * k));
}

}  // namespace p2e
