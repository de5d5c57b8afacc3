// p2e_field.cuh -- 9x29-bit limb arithmetic for the secp256k1 base/scalar fields, Goldilocks
// helpers, and the per-item witness-generator cores (one item per thread, everything in
// registers).  All functions are __host__ __device__ so that the same code builds the circuit
// constants on the host (p2e_tape.cpp) and can be unit-checked on a CPU-only box.
//
// Reference semantics restated here (file:line in Weobe/plonky2-ecdsa):
//   MulNonnativeGenerator::run_once      src/gates/mul_nonnative.rs:249-324
//   CheckSumGenerator::run_once          src/gates/mul_nonnative.rs:513-531
//   NonNativeAddition/MultipleAdds/Subtraction/Inverse generators
//                                        src/gadgets/nonnative.rs:626-645, 696-728, 792-810, 857-872
//   get/set_biguint_target               src/gadgets/biguint.rs:444-480
//   decompose_secp256k1_scalar           src/curve/glv.rs:39-77
#pragma once
#include <stdint.h>

namespace p2e {

typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

#define P2E_HD __host__ __device__ __forceinline__

constexpr int NL = 9;                 // limbs per nonnative value (mul_nonnative.rs:37-39)
constexpr int LB = 29;                // bits per limb (nonnative.rs:32)
constexpr u32 LMASK = 0x1fffffffu;
constexpr u32 TOPMASK = 0x00ffffffu;  // limb 8 holds bits 232..255 of a 256-bit value
constexpr u64 GL_P = 0xFFFFFFFF00000001ull;
constexpr u64 GL_EPS = 0xFFFFFFFFull;                 // 2^64 mod GL_P
constexpr u64 GL_INV_2_29 = 0xfffffff700000009ull;    // (2^29)^-1 mod GL_P

// per-item error bits: the reference panics / asserts in these situations
enum : u32 {
  ERR_U32 = 1,            // "Value doesn't fit in u32"            mul_nonnative.rs:262,271
  ERR_LIMB_RANGE = 2,     // assert limb < 2^29                    mul_nonnative.rs:275-276
  ERR_NONCANONICAL = 4,   // from_noncanonical_biguint >= 2^256    [plonky2 field pattern p256_base.rs:121-130]
  ERR_INVERSE_ZERO = 8,   // inverse() of zero
  ERR_CARRY_RANGE = 16,   // assert b_i + 2^33 < 2^34              mul_nonnative.rs:527
  ERR_DIV_ZERO = 32,      // BigUint div_rem by zero               biguint.rs:515
  ERR_LIMBS_OVERFLOW = 64,// assert target.num_limbs() >= limbs    biguint.rs:473
  ERR_CONNECT = 128,      // a connect_* / assert_* of the gadget is unsatisfiable
  ERR_INTERNAL = 256      // device-side self check failed (never expected)
};

// ------------------------------------------------------------------------------------------
// Moduli.  M = 2^256 - c ; d = 32*c = 2^261 mod M.  nt(k) = limb count of the running value
// before fold k (value bounds derived in DESIGN.md "exact division without a wide divide").
struct ModP {   // secp256k1 base field
  static constexpr int ID = 0;
  static constexpr int NC = 2, ND = 2, NFOLD = 3;
  P2E_HD static constexpr u32 m(int i) {
    constexpr u32 t[9] = {0x1ffffc2fu, 0x1ffffff7u, 0x1fffffffu, 0x1fffffffu, 0x1fffffffu,
                          0x1fffffffu, 0x1fffffffu, 0x1fffffffu, 0x00ffffffu};
    return t[i];
  }
  P2E_HD static constexpr u32 c(int i) { constexpr u32 t[2] = {0x000003d1u, 0x00000008u}; return t[i]; }
  P2E_HD static constexpr u32 d(int i) { constexpr u32 t[2] = {0x00007a20u, 0x00000100u}; return t[i]; }
  P2E_HD static constexpr u32 half(int i) {
    constexpr u32 t[9] = {0x1ffffe17u, 0x1ffffffbu, 0x1fffffffu, 0x1fffffffu, 0x1fffffffu,
                          0x1fffffffu, 0x1fffffffu, 0x1fffffffu, 0x007fffffu};
    return t[i];
  }
  P2E_HD static constexpr int nt(int k) { constexpr int t[4] = {18, 11, 10, 9}; return t[k]; }
};

struct ModN {   // secp256k1 scalar field
  static constexpr int ID = 1;
  static constexpr int NC = 5, ND = 5, NFOLD = 4;
  P2E_HD static constexpr u32 m(int i) {
    constexpr u32 t[9] = {0x10364141u, 0x1e92f466u, 0x12280eefu, 0x1db9cd5eu, 0x1fffebaau,
                          0x1fffffffu, 0x1fffffffu, 0x1fffffffu, 0x00ffffffu};
    return t[i];
  }
  P2E_HD static constexpr u32 c(int i) {
    constexpr u32 t[5] = {0x0fc9bebfu, 0x016d0b99u, 0x0dd7f110u, 0x024632a1u, 0x00001455u};
    return t[i];
  }
  P2E_HD static constexpr u32 d(int i) {
    constexpr u32 t[5] = {0x1937d7e0u, 0x0da1732fu, 0x1afe2201u, 0x08c6542du, 0x00028aa2u};
    return t[i];
  }
  P2E_HD static constexpr u32 half(int i) {   // floor(n / 2)
    constexpr u32 t[9] = {0x081b20a0u, 0x1f497a33u, 0x09140777u, 0x0edce6afu, 0x1ffff5d5u,
                          0x1fffffffu, 0x1fffffffu, 0x1fffffffu, 0x007fffffu};
    return t[i];
  }
  P2E_HD static constexpr int nt(int k) { constexpr int t[5] = {18, 14, 10, 10, 9}; return t[k]; }
};

// ------------------------------------------------------------------------------------------
// Goldilocks
P2E_HD u64 umulhi64(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
  return __umul64hi(a, b);
#else
  return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

P2E_HD u64 gl_add(u64 a, u64 b) {          // a, b canonical
  u64 s = a + b;
  if (s < a) s += GL_EPS;                  // wrapped: true sum - P == s + 2^32 - 1
  if (s >= GL_P) s -= GL_P;
  return s;
}

P2E_HD u64 gl_sub(u64 a, u64 b) {          // a, b canonical
  u64 d = a - b;
  if (a < b) d -= GL_EPS;                  // + P  (mod 2^64)
  return d;
}

P2E_HD u64 gl_reduce128(u64 lo, u64 hi) {  // (hi:lo) mod P, canonical
  u64 hh = hi >> 32, hl = hi & GL_EPS;
  u64 t0 = lo - hh;
  if (lo < hh) t0 -= GL_EPS;
  u64 t1 = hl * GL_EPS;
  u64 r = t0 + t1;
  if (r < t1) r += GL_EPS;
  if (r >= GL_P) r -= GL_P;
  return r;
}

P2E_HD u64 gl_mul(u64 a, u64 b) { return gl_reduce128(a * b, umulhi64(a, b)); }

// canonical Goldilocks representative of a signed integer |v| < 2^63
P2E_HD u64 gl_from_i64(i64 v) { return (u64)v + ((u64)(v >> 63) & (0ull - GL_EPS)); }   // + P when negative (mod 2^64)

#ifdef __CUDACC__
// Lazy Goldilocks for the constraint evaluators: wires are canonical, intermediates are ANY u64
// congruent to the value, every stored constraint is canonicalised once (glc).  The corrections use
// the carry flag (sub.cc / subc, add.cc / addc) instead of 64-bit compares + selects.
//   glm(any, any) -> any        gla(any, canonical) -> any        gls(any, canonical) -> any
__device__ __forceinline__ u64 gl_red_lazy(u64 lo, u64 hi) {   // (hi:lo) mod P, result < 2^64
  const u32 hh = (u32)(hi >> 32), hl = (u32)hi;
  u64 t0, r;
  u32 b, c;
  asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(t0), "=r"(b) : "l"(lo), "l"((u64)hh));
  t0 -= (u64)b;                                  // b = 0xffffffff = 2^64 mod P on borrow: + P
  u64 t1;                                        // hl * (2^32 - 1) as ONE IMAD.WIDE: the evaluators / NTT are bound by the
  asm("mul.wide.u32 %0, %1, 0xffffffff;" : "=l"(t1) : "r"(hl));   // ALU pipe (shift + 64-bit subtract), the IMAD pipe idles
  asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(r), "=r"(c) : "l"(t0), "l"(t1));
  return r + (u64)(0u - c);                      // 2^64 = 2^32 - 1 (mod P) on carry
}
__device__ __forceinline__ u64 glm(u64 a, u64 b) { return gl_red_lazy(a * b, __umul64hi(a, b)); }
__device__ __forceinline__ u64 gla(u64 a, u64 b) {             // b canonical
  u64 r;
  u32 c;
  asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(r), "=r"(c) : "l"(a), "l"(b));
  return r + (u64)(0u - c);
}
__device__ __forceinline__ u64 gls(u64 a, u64 b) {             // b canonical
  u64 r;
  u32 bw;
  asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(r), "=r"(bw) : "l"(a), "l"(b));
  return r - (u64)bw;
}
__device__ __forceinline__ u64 glc(u64 a) {                    // any -> canonical
  u64 t;
  u32 c;
  asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(t), "=r"(c) : "l"(a), "l"(GL_EPS));   // a - P, carry iff a >= P
  return c ? t : a;
}
#endif

// ------------------------------------------------------------------------------------------
// limb helpers (u32 limbs are always < 2^29 unless stated)
P2E_HD void set_zero9(u32 (&a)[NL]) {
#pragma unroll
  for (int i = 0; i < NL; i++) a[i] = 0;
}

P2E_HD bool is_zero9(const u32 (&a)[NL]) {
  u32 o = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) o |= a[i];
  return o == 0;
}

P2E_HD bool eq9(const u32 (&a)[NL], const u32 (&b)[NL]) {
  u32 o = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) o |= a[i] ^ b[i];
  return o == 0;
}

// a >= b for normalised 9-limb values
P2E_HD bool geq9(const u32 (&a)[NL], const u32 (&b)[NL]) {
  // borrow chain of a - b
  int32_t br = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) br = ((int32_t)a[i] - (int32_t)b[i] + br) >> LB;
  return br == 0;
}

template <class MOD>
P2E_HD bool geq_mod(const u32 (&a)[NL]) {
  int32_t br = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) br = ((int32_t)a[i] - (int32_t)MOD::m(i) + br) >> LB;
  return br == 0;
}

// r = a + b (no reduction); returns carry-out limb (value >> 261)
P2E_HD u32 add9(const u32 (&a)[NL], const u32 (&b)[NL], u32 (&r)[NL]) {
  u32 c = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    u32 v = a[i] + b[i] + c;
    r[i] = v & LMASK;
    c = v >> LB;
  }
  return c;
}

// r = a - b for a >= b
P2E_HD void sub9(const u32 (&a)[NL], const u32 (&b)[NL], u32 (&r)[NL]) {
  int32_t br = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    int32_t v = (int32_t)a[i] - (int32_t)b[i] + br;
    r[i] = (u32)v & LMASK;
    br = v >> LB;
  }
}

template <class MOD>
P2E_HD void sub_mod9(const u32 (&a)[NL], u32 (&r)[NL]) {   // r = a - M for a >= M
  int32_t br = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    int32_t v = (int32_t)a[i] - (int32_t)MOD::m(i) + br;
    r[i] = (u32)v & LMASK;
    br = v >> LB;
  }
}

template <class MOD>
P2E_HD void add_mod9(const u32 (&a)[NL], u32 (&r)[NL]) {   // r = a + M (caller guarantees < 2^261)
  u32 c = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    u32 v = a[i] + MOD::m(i) + c;
    r[i] = v & LMASK;
    c = v >> LB;
  }
}

// to_canonical_biguint: ONE conditional subtract
template <class MOD>
P2E_HD void canon9(u32 (&a)[NL]) {
  if (geq_mod<MOD>(a)) {
    u32 t[NL];
    sub_mod9<MOD>(a, t);
#pragma unroll
    for (int i = 0; i < NL; i++) a[i] = t[i];
  }
}

// get_biguint_target + from_noncanonical_biguint + to_canonical_biguint on 9 raw wires
// (Horner over whatever u64 each wire holds, biguint.rs:444-452; absent limbs are passed as 0).
template <class MOD>
P2E_HD u32 load_canon(const u64 (&w)[NL], u32 (&a)[NL]) {
  u64 c = 0;
  u32 err = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    u64 l = w[i];
    u64 lo = l + c;
    u64 ov = lo < l ? 1 : 0;
    a[i] = (u32)lo & LMASK;
    c = (lo >> LB) | (ov << (64 - LB));
  }
  if (c != 0 || (a[NL - 1] >> 24) != 0) err |= ERR_NONCANONICAL;
  canon9<MOD>(a);
  return err;
}

// ------------------------------------------------------------------------------------------
// 9x9 schoolbook product into 17 un-carried 64-bit columns (each < 9 * 2^58 < 2^62)
P2E_HD void mul_cols(const u32 (&x)[NL], const u32 (&y)[NL], u64 (&col)[2 * NL - 1]) {
#pragma unroll
  for (int i = 0; i < 2 * NL - 1; i++) col[i] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
#pragma unroll
    for (int j = 0; j < NL; j++) col[i + j] += (u64)x[i] * (u64)y[j];
  }
}

P2E_HD void sqr_cols(const u32 (&x)[NL], u64 (&col)[2 * NL - 1]) {
#pragma unroll
  for (int i = 0; i < 2 * NL - 1; i++) col[i] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    col[2 * i] += (u64)x[i] * (u64)x[i];
#pragma unroll
    for (int j = i + 1; j < NL; j++) col[i + j] += (u64)(2 * x[i]) * (u64)x[j];
  }
}

// Exact (q, r) = divmod(T, M) for T < 2^522 given as 17 lazy columns (each < 2^63).
// M = 2^256 - c:  T = Hh*2^261 + Ll = Hh*32*M + (Hh*d + Ll), folded NFOLD times, then the bits
// above 2^256 are folded with c and one conditional subtraction finishes.  q10[9] is the limb the
// reference silently drops (mul_nonnative.rs:287-294); it is non-zero only for inputs >= 2^256.
template <class MOD, bool WANT_Q>
P2E_HD void divmod_cols(const u64 (&col)[2 * NL - 1], u32 (&r)[NL], u32 (&q10)[NL + 1]) {
  u64 T[18];
  {
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < 17; i++) {
      u64 v = col[i] + c;
      T[i] = v & LMASK;
      c = v >> LB;
    }
    T[17] = c;
  }
  u64 S[NL + 1];
#pragma unroll
  for (int i = 0; i < NL + 1; i++) S[i] = 0;

#pragma unroll
  for (int k = 0; k < MOD::NFOLD; k++) {
    const int nt = MOD::nt(k), nt2 = MOD::nt(k + 1), nh = nt - NL;
    u64 nw[18];
#pragma unroll
    for (int i = 0; i < 18; i++) nw[i] = (i < NL) ? T[i] : 0;
#pragma unroll
    for (int i = 0; i < 18 - NL; i++) {
      if (i < nh) {
        if (WANT_Q) S[i] += T[NL + i];
#pragma unroll
        for (int j = 0; j < MOD::ND; j++) nw[i + j] += (u64)(u32)T[NL + i] * (u64)MOD::d(j);
      }
    }
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < 18; i++) {
      if (i < nt2) {
        u64 v = nw[i] + c;
        if (i < nt2 - 1) {
          T[i] = v & LMASK;
          c = v >> LB;
        } else {
          T[i] = v;
        }
      } else {
        T[i] = 0;
      }
    }
  }
  // T < 2^261 in 9 normalised limbs
  u32 extra = 0;
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    u32 t = (u32)T[8] >> 24;
    T[8] &= TOPMASK;
    extra += t;
#pragma unroll
    for (int j = 0; j < MOD::NC; j++) T[j] += (u64)t * (u64)MOD::c(j);
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < NL - 1; i++) {
      u64 v = T[i] + c;
      T[i] = v & LMASK;
      c = v >> LB;
    }
    T[8] += c;
  }
  // T < 2^256 ; final conditional subtraction: w = T + c = T - M + 2^256
  u32 w[NL];
  {
    u32 c = 0;
#pragma unroll
    for (int i = 0; i < NL; i++) {
      u32 v = (u32)T[i] + (i < MOD::NC ? MOD::c(i) : 0u) + c;
      if (i < NL - 1) {
        w[i] = v & LMASK;
        c = v >> LB;
      } else {
        w[i] = v;
      }
    }
  }
  u32 fin = w[8] >> 24;
  extra += fin;
#pragma unroll
  for (int i = 0; i < NL; i++) r[i] = fin ? (i == NL - 1 ? (w[i] & TOPMASK) : w[i]) : (u32)T[i];
  if (WANT_Q) {
    u64 c = extra;
#pragma unroll
    for (int i = 0; i < NL + 1; i++) {
      u64 v = (S[i] << 5) + c;
      q10[i] = (u32)v & LMASK;
      c = v >> LB;
    }
  }
}

// r = x*y mod M  (x, y any 9-limb values < 2^261)
template <class MOD>
P2E_HD void mulmod(const u32 (&x)[NL], const u32 (&y)[NL], u32 (&r)[NL]) {
  u64 col[2 * NL - 1];
  mul_cols(x, y, col);
  u32 q10[NL + 1];
  divmod_cols<MOD, false>(col, r, q10);
}

template <class MOD>
P2E_HD void sqrmod(const u32 (&x)[NL], u32 (&r)[NL]) {
  u64 col[2 * NL - 1];
  sqr_cols(x, col);
  u32 q10[NL + 1];
  divmod_cols<MOD, false>(col, r, q10);
}

template <class MOD>
P2E_HD void sqrmod_n(u32 (&x)[NL], int n) {
  for (int i = 0; i < n; i++) {
    u32 t[NL];
    sqrmod<MOD>(x, t);
#pragma unroll
    for (int j = 0; j < NL; j++) x[j] = t[j];
  }
}

P2E_HD void copy9(const u32 (&a)[NL], u32 (&r)[NL]) {
#pragma unroll
  for (int i = 0; i < NL; i++) r[i] = a[i];
}

// modular add / sub / neg on canonical values (used by the projective chain replay)
template <class MOD>
P2E_HD void addmod(const u32 (&a)[NL], const u32 (&b)[NL], u32 (&r)[NL]) {
  add9(a, b, r);          // < 2^257: no carry-out
  canon9<MOD>(r);
}

template <class MOD>
P2E_HD void submod(const u32 (&a)[NL], const u32 (&b)[NL], u32 (&r)[NL]) {
  if (geq9(a, b)) {
    sub9(a, b, r);
  } else {
    u32 t[NL];
    add_mod9<MOD>(a, t);
    sub9(t, b, r);
  }
}

// x^(p-2) mod p for the secp256k1 base field: 255 squarings + 15 multiplications
// (addition chain over the run structure of p-2 = 2^256 - 2^32 - 979).  The inverse is unique,
// so this equals the reference's Fermat exp_biguint(p-2) bit for bit.
P2E_HD void inv_p(const u32 (&a)[NL], u32 (&r)[NL]) {
  u32 x2[NL], x3[NL], x22[NL], x44[NL], t[NL], u[NL];
  sqrmod<ModP>(a, t); mulmod<ModP>(t, a, x2);
  sqrmod<ModP>(x2, t); mulmod<ModP>(t, a, x3);
  copy9(x3, t); sqrmod_n<ModP>(t, 3); mulmod<ModP>(t, x3, u);          // x6
  copy9(u, t); sqrmod_n<ModP>(t, 3); mulmod<ModP>(t, x3, u);           // x9
  copy9(u, t); sqrmod_n<ModP>(t, 2); mulmod<ModP>(t, x2, u);           // x11
  copy9(u, t); sqrmod_n<ModP>(t, 11); mulmod<ModP>(t, u, x22);         // x22
  copy9(x22, t); sqrmod_n<ModP>(t, 22); mulmod<ModP>(t, x22, x44);     // x44
  copy9(x44, t); sqrmod_n<ModP>(t, 44); mulmod<ModP>(t, x44, u);       // x88
  u32 x88[NL]; copy9(u, x88);
  copy9(x88, t); sqrmod_n<ModP>(t, 88); mulmod<ModP>(t, x88, u);       // x176
  copy9(u, t); sqrmod_n<ModP>(t, 44); mulmod<ModP>(t, x44, u);         // x220
  copy9(u, t); sqrmod_n<ModP>(t, 3); mulmod<ModP>(t, x3, u);           // x223
  copy9(u, t); sqrmod_n<ModP>(t, 23); mulmod<ModP>(t, x22, u);
  copy9(u, t); sqrmod_n<ModP>(t, 5); mulmod<ModP>(t, a, u);
  copy9(u, t); sqrmod_n<ModP>(t, 3); mulmod<ModP>(t, x2, u);
  copy9(u, t); sqrmod_n<ModP>(t, 2); mulmod<ModP>(t, a, r);
}

// x^(n-2) mod n for the scalar field: 4-bit fixed window over the constant exponent.
P2E_HD void inv_n(const u32 (&a)[NL], u32 (&r)[NL]) {
  // n - 2, little-endian 32-bit words
  const u32 e[8] = {0xd036413fu, 0xbfd25e8cu, 0xaf48a03bu, 0xbaaedce6u,
                    0xfffffffeu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
  u32 tab[16][NL];
  set_zero9(tab[0]);
  tab[0][0] = 1;
  copy9(a, tab[1]);
  for (int i = 2; i < 16; i++) mulmod<ModN>(tab[i - 1], a, tab[i]);
  u32 acc[NL];
  copy9(tab[15], acc);                     // top nibble of n-2 is 0xf
  for (int nib = 62; nib >= 0; nib--) {
    sqrmod_n<ModN>(acc, 4);
    u32 w = (e[nib >> 3] >> ((nib & 7) * 4)) & 15u;
    if (w) {
      u32 t[NL];
      mulmod<ModN>(acc, tab[w], t);
      copy9(t, acc);
    }
  }
  copy9(acc, r);
}

// ------------------------------------------------------------------------------------------
// Montgomery arithmetic mod n (scalar field) in 29-bit limbs, R = 2^261.  Because R > 32 n, products
// of values below 2^258 stay below 2^257 without a final subtraction, so a whole exponentiation runs
// on "almost reduced" values and is canonicalised once.  ~235 instructions per product with a
// 9-step dependency chain, against ~600 (and serial carry ripples) for the exact special-form
// reduction: this is what the per-signature scalar inverse c = s^-1 (ecdsa.rs:40) runs on.
constexpr u32 MONT_N_PRIME = 0x1588b13fu;           // -n^-1 mod 2^29
P2E_HD void mont_n_reduce(u64 (&col)[2 * NL], u32 (&r)[NL]) {
#pragma unroll
  for (int i = 0; i < NL; i++) {
    u32 m = ((u32)col[i] * MONT_N_PRIME) & LMASK;
#pragma unroll
    for (int j = 0; j < NL; j++) col[i + j] += (u64)m * (u64)ModN::m(j);
    col[i + 1] += col[i] >> LB;                       // low 29 bits of col[i] are now zero
  }
  u64 c = 0;
#pragma unroll
  for (int j = 0; j < NL; j++) {
    u64 v = col[NL + j] + c;
    r[j] = (u32)v & LMASK;
    c = v >> LB;
  }
}
P2E_HD void mont_n_mul(const u32 (&a)[NL], const u32 (&b)[NL], u32 (&r)[NL]) {
  u64 col[2 * NL];
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) col[i] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
#pragma unroll
    for (int j = 0; j < NL; j++) col[i + j] += (u64)a[i] * (u64)b[j];
  }
  mont_n_reduce(col, r);
}
P2E_HD void mont_n_sqr(const u32 (&a)[NL], u32 (&r)[NL]) {
  u64 col[2 * NL];
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) col[i] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    col[2 * i] += (u64)a[i] * (u64)a[i];
#pragma unroll
    for (int j = i + 1; j < NL; j++) col[i + j] += (u64)(2 * a[i]) * (u64)a[j];
  }
  mont_n_reduce(col, r);
}

// x^(n-2) mod n, canonical: 4-bit fixed window over the constant exponent, in the Montgomery domain
P2E_HD void inv_n_mont(const u32 (&a)[NL], u32 (&r)[NL]) {
  const u32 R2[NL] = {0x09f6ab4bu, 0x1f300d1eu, 0x1c0bd5d8u, 0x0c8ada8cu, 0x11cefa2bu,
                      0x08b79a0fu, 0x1e697f5eu, 0x00e34de2u, 0x009c7356u};          // R^2 mod n
  const u32 e[8] = {0xd036413fu, 0xbfd25e8cu, 0xaf48a03bu, 0xbaaedce6u,
                    0xfffffffeu, 0xffffffffu, 0xffffffffu, 0xffffffffu};              // n - 2
  u32 tab[16][NL];
  mont_n_mul(a, R2, tab[1]);                         // a R
  copy9(tab[1], tab[0]);                             // slot 0 unused (window 0 skips the multiply)
  for (int i = 2; i < 16; i++) mont_n_mul(tab[i - 1], tab[1], tab[i]);
  u32 acc[NL], t[NL];
  copy9(tab[15], acc);                               // top nibble of n-2 is 0xf
  for (int nib = 62; nib >= 0; nib--) {
    for (int k = 0; k < 4; k++) {
      mont_n_sqr(acc, t);
      copy9(t, acc);
    }
    u32 w = (e[nib >> 3] >> ((nib & 7) * 4)) & 15u;
    if (w) {
      mont_n_mul(acc, tab[w], t);
      copy9(t, acc);
    }
  }
  u32 one[NL];
  set_zero9(one);
  one[0] = 1;
  mont_n_mul(acc, one, t);                           // leave the Montgomery domain: < 2^257
  canon9<ModN>(t);
  canon9<ModN>(t);
  copy9(t, r);
}

// the reference's own algorithm (a ladder): kept as the cross-check of inv_mod in tests/hostcheck
template <class MOD>
P2E_HD void inv_mod_ladder(const u32 (&a)[NL], u32 (&r)[NL]) {
  if (MOD::ID == 0) inv_p(a, r); else inv_n_mont(a, r);
}
}  // namespace p2e
#include "p2e_modinv.cuh"
namespace p2e {
// canonical 29-bit limbs in (a < m), canonical 29-bit limbs out: division steps (p2e_modinv.cuh)
template <class MOD>
P2E_HD void inv_mod(const u32 (&a)[NL], u32 (&r)[NL]) {
  u32 w[8], o[8];
#pragma unroll
  for (int k = 0; k < 8; k++) w[k] = (a[k] >> (3 * k)) | (a[k + 1] << (LB - 3 * k));
  modinv_words<MOD>(w, o);
  r[0] = o[0] & LMASK;
#pragma unroll
  for (int i = 1; i < 8; i++) r[i] = ((o[i - 1] >> (32 - 3 * i)) | (o[i] << (3 * i))) & LMASK;
  r[8] = o[7] >> 8;
}

// ------------------------------------------------------------------------------------------
// "Lazy" base-field arithmetic for the latency-bound chain replay (p2e_chain.cu K1-K3).
// A value is 9 SIGNED 32-bit limbs in redundant form, congruent to the field element mod p:
//     invariant LZ:  -2^20 <= l_i <= 2^29 + 2^20
// Products are reduced with 2^261 = d (mod p) using only limb-parallel carry passes (no serial
// ripple), so a multiplication has a ~100-cycle dependency chain instead of ~1000.  Exact
// canonical limbs are produced once per stored output by lz_canon().
typedef int32_t i32;
constexpr i64 LZ_D0 = 0x7a20, LZ_D1 = 0x100;      // ModP::d

P2E_HD void lz_from_canon(const u32 (&a)[NL], i32 (&r)[NL]) {
#pragma unroll
  for (int i = 0; i < NL; i++) r[i] = (i32)a[i];
}

// columns (|col_i| < 2^62) -> LZ limbs
P2E_HD void lz_reduce_cols(const i64 (&col)[2 * NL - 1], i32 (&r)[NL]) {
  // pass A: split every column, push carries one limb up (all independent)
  i64 n[2 * NL];
  n[0] = col[0] & LMASK;
#pragma unroll
  for (int i = 1; i < 2 * NL - 1; i++) n[i] = (col[i] & LMASK) + (col[i - 1] >> LB);
  n[2 * NL - 1] = col[2 * NL - 2] >> LB;
  // fold A: limbs 9..17 times d = 2^261 mod p
  i64 t[NL + 1];
#pragma unroll
  for (int i = 0; i < NL; i++) t[i] = n[i] + n[NL + i] * LZ_D0 + (i > 0 ? n[NL + i - 1] * LZ_D1 : 0);
  t[NL] = n[2 * NL - 1] * LZ_D1;
  // pass B
  i64 u[NL + 2];
  u[0] = t[0] & LMASK;
#pragma unroll
  for (int i = 1; i < NL + 1; i++) u[i] = (t[i] & LMASK) + (t[i - 1] >> LB);
  u[NL + 1] = t[NL] >> LB;
  // fold B: limbs 9, 10
  i64 w0 = u[0] + u[NL] * LZ_D0;
  i64 w1 = u[1] + u[NL] * LZ_D1 + u[NL + 1] * LZ_D0;
  i64 w2 = u[2] + u[NL + 1] * LZ_D1;
  // pass C on the three touched limbs
  r[0] = (i32)(w0 & LMASK);
  r[1] = (i32)((w1 & LMASK) + (w0 >> LB));
  r[2] = (i32)((w2 & LMASK) + (w1 >> LB));
  r[3] = (i32)(u[3] + (w2 >> LB));
#pragma unroll
  for (int i = 4; i < NL; i++) r[i] = (i32)u[i];
}

P2E_HD void lz_mul(const i32 (&a)[NL], const i32 (&b)[NL], i32 (&r)[NL]) {
  i64 col[2 * NL - 1];
#pragma unroll
  for (int i = 0; i < 2 * NL - 1; i++) col[i] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
#pragma unroll
    for (int j = 0; j < NL; j++) col[i + j] += (i64)a[i] * (i64)b[j];
  }
  lz_reduce_cols(col, r);
}

P2E_HD void lz_sqr(const i32 (&a)[NL], i32 (&r)[NL]) {
  i64 col[2 * NL - 1];
#pragma unroll
  for (int i = 0; i < 2 * NL - 1; i++) col[i] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    col[2 * i] += (i64)a[i] * (i64)a[i];
#pragma unroll
    for (int j = i + 1; j < NL; j++) col[i + j] += (i64)(2 * a[i]) * (i64)a[j];
  }
  lz_reduce_cols(col, r);
}

// one limb-parallel carry pass on |l_i| < 2^34 -> LZ
P2E_HD void lz_weak(const i64 (&l)[NL], i32 (&r)[NL]) {
  const i64 c8 = l[NL - 1] >> LB;
  r[0] = (i32)((l[0] & LMASK) + c8 * LZ_D0);
  r[1] = (i32)((l[1] & LMASK) + (l[0] >> LB) + c8 * LZ_D1);
#pragma unroll
  for (int i = 2; i < NL; i++) r[i] = (i32)((l[i] & LMASK) + (l[i - 1] >> LB));
}

P2E_HD void lz_add(const i32 (&a)[NL], const i32 (&b)[NL], i32 (&r)[NL]) {
  i64 t[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) t[i] = (i64)a[i] + b[i];
  lz_weak(t, r);
}
P2E_HD void lz_sub(const i32 (&a)[NL], const i32 (&b)[NL], i32 (&r)[NL]) {
  i64 t[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) t[i] = (i64)a[i] - b[i];
  lz_weak(t, r);
}
// r = ka * a - kb * b for small non-negative constants (ka, kb <= 8)
P2E_HD void lz_lin(const i32 (&a)[NL], int ka, const i32 (&b)[NL], int kb, i32 (&r)[NL]) {
  i64 t[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) t[i] = (i64)a[i] * ka - (i64)b[i] * kb;
  lz_weak(t, r);
}
P2E_HD void lz_copy(const i32 (&a)[NL], i32 (&r)[NL]) {
#pragma unroll
  for (int i = 0; i < NL; i++) r[i] = a[i];
}
P2E_HD void lz_sqr_n(i32 (&x)[NL], int n) {
  for (int i = 0; i < n; i++) {
    i32 t[NL];
    lz_sqr(x, t);
    lz_copy(t, x);
  }
}

// exact canonical limbs in [0, p): add 4p limb-wise (all limbs become non-negative), then the
// exact fold / conditional-subtract tail of divmod_cols
P2E_HD void lz_canon(const i32 (&a)[NL], u32 (&r)[NL]) {
  u64 col[2 * NL - 1];
#pragma unroll
  for (int i = 0; i < 2 * NL - 1; i++) col[i] = (i < NL) ? (u64)((i64)a[i] + 4 * (i64)ModP::m(i)) : 0;
  u32 q10[NL + 1];
  divmod_cols<ModP, false>(col, r, q10);
}

// x^(p-2): same addition chain as inv_p, in lazy arithmetic
P2E_HD void lz_inv(const i32 (&a)[NL], i32 (&r)[NL]) {
  i32 x2[NL], x3[NL], x22[NL], x44[NL], x88[NL], t[NL], u[NL];
  lz_sqr(a, t); lz_mul(t, a, x2);
  lz_sqr(x2, t); lz_mul(t, a, x3);
  lz_copy(x3, t); lz_sqr_n(t, 3); lz_mul(t, x3, u);          // x6
  lz_copy(u, t); lz_sqr_n(t, 3); lz_mul(t, x3, u);           // x9
  lz_copy(u, t); lz_sqr_n(t, 2); lz_mul(t, x2, u);           // x11
  lz_copy(u, t); lz_sqr_n(t, 11); lz_mul(t, u, x22);         // x22
  lz_copy(x22, t); lz_sqr_n(t, 22); lz_mul(t, x22, x44);     // x44
  lz_copy(x44, t); lz_sqr_n(t, 44); lz_mul(t, x44, x88);     // x88
  lz_copy(x88, t); lz_sqr_n(t, 88); lz_mul(t, x88, u);       // x176
  lz_copy(u, t); lz_sqr_n(t, 44); lz_mul(t, x44, u);         // x220
  lz_copy(u, t); lz_sqr_n(t, 3); lz_mul(t, x3, u);           // x223
  lz_copy(u, t); lz_sqr_n(t, 23); lz_mul(t, x22, u);
  lz_copy(u, t); lz_sqr_n(t, 5); lz_mul(t, a, u);
  lz_copy(u, t); lz_sqr_n(t, 3); lz_mul(t, x2, u);
  lz_copy(u, t); lz_sqr_n(t, 2); lz_mul(t, a, r);
}

// ------------------------------------------------------------------------------------------
// A8 + A9 : MulNonnativeGenerator + CheckSumGenerator for one (x, y) pair.
// q = low 9 limbs of floor(x*y / M) (10th limb dropped like the reference), r = x*y mod M,
// cs[i] canonical Goldilocks, b[i] = carry_i + 2^33.
P2E_HD u32 check_sum_slow(const u64 (&cs)[2 * NL - 1], u64 (&b)[2 * NL - 2]) {
  // literal Goldilocks field division (mul_nonnative.rs:521-527); only reached when the
  // un-carried columns do not carry exactly, i.e. when the reference's assert would fire
  u64 last = 0;
  u32 err = 0;
  for (int i = 0; i < 2 * NL - 2; i++) {
    u64 bf = gl_mul(gl_add(cs[i], last), GL_INV_2_29);
    u64 o = gl_add(bf, 1ull << 33);
    b[i] = o;
    last = bf;
    if (o >= (1ull << 34)) err |= ERR_CARRY_RANGE;
  }
  return err;
}

template <class MOD>
P2E_HD u32 mul_witness_core(const u32 (&x)[NL], const u32 (&y)[NL], u32 (&r)[NL], u32 (&q)[NL],
                            u64 (&cs)[2 * NL - 1], u64 (&b)[2 * NL - 2]) {
  u64 col[2 * NL - 1];
  mul_cols(x, y, col);
  u32 q10[NL + 1];
  divmod_cols<MOD, true>(col, r, q10);
#pragma unroll
  for (int i = 0; i < NL; i++) q[i] = q10[i];
  // un-carried check_sum columns: sum_j q[i-j]*m[j] - x*y column + r[i]   (|v| < 2^62)
  i64 v[2 * NL - 1];
#pragma unroll
  for (int i = 0; i < 2 * NL - 1; i++) v[i] = -(i64)col[i] + (i < NL ? (i64)r[i] : 0);
#pragma unroll
  for (int i = 0; i < NL; i++) {
#pragma unroll
    for (int j = 0; j < NL; j++) v[i + j] += (i64)((u64)q[i] * (u64)MOD::m(j));
  }
  // exact carries: every partial sum must be a multiple of 2^29 and every offset carry must fit 34
  // bits; both conditions are accumulated branch-free and tested once
  u64 bad = 0;
  i64 c = 0;
  u32 err = 0;
#pragma unroll
  for (int i = 0; i < 2 * NL - 1; i++) {
    cs[i] = gl_from_i64(v[i]);
    if (i < 2 * NL - 2) {
      i64 s = v[i] + c;
      bad |= (u64)s & LMASK;
      c = s >> LB;
      u64 o = (u64)(c + ((i64)1 << 33));
      b[i] = o;
      bad |= o >> 34;
    }
  }
  if (bad) err |= check_sum_slow(cs, b);
  return err;
}

// CheckSumGenerator alone on 17 canonical Goldilocks wires (standalone A9 entry point)
P2E_HD u32 check_sum_core(const u64 (&a)[2 * NL - 1], u64 (&b)[2 * NL - 2]) {
  bool exact = true;
  i64 c = 0;
#pragma unroll
  for (int i = 0; i < 2 * NL - 2; i++) {
    // signed view of the wire; exact path only when |a| < 2^62
    u64 w = a[i];
    i64 sv;
    if (w < (1ull << 62)) sv = (i64)w;
    else if (w > GL_P - (1ull << 62)) sv = -(i64)(GL_P - w);
    else { exact = false; sv = 0; }
    i64 s = sv + c;
    if ((u64)s & LMASK) exact = false;
    c = s >> LB;
    i64 o = c + ((i64)1 << 33);
    b[i] = (u64)o;
    if (o < 0 || o >= ((i64)1 << 34)) exact = false;
  }
  if (!exact) return check_sum_slow(a, b);
  return 0;
}

// mul gate generator input checks (mul_nonnative.rs:256-277)
P2E_HD u32 load_mul_wires(const u64 (&w)[NL], u32 (&a)[NL]) {
  u32 err = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    if (w[i] >> 32) err |= ERR_U32;
    else if (w[i] >> LB) err |= ERR_LIMB_RANGE;
    a[i] = (u32)w[i] & LMASK;
  }
  return err;
}

// ------------------------------------------------------------------------------------------
// A4 NonNativeAdditionGenerator (strict '>' quirk, nonnative.rs:637): inputs canonical
template <class MOD>
P2E_HD u32 add_witness_core(const u32 (&a)[NL], const u32 (&b)[NL], u32 (&sum)[NL]) {
  u32 s[NL];
  add9(a, b, s);
  // s > M  <=>  s >= M and s != M
  bool gt = false;
  if (geq_mod<MOD>(s)) {
    u32 t[NL];
    sub_mod9<MOD>(s, t);
    gt = !is_zero9(t);
    if (gt) copy9(t, s);
  }
  copy9(s, sum);
  return gt ? 1u : 0u;
}

// A6 NonNativeSubtractionGenerator: inputs canonical
template <class MOD>
P2E_HD u32 sub_witness_core(const u32 (&a)[NL], const u32 (&b)[NL], u32 (&diff)[NL]) {
  if (geq9(a, b)) {
    sub9(a, b, diff);
    return 0;
  }
  u32 t[NL];
  add_mod9<MOD>(a, t);
  sub9(t, b, diff);
  return 1;
}

// A7 tail: div = floor(x_c * inv / M) given x_c canonical and inv = x_c^-1
template <class MOD>
P2E_HD u32 inv_div_core(const u32 (&xc)[NL], const u32 (&inv)[NL], u32 (&div)[NL]) {
  u64 col[2 * NL - 1];
  mul_cols(xc, inv, col);
  u32 r[NL], q10[NL + 1];
  divmod_cols<MOD, true>(col, r, q10);
#pragma unroll
  for (int i = 0; i < NL; i++) div[i] = q10[i];
  // self check: x * inv == 1 (mod M)
  u32 bad = r[0] ^ 1u;
#pragma unroll
  for (int i = 1; i < NL; i++) bad |= r[i];
  return bad ? ERR_INTERNAL : 0u;
}

// ------------------------------------------------------------------------------------------
// A10 decompose_secp256k1_scalar (curve/glv.rs:39-77).  k canonical scalar.
// round(B*k/n) = floor((2*B*k + n) / (2n)) = floor(divmod(2B*k + n, n).q / 2)   (n odd)
P2E_HD void glv_round_div(const u32 (&twoB)[NL], const u32 (&k)[NL], u32 (&cq)[NL]) {
  u64 col[2 * NL - 1];
  mul_cols(twoB, k, col);
#pragma unroll
  for (int i = 0; i < NL; i++) col[i] += ModN::m(i);
  u32 r[NL], q10[NL + 1];
  divmod_cols<ModN, true>(col, r, q10);
#pragma unroll
  for (int i = 0; i < NL; i++) cq[i] = ((q10[i] >> 1) | (q10[i + 1] << (LB - 1))) & LMASK;
}

P2E_HD void glv_decompose_core(const u32 (&k)[NL], u32 (&k1)[NL], u32 (&k2)[NL], u32& k1_neg, u32& k2_neg) {
  const u32 TWO_B2[NL] = {0x0509d62au, 0x06c90e49u, 0x0a35e6f4u, 0x1b48869fu, 0x00000610u, 0, 0, 0, 0};
  const u32 TWO_MB1[NL] = {0x157fc986u, 0x1547fa90u, 0x07441437u, 0x0dfb5804u, 0x00001c88u, 0, 0, 0, 0};
  const u32 A1[NL] = {0x1284eb15u, 0x03648724u, 0x151af37au, 0x0da4434fu, 0x00000308u, 0, 0, 0, 0};
  const u32 MB1[NL] = {0x0abfe4c3u, 0x1aa3fd48u, 0x03a20a1bu, 0x06fdac02u, 0x00000e44u, 0, 0, 0, 0};
  const u32 A2[NL] = {0x1d44cfd8u, 0x1e08846cu, 0x18bcfd95u, 0x14a1ef51u, 0x0000114cu, 0, 0, 0, 0};
  u32 c1[NL], c2[NL], t1[NL], t2[NL], t3[NL];
  glv_round_div(TWO_B2, k, c1);
  glv_round_div(TWO_MB1, k, c2);
  // k1_raw = k - c1*A1 - c2*A2 ; k2_raw = c1*MINUS_B1 - c2*B2   (B2 == A1)
  mulmod<ModN>(c1, A1, t1);
  mulmod<ModN>(c2, A2, t2);
  submod<ModN>(k, t1, t3);
  submod<ModN>(t3, t2, k1);
  mulmod<ModN>(c1, MB1, t1);
  mulmod<ModN>(c2, A1, t2);
  submod<ModN>(t1, t2, k2);
  u32 half[NL], mm[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) { half[i] = ModN::half(i); mm[i] = ModN::m(i); }
  // neg <=> value > floor(n/2)
  k1_neg = (geq9(k1, half) && !eq9(k1, half)) ? 1u : 0u;
  if (k1_neg) { sub9(mm, k1, t1); copy9(t1, k1); }
  k2_neg = (geq9(k2, half) && !eq9(k2, half)) ? 1u : 0u;
  if (k2_neg) { sub9(mm, k2, t1); copy9(t1, k2); }
}

}  // namespace p2e
