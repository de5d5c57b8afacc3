// p2e_f8.cuh -- secp256k1 base-field arithmetic on 8 saturated 32-bit words with hardware carry chains, for the
// chain kernels K1-K3 (p2e_chain.cu) where no 29-bit limb is ever an output: a product is 64 IMAD.WIDE.U32(.X)
// instead of 81 IMAD.WIDE plus carry passes, carries live in the CC flag instead of SHF / LOP3 / IADD3 triples
// (~140 SASS instructions per multiplication against ~240 for the lazy 29-bit form; scripts/sass_mix.py).
//
// A value is "weak": any 256-bit integer congruent to the field element mod p = 2^256 - 2^32 - 977.  Every
// operation takes weak inputs and returns a weak result; f8_canon() gives the representative in [0, p).
// The reference has no counterpart (it runs num::BigUint); everything here is an implementation detail behind
// the inverse-is-unique argument of DESIGN.md section 4 -- the witness bits come out of the generator cores.
//
// Each operation has two bodies computing the SAME function word for word: PTX carry chains (device) and a
// portable one (host: tape constants, tests/hostcheck).  tests/test_gpu_f8.py compares them on the GPU.
#pragma once
#include "p2e_field.cuh"

namespace p2e {

constexpr u32 F8_C = 977u;   // 2^256 = 2^32 + 977 (mod p)

struct F8 {
  u32 v[8];
};

// ------------------------------------------------------------------------------------------
// portable bodies (also the specification of the PTX bodies)
P2E_HD void f8p_mul_wide(const u32 (&a)[8], const u32 (&b)[8], u32 (&t)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u64 c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      u64 v = (u64)a[i] * (u64)b[j] + t[i + j] + c;
      t[i + j] = (u32)v;
      c = v >> 32;
    }
    t[i + 8] = (u32)c;
  }
}

// fold a 9-word value (w[8] <= 2^32 + 978) to 8 words: twice 2^256 -> 2^32 + 977
P2E_HD void f8p_fold9(const u32 (&lo)[8], u64 top, u32 (&r)[8]) {
  // S = lo + top * (2^32 + 977), top < 2^34
  u64 c = 0;
  u64 add0 = (top & 0xffffffffull) * F8_C;            // < 2^42
  u64 add1 = (top >> 32) * F8_C + (top & 0xffffffffull);   // word 1
  u64 add2 = top >> 32;                                    // word 2
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u64 v = c + lo[k] + (k == 0 ? (add0 & 0xffffffffull) : k == 1 ? (add0 >> 32) + add1 : k == 2 ? add2 : 0);
    r[k] = (u32)v;
    c = v >> 32;
  }
  // c = 1 only when r wrapped, i.e. r < 2^67: the second fold cannot carry past word 2
  u64 v = (u64)r[0] + c * F8_C;
  r[0] = (u32)v;
  v = (u64)r[1] + c + (v >> 32);
  r[1] = (u32)v;
  r[2] += (u32)(v >> 32);
}

P2E_HD void f8p_reduce16(const u32 (&t)[16], u32 (&r)[8]) {
  // R = L + H * (2^32 + 977), 10 words
  u32 R[8];
  u64 c = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u64 v = c + t[k] + (u64)t[8 + k] * F8_C + (k ? t[8 + k - 1] : 0u);
    R[k] = (u32)v;
    c = v >> 32;
  }
  const u64 top = c + t[15];                              // words 8 and 9 of R as one integer
  f8p_fold9(R, top, r);
}

P2E_HD void f8p_add(const u32 (&a)[8], const u32 (&b)[8], u32 (&r)[8]) {
  u32 s[8];
  u64 c = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u64 v = c + a[k] + b[k];
    s[k] = (u32)v;
    c = v >> 32;
  }
  f8p_fold9(s, c, r);
}

P2E_HD void f8p_sub(const u32 (&a)[8], const u32 (&b)[8], u32 (&r)[8]) {
  // d = a - b + 2^256 beta ; true value = d - beta (2^32 + 977), applied twice (the second borrow is rare)
  i64 br = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    i64 v = (i64)a[k] - (i64)b[k] + br;
    r[k] = (u32)v;
    br = v >> 32;
  }
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    const i64 beta = -br;                                 // 0 or 1
    br = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      i64 v = (i64)r[k] - (k == 0 ? beta * (i64)F8_C : k == 1 ? beta : 0) + br;
      r[k] = (u32)v;
      br = v >> 32;
    }
  }
}

template <int K>
P2E_HD void f8p_mul_small(const u32 (&a)[8], u32 (&r)[8]) {
  u32 s[8];
  u64 c = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u64 v = c + (u64)a[k] * (u64)K;
    s[k] = (u32)v;
    c = v >> 32;
  }
  f8p_fold9(s, c, r);
}

P2E_HD void f8p_canon(u32 (&r)[8]) {
  u32 t[8];
  u64 c = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u64 v = c + r[k] + (k == 0 ? F8_C : k == 1 ? 1u : 0u);
    t[k] = (u32)v;
    c = v >> 32;
  }
  if (c) {
#pragma unroll
    for (int k = 0; k < 8; k++) r[k] = t[k];
  }
}

// ------------------------------------------------------------------------------------------
// PTX bodies.  A "row" is a carry chain of aligned (lo, hi) word pairs: products a_i * b land on consecutive
// pairs when the a_i are two words apart, and ptxas fuses each mad.lo.cc / madc.hi.cc pair into ONE
// IMAD.WIDE.U32(.X) on an even-aligned register pair.  Two accumulators: `e` collects the products with i + j even
// (weight 2^(32 (i+j))), `o` those with i + j odd (array index = i + j - 1).  A row whose top pair has a fresh
// (zero) high word cannot carry out (x + (2^32-1)^2 + 1 < 2^64); the other rows carry into a fresh word.
#ifdef __CUDA_ARCH__
#define F8_ROW4_FIRST(e0, e1, e2, e3, e4, e5, e6, e7, a0, a1, a2, a3, b) \
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %8, %12;\n\tmov.b64 {%0, %1}, t;\n\tmul.wide.u32 t, %9, %12;\n\tmov.b64 {%2, %3}, t;\n\tmul.wide.u32 t, %10, %12;\n\tmov.b64 {%4, %5}, t;\n\tmul.wide.u32 t, %11, %12;\n\tmov.b64 {%6, %7}, t;\n\t}" : "=r"(e0), "=r"(e1), "=r"(e2), "=r"(e3), "=r"(e4), "=r"(e5), "=r"(e6), "=r"(e7) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b))
#define F8_ROW3_FIRST(e0, e1, e2, e3, e4, e5, a0, a1, a2, b) \
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %6, %9;\n\tmov.b64 {%0, %1}, t;\n\tmul.wide.u32 t, %7, %9;\n\tmov.b64 {%2, %3}, t;\n\tmul.wide.u32 t, %8, %9;\n\tmov.b64 {%4, %5}, t;\n\t}" : "=r"(e0), "=r"(e1), "=r"(e2), "=r"(e3), "=r"(e4), "=r"(e5) : "r"(a0), "r"(a1), "r"(a2), "r"(b))
#define F8_ROW4_NC(e0, e1, e2, e3, e4, e5, e6, e7, a0, a1, a2, a3, b)                                             \
  asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"                                     \
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"                                    \
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"                                  \
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"                                         \
      : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(e5), "+r"(e6), "+r"(e7)                           \
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b))
#define F8_ROW4_C(e0, e1, e2, e3, e4, e5, e6, e7, e8, a0, a1, a2, a3, b)                                          \
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                     \
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                  \
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                  \
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, 0, 0;"                \
      : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(e5), "+r"(e6), "+r"(e7), "=r"(e8)                 \
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b))
#define F8_ROW3_NC(e0, e1, e2, e3, e4, e5, a0, a1, a2, b)                                                         \
  asm("mad.lo.cc.u32 %0, %6, %9, %0;\n\tmadc.hi.cc.u32 %1, %6, %9, %1;\n\t"                                       \
      "madc.lo.cc.u32 %2, %7, %9, %2;\n\tmadc.hi.cc.u32 %3, %7, %9, %3;\n\t"                                      \
      "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.u32 %5, %8, %9, %5;"                                             \
      : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(e5)                                               \
      : "r"(a0), "r"(a1), "r"(a2), "r"(b))
#define F8_ROW3_C(e0, e1, e2, e3, e4, e5, e6, a0, a1, a2, b)                                                      \
  asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\tmadc.hi.cc.u32 %1, %7, %10, %1;\n\t"                                     \
      "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"                                    \
      "madc.lo.cc.u32 %4, %9, %10, %4;\n\tmadc.hi.cc.u32 %5, %9, %10, %5;\n\taddc.u32 %6, 0, 0;"                  \
      : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(e5), "=r"(e6)                                     \
      : "r"(a0), "r"(a1), "r"(a2), "r"(b))
#define F8_ROW2_NC(e0, e1, e2, e3, a0, a1, b)                                                                     \
  asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\tmadc.hi.cc.u32 %1, %4, %6, %1;\n\t"                                       \
      "madc.lo.cc.u32 %2, %5, %6, %2;\n\tmadc.hi.u32 %3, %5, %6, %3;"                                             \
      : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3)                                                                    \
      : "r"(a0), "r"(a1), "r"(b))
#define F8_ROW2_C(e0, e1, e2, e3, e4, a0, a1, b)                                                                  \
  asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\t"                                       \
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\tmadc.hi.cc.u32 %3, %6, %7, %3;\n\taddc.u32 %4, 0, 0;"                    \
      : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "=r"(e4)                                                          \
      : "r"(a0), "r"(a1), "r"(b))
#define F8_ROW1_NC(e0, e1, a0, b) \
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(e0), "+r"(e1) : "r"(a0), "r"(b))
#define F8_ROW1_C(e0, e1, e2, a0, b)                                                            \
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, 0, 0;" \
      : "+r"(e0), "+r"(e1), "=r"(e2)                                                            \
      : "r"(a0), "r"(b))

// t = e + (o << 32) for words lo..15 where e[k] pairs with o[k-1]
__device__ __forceinline__ void f8d_merge16(const u32 (&e)[16], const u32 (&o)[16], u32 (&t)[16]) {
  t[0] = e[0];
  asm("add.cc.u32 %0, %1, %2;" : "=r"(t[1]) : "r"(e[1]), "r"(o[0]));
#pragma unroll
  for (int k = 2; k < 15; k++) asm("addc.cc.u32 %0, %1, %2;" : "=r"(t[k]) : "r"(e[k]), "r"(o[k - 1]));
  asm("addc.u32 %0, %1, %2;" : "=r"(t[15]) : "r"(e[15]), "r"(o[14]));
}

__device__ __forceinline__ void f8d_mul_wide(const u32 (&a)[8], const u32 (&b)[8], u32 (&t)[16]) {
  u32 e[16], o[16];
#pragma unroll
  for (int i = 8; i < 16; i++) { e[i] = 0; o[i] = 0; }
  F8_ROW4_FIRST(e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], a[0], a[2], a[4], a[6], b[0]);
  F8_ROW4_FIRST(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], a[1], a[3], a[5], a[7], b[0]);
#pragma unroll
  for (int j = 1; j < 8; j++) {
    if (j & 1) {
      F8_ROW4_C(o[j - 1], o[j], o[j + 1], o[j + 2], o[j + 3], o[j + 4], o[j + 5], o[j + 6], o[j + 7], a[0], a[2], a[4], a[6], b[j]);
      F8_ROW4_NC(e[j + 1], e[j + 2], e[j + 3], e[j + 4], e[j + 5], e[j + 6], e[j + 7], e[j + 8], a[1], a[3], a[5], a[7], b[j]);
    } else {
      F8_ROW4_C(e[j], e[j + 1], e[j + 2], e[j + 3], e[j + 4], e[j + 5], e[j + 6], e[j + 7], e[j + 8], a[0], a[2], a[4], a[6], b[j]);
      F8_ROW4_NC(o[j], o[j + 1], o[j + 2], o[j + 3], o[j + 4], o[j + 5], o[j + 6], o[j + 7], a[1], a[3], a[5], a[7], b[j]);
    }
  }
  f8d_merge16(e, o, t);
}

// a^2: the 28 off-diagonal products once, doubled with funnel shifts, plus the 8 squares as one carry chain
__device__ __forceinline__ void f8d_sqr_wide(const u32 (&a)[8], u32 (&t)[16]) {
  u32 e[16], o[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { e[i] = 0; o[i] = 0; }
  F8_ROW3_FIRST(e[2], e[3], e[4], e[5], e[6], e[7], a[2], a[4], a[6], a[0]);
  F8_ROW3_NC(e[4], e[5], e[6], e[7], e[8], e[9], a[3], a[5], a[7], a[1]);
  F8_ROW2_C(e[6], e[7], e[8], e[9], e[10], a[4], a[6], a[2]);
  F8_ROW2_NC(e[8], e[9], e[10], e[11], a[5], a[7], a[3]);
  F8_ROW1_C(e[10], e[11], e[12], a[6], a[4]);
  F8_ROW1_NC(e[12], e[13], a[7], a[5]);
  F8_ROW4_FIRST(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], a[1], a[3], a[5], a[7], a[0]);
  F8_ROW3_C(o[2], o[3], o[4], o[5], o[6], o[7], o[8], a[2], a[4], a[6], a[1]);
  F8_ROW3_NC(o[4], o[5], o[6], o[7], o[8], o[9], a[3], a[5], a[7], a[2]);
  F8_ROW2_C(o[6], o[7], o[8], o[9], o[10], a[4], a[6], a[3]);
  F8_ROW2_NC(o[8], o[9], o[10], o[11], a[5], a[7], a[4]);
  F8_ROW1_C(o[10], o[11], o[12], a[6], a[5]);
  F8_ROW1_NC(o[12], o[13], a[7], a[6]);
  u32 s[16];
  f8d_merge16(e, o, s);                              // s[0] = 0, s < 2^511
#pragma unroll
  for (int k = 15; k >= 1; k--) asm("shf.l.clamp.b32 %0, %1, %2, 1;" : "=r"(t[k]) : "r"(s[k - 1]), "r"(s[k]));
  t[0] = 0;
  asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\tmadc.hi.cc.u32 %1, %16, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %17, %17, %2;\n\tmadc.hi.cc.u32 %3, %17, %17, %3;\n\t"
      "madc.lo.cc.u32 %4, %18, %18, %4;\n\tmadc.hi.cc.u32 %5, %18, %18, %5;\n\t"
      "madc.lo.cc.u32 %6, %19, %19, %6;\n\tmadc.hi.cc.u32 %7, %19, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %20, %20, %8;\n\tmadc.hi.cc.u32 %9, %20, %20, %9;\n\t"
      "madc.lo.cc.u32 %10, %21, %21, %10;\n\tmadc.hi.cc.u32 %11, %21, %21, %11;\n\t"
      "madc.lo.cc.u32 %12, %22, %22, %12;\n\tmadc.hi.cc.u32 %13, %22, %22, %13;\n\t"
      "madc.lo.cc.u32 %14, %23, %23, %14;\n\tmadc.hi.u32 %15, %23, %23, %15;"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]),
        "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
}

// r = lo + top * (2^32 + 977) folded to 8 words; top = t8 + 2^32 t9 with t9 <= 1
__device__ __forceinline__ void f8d_fold9(const u32 (&lo)[8], u32 t8, u32 t9, u32 (&r)[8]) {
  u32 a0, a1, a2, c8;
  asm("mul.lo.u32 %0, %3, %5;\n\tmul.hi.u32 %1, %3, %5;\n\t"
      "add.cc.u32 %1, %1, %3;\n\taddc.u32 %2, %4, 0;\n\t"          // + t8 << 32, + t9 << 64
      "mad.lo.cc.u32 %1, %4, %5, %1;\n\taddc.u32 %2, %2, 0;"         // + t9 * 977 << 32
      : "=&r"(a0), "=&r"(a1), "=&r"(a2)
      : "r"(t8), "r"(t9), "r"(F8_C));
  asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, 0;\n\taddc.cc.u32 %4, %13, 0;\n\taddc.cc.u32 %5, %14, 0;\n\t"
      "addc.cc.u32 %6, %15, 0;\n\taddc.cc.u32 %7, %16, 0;\n\taddc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c8)
      : "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]), "r"(a0), "r"(a1), "r"(a2));
  // c8 = 1 only when the sum wrapped (r < 2^67): the last fold cannot carry past word 2
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\taddc.cc.u32 %1, %1, %3;\n\taddc.u32 %2, %2, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2])
      : "r"(c8), "r"(F8_C));
}

// small top (t8 < 2^20, e.g. a carry or the top word of a small multiple): one multiply-add instead of three words
__device__ __forceinline__ void f8d_fold_small(const u32 (&lo)[8], u32 t8, u32 (&r)[8]) {
  u32 c8;
  asm("mad.lo.cc.u32 %0, %17, %18, %9;\n\taddc.cc.u32 %1, %10, %17;\n\taddc.cc.u32 %2, %11, 0;\n\t"
      "addc.cc.u32 %3, %12, 0;\n\taddc.cc.u32 %4, %13, 0;\n\taddc.cc.u32 %5, %14, 0;\n\t"
      "addc.cc.u32 %6, %15, 0;\n\taddc.cc.u32 %7, %16, 0;\n\taddc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c8)
      : "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]), "r"(t8), "r"(F8_C));
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\taddc.cc.u32 %1, %1, %3;\n\taddc.u32 %2, %2, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2])
      : "r"(c8), "r"(F8_C));
}

__device__ __forceinline__ void f8d_reduce16(const u32 (&t)[16], u32 (&r)[8]) {
  u32 e[9], o[9];
  // even words of H times 977 on top of L (aligned pairs)
  asm("mad.lo.cc.u32 %0, %9, %13, %14;\n\tmadc.hi.cc.u32 %1, %9, %13, %15;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %16;\n\tmadc.hi.cc.u32 %3, %10, %13, %17;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %18;\n\tmadc.hi.cc.u32 %5, %11, %13, %19;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %20;\n\tmadc.hi.cc.u32 %7, %12, %13, %21;\n\taddc.u32 %8, 0, 0;"
      : "=&r"(e[0]), "=&r"(e[1]), "=&r"(e[2]), "=&r"(e[3]), "=&r"(e[4]), "=&r"(e[5]), "=&r"(e[6]), "=&r"(e[7]), "=&r"(e[8])
      : "r"(t[8]), "r"(t[10]), "r"(t[12]), "r"(t[14]), "r"(F8_C), "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]),
        "r"(t[5]), "r"(t[6]), "r"(t[7]));
  // odd words of H times 977 plus H itself (the 2^32 term) in the frame shifted by one word
  asm("mad.lo.cc.u32 %0, %9, %13, %14;\n\tmadc.hi.cc.u32 %1, %9, %13, %15;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %16;\n\tmadc.hi.cc.u32 %3, %10, %13, %17;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %18;\n\tmadc.hi.cc.u32 %5, %11, %13, %19;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %20;\n\tmadc.hi.cc.u32 %7, %12, %13, %21;\n\taddc.u32 %8, 0, 0;"
      : "=&r"(o[0]), "=&r"(o[1]), "=&r"(o[2]), "=&r"(o[3]), "=&r"(o[4]), "=&r"(o[5]), "=&r"(o[6]), "=&r"(o[7]), "=&r"(o[8])
      : "r"(t[9]), "r"(t[11]), "r"(t[13]), "r"(t[15]), "r"(F8_C), "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]),
        "r"(t[13]), "r"(t[14]), "r"(t[15]));
  u32 R[8], t8, t9;
  R[0] = e[0];
  asm("add.cc.u32 %0, %1, %2;" : "=r"(R[1]) : "r"(e[1]), "r"(o[0]));
#pragma unroll
  for (int k = 2; k < 8; k++) asm("addc.cc.u32 %0, %1, %2;" : "=r"(R[k]) : "r"(e[k]), "r"(o[k - 1]));
  asm("addc.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=r"(t8), "=r"(t9) : "r"(e[8]), "r"(o[7]), "r"(o[8]));
  f8d_fold9(R, t8, t9, r);
}

__device__ __forceinline__ void f8d_add(const u32 (&a)[8], const u32 (&b)[8], u32 (&r)[8]) {
  u32 s[8], c;
  asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\taddc.cc.u32 %4, %13, %21;\n\taddc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\taddc.cc.u32 %7, %16, %24;\n\taddc.u32 %8, 0, 0;"
      : "=&r"(s[0]), "=&r"(s[1]), "=&r"(s[2]), "=&r"(s[3]), "=&r"(s[4]), "=&r"(s[5]), "=&r"(s[6]), "=&r"(s[7]), "=&r"(c)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  f8d_fold_small(s, c, r);
}

__device__ __forceinline__ void f8d_sub(const u32 (&a)[8], const u32 (&b)[8], u32 (&r)[8]) {
  u32 br;
  asm("sub.cc.u32 %0, %9, %17;\n\tsubc.cc.u32 %1, %10, %18;\n\tsubc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\tsubc.cc.u32 %4, %13, %21;\n\tsubc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\tsubc.cc.u32 %7, %16, %24;\n\tsubc.u32 %8, 0, 0;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(br)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  // br = 0xffffffff on borrow: subtract 2^32 + 977 (i.e. add p, mod 2^256); a second borrow is possible, a third is not
  u32 m0 = br & F8_C, m1 = br & 1u;
  asm("sub.cc.u32 %0, %0, %9;\n\tsubc.cc.u32 %1, %1, %10;\n\tsubc.cc.u32 %2, %2, 0;\n\t"
      "subc.cc.u32 %3, %3, 0;\n\tsubc.cc.u32 %4, %4, 0;\n\tsubc.cc.u32 %5, %5, 0;\n\t"
      "subc.cc.u32 %6, %6, 0;\n\tsubc.cc.u32 %7, %7, 0;\n\tsubc.u32 %8, 0, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=r"(br)
      : "r"(m0), "r"(m1));
  m0 = br & F8_C;
  m1 = br & 1u;
  asm("sub.cc.u32 %0, %0, %8;\n\tsubc.cc.u32 %1, %1, %9;\n\tsubc.cc.u32 %2, %2, 0;\n\t"
      "subc.cc.u32 %3, %3, 0;\n\tsubc.cc.u32 %4, %4, 0;\n\tsubc.cc.u32 %5, %5, 0;\n\t"
      "subc.cc.u32 %6, %6, 0;\n\tsubc.u32 %7, %7, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
      : "r"(m0), "r"(m1));
}

template <int K>
__device__ __forceinline__ void f8d_mul_small(const u32 (&a)[8], u32 (&r)[8]) {
  u32 e[8], o[8], s[8], t8;
  const u32 k = (u32)K;
  F8_ROW4_FIRST(e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], a[0], a[2], a[4], a[6], k);
  F8_ROW4_FIRST(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], a[1], a[3], a[5], a[7], k);
  s[0] = e[0];
  asm("add.cc.u32 %0, %1, %2;" : "=r"(s[1]) : "r"(e[1]), "r"(o[0]));
#pragma unroll
  for (int i = 2; i < 8; i++) asm("addc.cc.u32 %0, %1, %2;" : "=r"(s[i]) : "r"(e[i]), "r"(o[i - 1]));
  asm("addc.u32 %0, %1, 0;" : "=r"(t8) : "r"(o[7]));
  f8d_fold_small(s, t8, r);
}

__device__ __forceinline__ void f8d_canon(u32 (&r)[8]) {
  u32 t[8], c;
  asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, 1;\n\taddc.cc.u32 %2, %11, 0;\n\t"
      "addc.cc.u32 %3, %12, 0;\n\taddc.cc.u32 %4, %13, 0;\n\taddc.cc.u32 %5, %14, 0;\n\t"
      "addc.cc.u32 %6, %15, 0;\n\taddc.cc.u32 %7, %16, 0;\n\taddc.u32 %8, 0, 0;"
      : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(c)
      : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(F8_C));
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = c ? t[i] : r[i];
}
#endif  // __CUDA_ARCH__

// ------------------------------------------------------------------------------------------
// public operations (device: PTX bodies; host: portable bodies)
P2E_HD void f8_mul(const u32 (&a)[8], const u32 (&b)[8], u32 (&r)[8]) {
  u32 t[16];
#ifdef __CUDA_ARCH__
  f8d_mul_wide(a, b, t);
  f8d_reduce16(t, r);
#else
  f8p_mul_wide(a, b, t);
  f8p_reduce16(t, r);
#endif
}
P2E_HD void f8_mul_wide(const u32 (&a)[8], const u32 (&b)[8], u32 (&t)[16]) {
#ifdef __CUDA_ARCH__
  f8d_mul_wide(a, b, t);
#else
  f8p_mul_wide(a, b, t);
#endif
}
P2E_HD void f8_sqr(const u32 (&a)[8], u32 (&r)[8]) {
  u32 t[16];
#ifdef __CUDA_ARCH__
  f8d_sqr_wide(a, t);
  f8d_reduce16(t, r);
#else
  f8p_mul_wide(a, a, t);
  f8p_reduce16(t, r);
#endif
}
P2E_HD void f8_add(const u32 (&a)[8], const u32 (&b)[8], u32 (&r)[8]) {
#ifdef __CUDA_ARCH__
  f8d_add(a, b, r);
#else
  f8p_add(a, b, r);
#endif
}
P2E_HD void f8_sub(const u32 (&a)[8], const u32 (&b)[8], u32 (&r)[8]) {
#ifdef __CUDA_ARCH__
  f8d_sub(a, b, r);
#else
  f8p_sub(a, b, r);
#endif
}
template <int K>
P2E_HD void f8_mul_small(const u32 (&a)[8], u32 (&r)[8]) {
#ifdef __CUDA_ARCH__
  f8d_mul_small<K>(a, r);
#else
  f8p_mul_small<K>(a, r);
#endif
}
P2E_HD void f8_canon(u32 (&r)[8]) {
#ifdef __CUDA_ARCH__
  f8d_canon(r);
#else
  f8p_canon(r);
#endif
}
P2E_HD void f8_copy(const u32 (&a)[8], u32 (&r)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = a[i];
}
P2E_HD void f8_one(u32 (&r)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = (i == 0);
}
// zero as a field element (weak values: 0 or p)
P2E_HD bool f8_is_zero(const u32 (&a)[8]) {
  u32 t[8];
  f8_copy(a, t);
  f8_canon(t);
  u32 o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= t[i];
  return o == 0;
}

// 9 x 29-bit limbs (value < 2^256) <-> 8 x 32-bit words
P2E_HD void f8_from_l29(const u32 (&l)[NL], u32 (&w)[8]) {
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int sh = 3 * k;                               // word k starts at bit 32 k = 29 k + 3 k
    w[k] = (l[k] >> sh) | (l[k + 1] << (LB - sh));
  }
}
P2E_HD void f8_to_l29(const u32 (&w)[8], u32 (&l)[NL]) {
  l[0] = w[0] & LMASK;
#pragma unroll
  for (int i = 1; i < 8; i++) {
    const int sh = 32 - 3 * i;                          // limb i starts at bit 29 i = 32 (i - 1) + sh
    l[i] = ((w[i - 1] >> sh) | (w[i] << (32 - sh))) & LMASK;
  }
  l[8] = w[7] >> 8;
}
// weak words -> canonical 29-bit limbs
P2E_HD void f8_to_canon_l29(const u32 (&w)[8], u32 (&l)[NL]) {
  u32 t[8];
  f8_copy(w, t);
  f8_canon(t);
  f8_to_l29(t, l);
}

// ------------------------------------------------------------------------------------------
// Fast MulNonnativeGenerator + CheckSumGenerator core for the chain expansion (K4).
// words of a 522-bit value given as 18 normalised 29-bit limbs
template <int NW>
P2E_HD void l29_to_words(const u32 (&T)[18], u32 (&W)[NW]) {
#pragma unroll
  for (int k = 0; k < NW; k++) {
    const int a = (32 * k) / 29, s = 32 * k - 29 * a;
    u32 w = T[a] >> s;
    if (a + 1 < 18) w |= T[a + 1] << (29 - s);
    if (s > 26 && a + 2 < 18) w |= T[a + 2] << (58 - s);
    W[k] = w;
  }
}

#ifdef __CUDA_ARCH__
// PTX body of f8_divmod16 (same function as the portable one below: r canonical, q exact)
__device__ __forceinline__ u32 f8d_divmod16(const u32 (&t)[16], u32 (&r)[8], u32 (&q)[8]) {
  u32 e[9], o[9];
  asm("mad.lo.cc.u32 %0, %9, %13, %14;\n\tmadc.hi.cc.u32 %1, %9, %13, %15;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %16;\n\tmadc.hi.cc.u32 %3, %10, %13, %17;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %18;\n\tmadc.hi.cc.u32 %5, %11, %13, %19;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %20;\n\tmadc.hi.cc.u32 %7, %12, %13, %21;\n\taddc.u32 %8, 0, 0;"
      : "=&r"(e[0]), "=&r"(e[1]), "=&r"(e[2]), "=&r"(e[3]), "=&r"(e[4]), "=&r"(e[5]), "=&r"(e[6]), "=&r"(e[7]), "=&r"(e[8])
      : "r"(t[8]), "r"(t[10]), "r"(t[12]), "r"(t[14]), "r"(F8_C), "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]),
        "r"(t[5]), "r"(t[6]), "r"(t[7]));
  asm("mad.lo.cc.u32 %0, %9, %13, %14;\n\tmadc.hi.cc.u32 %1, %9, %13, %15;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %16;\n\tmadc.hi.cc.u32 %3, %10, %13, %17;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %18;\n\tmadc.hi.cc.u32 %5, %11, %13, %19;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %20;\n\tmadc.hi.cc.u32 %7, %12, %13, %21;\n\taddc.u32 %8, 0, 0;"
      : "=&r"(o[0]), "=&r"(o[1]), "=&r"(o[2]), "=&r"(o[3]), "=&r"(o[4]), "=&r"(o[5]), "=&r"(o[6]), "=&r"(o[7]), "=&r"(o[8])
      : "r"(t[9]), "r"(t[11]), "r"(t[13]), "r"(t[15]), "r"(F8_C), "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]),
        "r"(t[13]), "r"(t[14]), "r"(t[15]));
  u32 R[8], t8, t9;
  R[0] = e[0];
  asm("add.cc.u32 %0, %1, %2;" : "=r"(R[1]) : "r"(e[1]), "r"(o[0]));
#pragma unroll
  for (int k = 2; k < 8; k++) asm("addc.cc.u32 %0, %1, %2;" : "=r"(R[k]) : "r"(e[k]), "r"(o[k - 1]));
  asm("addc.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=r"(t8), "=r"(t9) : "r"(e[8]), "r"(o[7]), "r"(o[8]));
  // S = R + (t8 + 2^32 t9)(2^32 + 977), then the wrap bit c8 once more (cannot carry past word 2)
  u32 a0, a1, a2, c8, S[8];
  asm("mul.lo.u32 %0, %3, %5;\n\tmul.hi.u32 %1, %3, %5;\n\tadd.cc.u32 %1, %1, %3;\n\taddc.u32 %2, %4, 0;\n\t"
      "mad.lo.cc.u32 %1, %4, %5, %1;\n\taddc.u32 %2, %2, 0;"
      : "=&r"(a0), "=&r"(a1), "=&r"(a2)
      : "r"(t8), "r"(t9), "r"(F8_C));
  asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, 0;\n\taddc.cc.u32 %4, %13, 0;\n\taddc.cc.u32 %5, %14, 0;\n\t"
      "addc.cc.u32 %6, %15, 0;\n\taddc.cc.u32 %7, %16, 0;\n\taddc.u32 %8, 0, 0;"
      : "=r"(S[0]), "=r"(S[1]), "=r"(S[2]), "=r"(S[3]), "=r"(S[4]), "=r"(S[5]), "=r"(S[6]), "=r"(S[7]), "=r"(c8)
      : "r"(R[0]), "r"(R[1]), "r"(R[2]), "r"(R[3]), "r"(R[4]), "r"(R[5]), "r"(R[6]), "r"(R[7]), "r"(a0), "r"(a1), "r"(a2));
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\taddc.cc.u32 %1, %1, %3;\n\taddc.u32 %2, %2, 0;"
      : "+r"(S[0]), "+r"(S[1]), "+r"(S[2])
      : "r"(c8), "r"(F8_C));
  // final conditional subtraction: S + (2^32 + 977) carries out iff S >= p
  u32 w[8], fin;
  asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, 1;\n\taddc.cc.u32 %2, %11, 0;\n\t"
      "addc.cc.u32 %3, %12, 0;\n\taddc.cc.u32 %4, %13, 0;\n\taddc.cc.u32 %5, %14, 0;\n\t"
      "addc.cc.u32 %6, %15, 0;\n\taddc.cc.u32 %7, %16, 0;\n\taddc.u32 %8, 0, 0;"
      : "=&r"(w[0]), "=&r"(w[1]), "=&r"(w[2]), "=&r"(w[3]), "=&r"(w[4]), "=&r"(w[5]), "=&r"(w[6]), "=&r"(w[7]), "=&r"(fin)
      : "r"(S[0]), "r"(S[1]), "r"(S[2]), "r"(S[3]), "r"(S[4]), "r"(S[5]), "r"(S[6]), "r"(S[7]), "r"(F8_C));
#pragma unroll
  for (int k = 0; k < 8; k++) r[k] = fin ? w[k] : S[k];
  // q = H + t8 + 2^32 t9 + c8 + fin
  u32 s0, s1, qov;
  const u32 small = c8 + fin;
  asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, 0;" : "=r"(s0), "=r"(s1) : "r"(t8), "r"(t9), "r"(small));
  asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, 0;\n\t"
      "addc.cc.u32 %3, %12, 0;\n\taddc.cc.u32 %4, %13, 0;\n\taddc.cc.u32 %5, %14, 0;\n\t"
      "addc.cc.u32 %6, %15, 0;\n\taddc.cc.u32 %7, %16, 0;\n\taddc.u32 %8, 0, 0;"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(qov)
      : "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]), "r"(s0), "r"(s1));
  return qov;
}
#endif

// (q, r) = divmod(W, p) for W < 2^512 with q < 2^256 (guaranteed when both factors are < p)
P2E_HD u32 f8_divmod16(const u32 (&t)[16], u32 (&r)[8], u32 (&q)[8]) {
#ifdef __CUDA_ARCH__
  return f8d_divmod16(t, r, q);
#else
  u32 R[8];
  u64 c = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u64 v = c + t[k] + (u64)t[8 + k] * F8_C + (k ? t[8 + k - 1] : 0u);
    R[k] = (u32)v;
    c = v >> 32;
  }
  const u64 top = c + t[15];
  // S = R + top * (2^32 + 977)
  u32 S[8];
  {
    u64 cc = 0;
    u64 add0 = (top & 0xffffffffull) * F8_C;
    u64 add1 = (top >> 32) * F8_C + (top & 0xffffffffull);
    u64 add2 = top >> 32;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      u64 v = cc + R[k] + (k == 0 ? (add0 & 0xffffffffull) : k == 1 ? (add0 >> 32) + add1 : k == 2 ? add2 : 0);
      S[k] = (u32)v;
      cc = v >> 32;
    }
    c = cc;   // c8
  }
  const u64 c8 = c;
  {
    u64 v = (u64)S[0] + c8 * F8_C;
    S[0] = (u32)v;
    v = (u64)S[1] + c8 + (v >> 32);
    S[1] = (u32)v;
    S[2] += (u32)(v >> 32);
  }
  // final conditional subtraction: S + (2^32 + 977) carries out iff S >= p
  u32 w[8];
  u64 cc = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u64 v = cc + S[k] + (k == 0 ? F8_C : k == 1 ? 1u : 0u);
    w[k] = (u32)v;
    cc = v >> 32;
  }
  const u64 fin = cc;
#pragma unroll
  for (int k = 0; k < 8; k++) r[k] = fin ? w[k] : S[k];
  // q = H + top + c8 + fin
  u64 small = top + c8 + fin;
  cc = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u64 v = cc + t[8 + k] + (k == 0 ? (small & 0xffffffffull) : k == 1 ? (small >> 32) : 0);
    q[k] = (u32)v;
    cc = v >> 32;
  }
  return (u32)cc;     // quotient overflow (never for factors < p)
#endif
}

// A8 + A9 for factors below 2^256 with q < 2^256 (chain kernel K4: every operand is a canonical generator output).
// r, q: 29-bit limbs; v: the 17 un-carried check_sum columns as signed integers; blo / tops: the CheckSum carries
// b_i = c_i + 2^33 as low words + 2-bit tops (the compact host format).
P2E_HD u32 mulw_p_fast(const u32 (&x)[NL], const u32 (&y)[NL], u32 (&r)[NL], u32 (&q)[NL], i64 (&v)[17], u32 (&blo)[16], u32& tops) {
  u64 X[17];
  mul_cols(x, y, X);
  u32 W[16], r8[8], q8[8];
  {
    u32 T[18];
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < 17; i++) {
      u64 s = X[i] + c;
      T[i] = (u32)s & LMASK;
      c = s >> LB;
    }
    T[17] = (u32)c;
    l29_to_words<16>(T, W);
  }
  u32 bad = f8_divmod16(W, r8, q8);
  f8_to_l29(r8, r);
  f8_to_l29(q8, q);
  // check_sum columns: m = [m0, m1, M, M, M, M, M, M, m8]
  constexpr u32 m0 = 0x1ffffc2fu, m1 = 0x1ffffff7u, m8 = 0x00ffffffu;
  u32 S = 0;
#pragma unroll
  for (int i = 0; i < 17; i++) {
    if (i >= 2 && i - 2 <= 8) S += q[i - 2];
    if (i >= 8 && i - 8 <= 8) S -= q[i - 8];
    u64 a = (i < NL ? (u64)r[i] : 0) - X[i];
    if (i < NL) a += (u64)q[i] * m0;
    if (i >= 1 && i - 1 < NL) a += (u64)q[i - 1] * m1;
    if (i >= 8) a += (u64)q[i - 8] * m8;
    a += (u64)S * LMASK;
    v[i] = (i64)a;
  }
  // exact carries c_i = (v_i + c_{i-1}) >> 29 (the partial sums are multiples of 2^29 because x y = q p + r holds as
  // integers; the last column closes the chain, checked once); b_i = c_i + 2^33 as low word + two top bits
  i64 c = 0;
  u32 tp = 0, hacc = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const i64 s = v[i] + c;
    c = s >> LB;
    blo[i] = (u32)c;                               // + 2^33 does not touch the low word
    const u32 h = (u32)((u64)c >> 32) + 2u;        // top word of c + 2^33: 0..3 for |c| <= 2^33
    hacc |= h;
    tp += h << (2 * i);
  }
  tops = tp;
  bad |= (hacc >> 2) | (u32)((u64)(v[16] + c) != 0);
  return bad ? ERR_INTERNAL : 0u;
}


}  // namespace p2e
