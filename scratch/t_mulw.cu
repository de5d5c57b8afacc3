#include "../plonky2_ecdsa_b200/csrc/p2e_f8.cuh"
using namespace p2e;

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

// (q, r) = divmod(W, p) for W < 2^512 with q < 2^256 (guaranteed when both factors are < p); portable body
P2E_HD u32 f8p_divmod16(const u32 (&t)[16], u32 (&r)[8], u32 (&q)[8]) {
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
}

// A8 + A9 for factors below 2^256 with q < 2^256 (chain kernel K4: every operand is a canonical generator output).
// r, q: 29-bit limbs; v: the 17 un-carried check_sum columns as signed integers; blo / tops: the CheckSum carries
// b_i = c_i + 2^33 as low words + 2-bit tops (the compact host format).
P2E_HD u32 mulw_p_fast(const u32 (&x)[NL], const u32 (&y)[NL], u32 (&r)[NL], u32 (&q)[NL], i64 (&v)[17], u32 (&blo)[16], u32& tops) {
  u64 X[17];
  mul_cols(x, y, X);
  u32 T[18];
  {
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < 17; i++) {
      u64 s = X[i] + c;
      T[i] = (u32)s & LMASK;
      c = s >> LB;
    }
    T[17] = (u32)c;
  }
  u32 W[16], r8[8], q8[8];
  l29_to_words<16>(T, W);
  u32 bad = f8p_divmod16(W, r8, q8);
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
  i64 c = 0;
  u32 low = 0, tp = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    i64 s = v[i] + c;
    low |= (u32)s;
    c = s >> LB;
    u64 o = (u64)(c + ((i64)1 << 33));
    blo[i] = (u32)o;
    tp |= ((u32)(o >> 32)) << (2 * i);     // o < 2^34 for in-range carries (checked below through bit growth)
    bad |= (u32)(o >> 34);
  }
  tops = tp;
  bad |= (low & LMASK) | (u32)((u64)(v[16] + c) != 0);
  return bad ? ERR_INTERNAL : 0u;
}

extern "C" __global__ void k_mulw(const u32* a, const u32* b, u32* o, u32* err) {
  u32 x[9], y[9];
  for (int i = 0; i < 9; i++) { x[i] = a[i*32+threadIdx.x]; y[i] = b[i*32+threadIdx.x]; }
  u32 r[9], q[9], blo[16], tops; i64 v[17];
  u32 e = mulw_p_fast(x, y, r, q, v, blo, tops);
  for (int i = 0; i < 9; i++) { o[i*32+threadIdx.x] = r[i]; o[(9+i)*32+threadIdx.x] = q[i]; }
  for (int i = 0; i < 16; i++) o[(18+i)*32+threadIdx.x] = blo[i];
  o[34*32+threadIdx.x] = tops;
  err[threadIdx.x] = e;
}
