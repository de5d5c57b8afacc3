#include "../plonky2_ecdsa_b200/csrc/p2e_f8.cuh"
using namespace p2e;

P2E_HD u32 mulw_p_fast2(const u32 (&x)[NL], const u32 (&y)[NL], u32 (&r)[NL], u32 (&q)[NL], i64 (&v)[17], u32 (&blo)[16], u32& tops) {
  // integer side: words
  u32 x8[8], y8[8], W[16], r8[8], q8[8];
  f8_from_l29(x, x8);
  f8_from_l29(y, y8);
  f8_mul_wide(x8, y8, W);
  u32 bad = f8_divmod16(W, r8, q8);
  f8_to_l29(r8, r);
  f8_to_l29(q8, q);
  // column side
  u64 X[17];
  mul_cols(x, y, X);
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
  u32 tp = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    i64 s = v[i] + c;
    c = s >> LB;
    u64 o = (u64)(c + ((i64)1 << 33));
    blo[i] = (u32)o;
    tp |= (((u32)(o >> 32)) & 3u) << (2 * i);
  }
  tops = tp;
  bad |= (u32)((u64)(v[16] + c) != 0);
  return bad ? ERR_INTERNAL : 0u;
}
extern "C" __global__ void k_mulw2(const u32* a, const u32* b, u32* o, u32* err) {
  u32 x[9], y[9];
  for (int i = 0; i < 9; i++) { x[i] = a[i*32+threadIdx.x]; y[i] = b[i*32+threadIdx.x]; }
  u32 r[9], q[9], blo[16], tops; i64 v[17];
  u32 e = mulw_p_fast2(x, y, r, q, v, blo, tops);
  for (int i = 0; i < 9; i++) { o[i*32+threadIdx.x] = r[i]; o[(9+i)*32+threadIdx.x] = q[i]; }
  for (int i = 0; i < 16; i++) o[(18+i)*32+threadIdx.x] = blo[i];
  o[34*32+threadIdx.x] = tops;
  err[threadIdx.x] = e;
}
extern "C" __global__ void k_mulw1(const u32* a, const u32* b, u32* o, u32* err) {
  u32 x[9], y[9];
  for (int i = 0; i < 9; i++) { x[i] = a[i*32+threadIdx.x]; y[i] = b[i*32+threadIdx.x]; }
  u32 r[9], q[9], blo[16], tops; i64 v[17];
  u32 e = mulw_p_fast(x, y, r, q, v, blo, tops);
  for (int i = 0; i < 9; i++) { o[i*32+threadIdx.x] = r[i]; o[(9+i)*32+threadIdx.x] = q[i]; }
  for (int i = 0; i < 16; i++) o[(18+i)*32+threadIdx.x] = blo[i];
  o[34*32+threadIdx.x] = tops;
  err[threadIdx.x] = e;
}
