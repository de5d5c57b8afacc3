// TEST INFRASTRUCTURE ONLY.  Compiles the __host__ __device__ arithmetic cores of
// plonky2_ecdsa_b200/csrc/p2e_field.cuh as plain host C++ so that the CPU-only test tier can
// check the exact code the CUDA kernels run against the oracle without a GPU.  It is never
// loaded by the product package (which has no CPU path).
#define __host__
#define __device__
#define __forceinline__ inline __attribute__((always_inline))
#include <stddef.h>
#include "../../plonky2_ecdsa_b200/csrc/p2e_field.cuh"
#include "../../plonky2_ecdsa_b200/csrc/p2e_f8.cuh"

using namespace p2e;

template <class MOD>
static void mul_one(const u64* x, const u64* y, u64* out, u32* err) {
  u64 xw[NL], yw[NL];
  for (int l = 0; l < NL; l++) { xw[l] = x[l]; yw[l] = y[l]; }
  u32 xa[NL], ya[NL], r[NL], q[NL];
  u64 cs[17], b[16];
  u32 e = load_mul_wires(xw, xa) | load_mul_wires(yw, ya);
  if (e == 0) {
    e = mul_witness_core<MOD>(xa, ya, r, q, cs, b);
    for (int l = 0; l < NL; l++) { out[l] = r[l]; out[9 + l] = q[l]; }
    for (int l = 0; l < 17; l++) out[18 + l] = cs[l];
    for (int l = 0; l < 16; l++) out[35 + l] = b[l];
  } else {
    for (int l = 0; l < 51; l++) out[l] = 0;
  }
  *err = e;
}

template <class MOD>
static void addsub_one(int is_sub, const u64* a, const u64* b, u64* out, u32* err) {
  u64 aw[NL], bw[NL];
  for (int l = 0; l < NL; l++) { aw[l] = a[l]; bw[l] = b[l]; }
  u32 ac[NL], bc[NL], r[NL];
  u32 e = load_canon<MOD>(aw, ac) | load_canon<MOD>(bw, bc);
  u32 ov = 0;
  if (e == 0) ov = is_sub ? sub_witness_core<MOD>(ac, bc, r) : add_witness_core<MOD>(ac, bc, r);
  else set_zero9(r);
  for (int l = 0; l < NL; l++) out[l] = r[l];
  out[9] = ov;
  *err = e;
}

template <class MOD>
static void inv_one(const u64* x, u64* out, u32* err) {
  u64 w[NL];
  for (int l = 0; l < NL; l++) w[l] = x[l];
  u32 xc[NL], inv[NL], div[NL];
  set_zero9(inv);
  set_zero9(div);
  u32 e = load_canon<MOD>(w, xc);
  if (e == 0 && is_zero9(xc)) e = ERR_INVERSE_ZERO;
  if (e == 0) {
    inv_mod<MOD>(xc, inv);
    e |= inv_div_core<MOD>(xc, inv, div);
  }
  for (int l = 0; l < NL; l++) { out[l] = div[l]; out[9 + l] = inv[l]; }
  *err = e;
}

extern "C" {

// x, y: [n][9]; out: [n][51]; err: [n]
void hc_mul_witness(int mod, size_t n, const u64* x, const u64* y, u64* out, u32* err) {
  for (size_t i = 0; i < n; i++) {
    if (mod == 0) mul_one<ModP>(x + 9 * i, y + 9 * i, out + 51 * i, err + i);
    else mul_one<ModN>(x + 9 * i, y + 9 * i, out + 51 * i, err + i);
  }
}

void hc_check_sum(size_t n, const u64* a, u64* b, u32* err) {
  for (size_t i = 0; i < n; i++) {
    u64 aw[17], bw[16];
    for (int l = 0; l < 17; l++) aw[l] = a[17 * i + l];
    err[i] = check_sum_core(aw, bw);
    for (int l = 0; l < 16; l++) b[16 * i + l] = bw[l];
  }
}

void hc_addsub(int mod, int is_sub, size_t n, const u64* a, const u64* b, u64* out, u32* err) {
  for (size_t i = 0; i < n; i++) {
    if (mod == 0) addsub_one<ModP>(is_sub, a + 9 * i, b + 9 * i, out + 10 * i, err + i);
    else addsub_one<ModN>(is_sub, a + 9 * i, b + 9 * i, out + 10 * i, err + i);
  }
}

void hc_inv(int mod, size_t n, const u64* x, u64* out, u32* err) {
  for (size_t i = 0; i < n; i++) {
    if (mod == 0) inv_one<ModP>(x + 9 * i, out + 18 * i, err + i);
    else inv_one<ModN>(x + 9 * i, out + 18 * i, err + i);
  }
}

void hc_glv(size_t n, const u64* k, u64* out, u32* err) {
  for (size_t i = 0; i < n; i++) {
    u64 w[NL];
    for (int l = 0; l < NL; l++) w[l] = k[9 * i + l];
    u32 kc[NL], k1[NL], k2[NL], n1 = 0, n2 = 0;
    u32 e = load_canon<ModN>(w, kc);
    set_zero9(k1);
    set_zero9(k2);
    if (e == 0) glv_decompose_core(kc, k1, k2, n1, n2);
    for (int l = 0; l < 5; l++) { out[12 * i + l] = k1[l]; out[12 * i + 5 + l] = k2[l]; }
    out[12 * i + 10] = n1;
    out[12 * i + 11] = n2;
    u32 hi = 0;
    for (int l = 5; l < NL; l++) hi |= k1[l] | k2[l];
    if (hi) e |= ERR_LIMBS_OVERFLOW;
    err[i] = e;
  }
}

// lazy field: ops on raw signed limbs; out = canonical result + raw lazy limbs (for bound checks)
// op: 0 mul, 1 sqr, 2 add, 3 sub, 4 lin(ka,kb), 5 inv
void hc_lazy(int op, int ka, int kb, size_t n, const int32_t* a, const int32_t* b, u32* canon, int32_t* raw) {
  for (size_t i = 0; i < n; i++) {
    i32 x[NL], y[NL], r[NL];
    for (int l = 0; l < NL; l++) { x[l] = a[9 * i + l]; y[l] = b[9 * i + l]; }
    if (op == 0) lz_mul(x, y, r);
    else if (op == 1) lz_sqr(x, r);
    else if (op == 2) lz_add(x, y, r);
    else if (op == 3) lz_sub(x, y, r);
    else if (op == 4) lz_lin(x, ka, y, kb, r);
    else lz_inv(x, r);
    u32 c[NL];
    lz_canon(r, c);
    for (int l = 0; l < NL; l++) { canon[9 * i + l] = c[l]; raw[9 * i + l] = r[l]; }
  }
}

// saturated 8 x 32-bit field (p2e_f8.cuh, portable bodies): out = raw weak words
// op: 0 mul, 1 sqr, 2 add, 3 sub, 4 mul_small<2>, 5 <3>, 6 <4>, 7 <8>, 8 canon, 9 from_l29 (a = 9 limbs in a[0..8]) , 10 to_l29 (out 9 words)
void hc_f8(int op, size_t n, const u32* a, const u32* b, u32* out) {
  for (size_t i = 0; i < n; i++) {
    u32 x[8], y[8], r[8];
    for (int l = 0; l < 8; l++) { x[l] = a[9 * i + l]; y[l] = b[9 * i + l]; r[l] = 0; }
    if (op == 0) f8_mul(x, y, r);
    else if (op == 1) f8_sqr(x, r);
    else if (op == 2) f8_add(x, y, r);
    else if (op == 3) f8_sub(x, y, r);
    else if (op == 4) f8_mul_small<2>(x, r);
    else if (op == 5) f8_mul_small<3>(x, r);
    else if (op == 6) f8_mul_small<4>(x, r);
    else if (op == 7) f8_mul_small<8>(x, r);
    else if (op == 8) { f8_copy(x, r); f8_canon(r); }
    else if (op == 9) { u32 l9[NL]; for (int l = 0; l < NL; l++) l9[l] = a[9 * i + l]; f8_from_l29(l9, r); }
    if (op == 10) { u32 l9[NL]; f8_to_l29(x, l9); for (int l = 0; l < NL; l++) out[9 * i + l] = l9[l]; }
    else { for (int l = 0; l < 8; l++) out[9 * i + l] = r[l]; out[9 * i + 8] = 0; }
  }
}

// fast mul-witness core of the chain expansion (ModP, factors < 2^256): out = r[9] q[9] cs[17] b[16] like hc_mul_witness
void hc_mulw_fast(size_t n, const u64* x, const u64* y, u64* out, u32* err) {
  for (size_t i = 0; i < n; i++) {
    u32 xa[NL], ya[NL], r[NL], q[NL], blo[16], tops;
    i64 v[17];
    for (int l = 0; l < NL; l++) { xa[l] = (u32)x[9 * i + l]; ya[l] = (u32)y[9 * i + l]; }
    err[i] = mulw_p_fast(xa, ya, r, q, v, blo, tops);
    for (int l = 0; l < NL; l++) { out[51 * i + l] = r[l]; out[51 * i + 9 + l] = q[l]; }
    for (int l = 0; l < 17; l++) out[51 * i + 18 + l] = gl_from_i64(v[l]);
    for (int l = 0; l < 16; l++) out[51 * i + 35 + l] = (u64)blo[l] | ((u64)((tops >> (2 * l)) & 3u) << 32);
  }
}

// division-step inverse (p2e_modinv.cuh) on 8 words x < m: out = x^-1 mod m; ladder != 0 -> the Fermat / Montgomery ladder
// on the same value (through 29-bit limbs), the algorithm the reference runs
void hc_modinv(int mod, int ladder, size_t n, const u32* a, u32* out) {
  for (size_t i = 0; i < n; i++) {
    u32 x[8], r[8];
    for (int l = 0; l < 8; l++) x[l] = a[8 * i + l];
    if (!ladder) {
      if (mod == 0) modinv_words<ModP>(x, r); else modinv_words<ModN>(x, r);
    } else {
      u32 l9[NL], o9[NL];
      f8_to_l29(x, l9);
      if (mod == 0) inv_mod_ladder<ModP>(l9, o9); else inv_mod_ladder<ModN>(l9, o9);
      f8_from_l29(o9, r);
    }
    for (int l = 0; l < 8; l++) out[8 * i + l] = r[l];
  }
}

u64 hc_gl_mul(u64 a, u64 b) { return gl_mul(a, b); }
u64 hc_gl_add(u64 a, u64 b) { return gl_add(a, b); }
u64 hc_gl_sub(u64 a, u64 b) { return gl_sub(a, b); }
u64 hc_gl_reduce128(u64 lo, u64 hi) { return gl_reduce128(lo, hi); }

}  // extern "C"
