/* CPU ORACLE (test infrastructure / CPU baseline, NOT product code).
 *
 * Plain-C restatement of the reference's CPU algorithms for the hot path, in the reference's own
 * algorithm class: little-endian base-2^32 big integers (num::BigUint's representation),
 * schoolbook multiplication, Knuth long division, convert_base repacking, Goldilocks field
 * arithmetic with a field division per carry, one Fermat ladder per inverse.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call into this file.  Wire-major layout [rows][n] like include/p2e.h so outputs compare 1:1.
 *
 * Restates (file:line of Weobe/plonky2-ecdsa):
 *   convert_base                         src/gadgets/biguint.rs:27-51
 *   MulNonnativeGenerator::run_once      src/gates/mul_nonnative.rs:249-324
 *   CheckSumGenerator::run_once          src/gates/mul_nonnative.rs:513-531
 *   NonNativeInverseGenerator::run_once  src/gadgets/nonnative.rs:857-872
 *   MulNonnativeGate::eval_unfiltered    src/gates/mul_nonnative.rs:101-130
 *   CheckSumGate::eval_unfiltered        src/gates/mul_nonnative.rs:411-427
 * It is validated against oracle/spec.py (the Python big-int spec) by tests/test_oracle_c.py.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef uint32_t u32;
typedef uint64_t u64;
typedef unsigned __int128 u128;

#define MAXW 40
typedef struct { int n; u32 w[MAXW]; } bn;     /* n = significant words (0 for zero) */

static void bn_trim(bn* a) { while (a->n > 0 && a->w[a->n - 1] == 0) a->n--; }

static const u32 P_W[8] = {0xFFFFFC2Fu, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
static const u32 N_W[8] = {0xD0364141u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};

static void bn_modulus(int mod, bn* m) { m->n = 8; memcpy(m->w, mod == 0 ? P_W : N_W, 32); }

/* convert_base, biguint.rs:27-51 (vector in, vector out; returns length) */
static int convert_base(const u32* x, int len, int from, int to, u32* res) {
  int k = 0, offset = 0;
  u64 rem = 0, mask = (1ull << to) - 1;
  for (int i = 0; i < len; i++) {
    rem += (u64)x[i] << offset;
    offset += from;
    while (offset >= to) { res[k++] = (u32)(rem & mask); rem >>= to; offset -= to; }
  }
  if (rem != 0) res[k++] = (u32)rem;
  return k;
}

static void bn_mul(const bn* a, const bn* b, bn* r) {
  memset(r->w, 0, sizeof(u32) * (a->n + b->n + 1));
  for (int i = 0; i < a->n; i++) {
    u64 c = 0;
    for (int j = 0; j < b->n; j++) {
      u64 t = (u64)a->w[i] * b->w[j] + r->w[i + j] + c;
      r->w[i + j] = (u32)t;
      c = t >> 32;
    }
    r->w[i + b->n] = (u32)c;
  }
  r->n = a->n + b->n;
  bn_trim(r);
}

static int bn_cmp(const bn* a, const bn* b) {
  if (a->n != b->n) return a->n < b->n ? -1 : 1;
  for (int i = a->n - 1; i >= 0; i--) if (a->w[i] != b->w[i]) return a->w[i] < b->w[i] ? -1 : 1;
  return 0;
}

/* Knuth algorithm D, base 2^32.  b != 0. */
static void bn_divrem(const bn* a, const bn* b, bn* q, bn* r) {
  if (bn_cmp(a, b) < 0) { q->n = 0; *r = *a; return; }
  int n = b->n, m = a->n - b->n;
  if (n == 1) {
    u64 rem = 0;
    for (int i = a->n - 1; i >= 0; i--) { u64 cur = (rem << 32) | a->w[i]; q->w[i] = (u32)(cur / b->w[0]); rem = cur % b->w[0]; }
    q->n = a->n; bn_trim(q);
    r->n = 1; r->w[0] = (u32)rem; bn_trim(r);
    return;
  }
  int s = __builtin_clz(b->w[n - 1]);
  u32 v[MAXW], u[MAXW + 1];
  for (int i = n - 1; i > 0; i--) v[i] = s ? (b->w[i] << s) | (b->w[i - 1] >> (32 - s)) : b->w[i];
  v[0] = b->w[0] << s;
  u[a->n] = s ? a->w[a->n - 1] >> (32 - s) : 0;
  for (int i = a->n - 1; i > 0; i--) u[i] = s ? (a->w[i] << s) | (a->w[i - 1] >> (32 - s)) : a->w[i];
  u[0] = a->w[0] << s;
  for (int j = m; j >= 0; j--) {
    u64 num = ((u64)u[j + n] << 32) | u[j + n - 1];
    u64 qh = num / v[n - 1], rh = num % v[n - 1];
    while (qh >> 32 || qh * v[n - 2] > ((rh << 32) | u[j + n - 2])) { qh--; rh += v[n - 1]; if (rh >> 32) break; }
    int64_t br = 0; u64 cy = 0;
    for (int i = 0; i < n; i++) {
      u64 p = qh * v[i] + cy; cy = p >> 32;
      int64_t t = (int64_t)u[j + i] - (int64_t)(u32)p + br;
      u[j + i] = (u32)t; br = t >> 32;
    }
    int64_t t = (int64_t)u[j + n] - (int64_t)cy + br;
    u[j + n] = (u32)t;
    if (t < 0) {
      qh--;
      u64 c2 = 0;
      for (int i = 0; i < n; i++) { u64 t2 = (u64)u[j + i] + v[i] + c2; u[j + i] = (u32)t2; c2 = t2 >> 32; }
      u[j + n] += (u32)c2;
    }
    q->w[j] = (u32)qh;
  }
  q->n = m + 1; bn_trim(q);
  for (int i = 0; i < n; i++) r->w[i] = s ? (u[i] >> s) | ((u64)u[i + 1] << (32 - s)) : u[i];
  r->n = n; bn_trim(r);
}

/* ---- Goldilocks */
#define GLP 0xFFFFFFFF00000001ull
static u64 gl_red(u128 x) {
  u64 lo = (u64)x, hi = (u64)(x >> 64), hh = hi >> 32, hl = hi & 0xFFFFFFFFull;
  u64 t0 = lo - hh; if (lo < hh) t0 -= 0xFFFFFFFFull;
  u64 t1 = hl * 0xFFFFFFFFull, r = t0 + t1; if (r < t1) r += 0xFFFFFFFFull;
  if (r >= GLP) r -= GLP;
  return r;
}
static u64 gl_mul(u64 a, u64 b) { return gl_red((u128)a * b); }
static u64 gl_add(u64 a, u64 b) { u64 s = a + b; if (s < a) s += 0xFFFFFFFFull; if (s >= GLP) s -= GLP; return s; }
static u64 gl_sub(u64 a, u64 b) { u64 d = a - b; if (a < b) d -= 0xFFFFFFFFull; return d; }
static u64 gl_pow(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = gl_mul(r, a); a = gl_mul(a, a); e >>= 1; } return r; }

/* ---- A8 + A9 for one instance.  Returns an error bitmask compatible with P2E_ERR_*. */
static u32 mul_nonnative_one(int mod, const u64* xw, const u64* yw, u64* out /*51: r q cs b*/) {
  for (int i = 0; i < 9; i++) if ((xw[i] >> 32) || (yw[i] >> 32)) { memset(out, 0, 51 * 8); return 1; }
  for (int i = 0; i < 9; i++) if ((xw[i] >> 29) || (yw[i] >> 29)) { memset(out, 0, 51 * 8); return 2; }
  bn m; bn_modulus(mod, &m);
  u32 m29[12] = {0}, x29[9], y29[9], q29[24] = {0}, r29[12] = {0};
  convert_base(m.w, m.n, 32, 29, m29);
  for (int i = 0; i < 9; i++) { x29[i] = (u32)xw[i]; y29[i] = (u32)yw[i]; }
  bn x, y, prod, q, r;
  x.n = convert_base(x29, 9, 29, 32, x.w); bn_trim(&x);
  y.n = convert_base(y29, 9, 29, 32, y.w); bn_trim(&y);
  bn_mul(&x, &y, &prod);
  bn_divrem(&prod, &m, &q, &r);
  convert_base(q.w, q.n, 32, 29, q29);
  convert_base(r.w, r.n, 32, 29, r29);
  u64 cs[17];
  for (int i = 0; i < 17; i++) {
    int lo = i - 8 > 0 ? i - 8 : 0, hi = i + 1 < 9 ? i + 1 : 9;
    u64 acc = 0;
    for (int j = lo; j < hi; j++)
      acc = gl_add(acc, gl_sub(gl_mul(q29[i - j], m29[j]), gl_mul(x29[j], y29[i - j])));
    if (i < 9) acc = gl_add(acc, r29[i]);
    cs[i] = acc;
  }
  for (int i = 0; i < 9; i++) { out[i] = r29[i]; out[9 + i] = q29[i]; }
  for (int i = 0; i < 17; i++) out[18 + i] = cs[i];
  /* CheckSumGenerator: field division by 2^29 per carry (plonky2 Field::div = mul by inverse) */
  u64 last = 0, inv_base = gl_pow(1ull << 29, GLP - 2);
  u32 err = 0;
  for (int i = 0; i < 16; i++) {
    u64 b = gl_mul(gl_add(cs[i], last), inv_base);
    u64 o = gl_add(b, 1ull << 33);
    out[35 + i] = o; last = b;
    if (o >= (1ull << 34)) err |= 16;
  }
  return err;
}

void orc_mul_nonnative(int mod, size_t n, size_t i0, size_t i1, const u64* x, const u64* y, u64* out, u32* err) {
  for (size_t i = i0; i < i1; i++) {
    u64 xw[9], yw[9], o[51];
    for (int l = 0; l < 9; l++) { xw[l] = x[l * n + i]; yw[l] = y[l * n + i]; }
    err[i] = mul_nonnative_one(mod, xw, yw, o);
    for (int l = 0; l < 51; l++) out[l * n + i] = o[l];
  }
}

/* ---- A7: Fermat ladder (square-and-multiply over order-2), BigUint mul + divrem per step */
static void bn_mulmod(const bn* a, const bn* b, const bn* m, bn* r) { bn p, q; bn_mul(a, b, &p); bn_divrem(&p, m, &q, r); }

static u32 inv_nonnative_one(int mod, const u64* xw, u64* out /*18: div inv*/) {
  bn m; bn_modulus(mod, &m);
  /* get_biguint_target (Horner) for in-range limbs, then from_noncanonical / canonical */
  u32 x29[9];
  for (int i = 0; i < 9; i++) { if (xw[i] >> 29) { memset(out, 0, 18 * 8); return 4; } x29[i] = (u32)xw[i]; }
  bn x; x.n = convert_base(x29, 9, 29, 32, x.w); bn_trim(&x);
  memset(out, 0, 18 * 8);
  if (x.n > 8) return 4;
  if (bn_cmp(&x, &m) >= 0) { /* one conditional subtract */
    int64_t br = 0;
    for (int i = 0; i < 8; i++) { int64_t t = (int64_t)x.w[i] - m.w[i] + br; x.w[i] = (u32)t; br = t >> 32; }
    bn_trim(&x);
  }
  if (x.n == 0) return 8;
  bn e = m; e.w[0] -= 2;                    /* order - 2 (no borrow for either modulus) */
  bn acc; acc.n = 1; acc.w[0] = 1;
  bn base = x, t;
  for (int bit = 0; bit < 256; bit++) {     /* exp_biguint: LSB-first square and multiply */
    if ((e.w[bit >> 5] >> (bit & 31)) & 1) { bn_mulmod(&acc, &base, &m, &t); acc = t; }
    bn_mulmod(&base, &base, &m, &t); base = t;
  }
  bn prod, div, rem;
  bn_mul(&x, &acc, &prod);
  bn_divrem(&prod, &m, &div, &rem);
  u32 d29[12] = {0}, i29[12] = {0};
  convert_base(div.w, div.n, 32, 29, d29);
  convert_base(acc.w, acc.n, 32, 29, i29);
  for (int i = 0; i < 9; i++) { out[i] = d29[i]; out[9 + i] = i29[i]; }
  return 0;
}

void orc_inv_nonnative(int mod, size_t n, size_t i0, size_t i1, const u64* x, u64* out, u32* err) {
  for (size_t i = i0; i < i1; i++) {
    u64 xw[9], o[18];
    for (int l = 0; l < 9; l++) xw[l] = x[l * n + i];
    err[i] = inv_nonnative_one(mod, xw, o);
    for (int l = 0; l < 18; l++) out[l * n + i] = o[l];
  }
}

/* ---- A11 / A12: eval_unfiltered per point, recomputing the modulus limbs per call like the
 * reference does (mul_nonnative.rs:103-109) */
void orc_eval_mul_nonnative(int mod, size_t n, size_t i0, size_t i1, const u64* w, u64* out) {
  for (size_t i = i0; i < i1; i++) {
    bn m; bn_modulus(mod, &m);
    u32 m29[12] = {0};
    convert_base(m.w, m.n, 32, 29, m29);
    for (int c = 0; c < 17; c++) {
      int lo = c - 8 > 0 ? c - 8 : 0, hi = c + 1 < 9 ? c + 1 : 9;
      u64 acc = 0;
      for (int j = lo; j < hi; j++) {
        u64 qm = gl_mul(m29[j], w[(27 + c - j) * n + i]);
        u64 xy = gl_mul(w[j * n + i], w[(9 + c - j) * n + i]);
        acc = gl_sub(gl_add(acc, qm), xy);
      }
      if (c < 9) acc = gl_add(acc, w[(18 + c) * n + i]);
      out[c * n + i] = gl_sub(acc, w[(36 + c) * n + i]);
    }
  }
}

void orc_eval_check_sum(size_t n, size_t i0, size_t i1, const u64* w, u64* out) {
  for (size_t i = i0; i < i1; i++) {
    u64 last = 0;
    for (int c = 0; c < 17; c++) {
      if (c < 16) {
        u64 o = gl_sub(w[(17 + c) * n + i], 1ull << 33);
        out[c * n + i] = gl_sub(gl_add(w[c * n + i], last), gl_mul(1ull << 29, o));
        last = o;
      } else {
        out[c * n + i] = gl_add(w[c * n + i], last);
      }
    }
  }
}
