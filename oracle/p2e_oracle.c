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

/* "best-CPU" variant of the inverse: extended Euclid on the same BigUint primitives instead of
 * the reference's Fermat ladder.  The inverse is unique, so every witness bit is unchanged; only the CPU time is
 * (about 10x fewer word operations per inverse).  orc_set_inverse_mode(1) selects it for the timed baseline. */
static int g_inverse_mode = 0;
void orc_set_inverse_mode(int mode) { g_inverse_mode = mode; }
static void bn_add(const bn* a, const bn* b, bn* r);
static void bn_sub(const bn* a, const bn* b, bn* r);
static void bn_inv_euclid(const bn* x, const bn* m, bn* inv) {
  /* r0 = m, r1 = x; cofactors alternate in sign: |t_{i+1}| = |t_{i-1}| + q |t_i| */
  bn r0 = *m, r1 = *x, t0, t1, q, r, p, s;
  t0.n = 0;
  t1.n = 1; t1.w[0] = 1;
  int neg = 0;                                /* sign of t1 */
  while (r1.n != 0) {
    bn_divrem(&r0, &r1, &q, &r);
    bn_mul(&q, &t1, &p);
    bn_add(&t0, &p, &s);
    t0 = t1; t1 = s;
    r0 = r1; r1 = r;
    neg = !neg;
  }
  /* gcd = r0 = 1; x * (+-t0) = 1 (mod m), sign of t0 is !neg */
  if (!neg) { bn_sub(m, &t0, inv); } else { *inv = t0; }
  bn_trim(inv);
}

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
  if (g_inverse_mode == 1) {
    bn_inv_euclid(&x, &m, &acc);
  } else {
    for (int bit = 0; bit < 256; bit++) {     /* exp_biguint: LSB-first square and multiply */
      if ((e.w[bit >> 5] >> (bit & 31)) & 1) { bn_mulmod(&acc, &base, &m, &t); acc = t; }
      bn_mulmod(&base, &base, &m, &t); base = t;
    }
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

/* =========================================================================================
 * A13: the whole verify_secp256k1_message_circuit tape on the CPU, reference style: every
 * generator runs its own BigUint arithmetic and every curve_add / curve_double pays one Fermat
 * ladder (gadgets/curve.rs:160-223 -> gadgets/nonnative.rs:857-872).  Circuit constants
 * (blinding points, fixed-base window tables, GLV constants) are passed in by the caller
 * (computed by oracle/spec.py), exactly as the reference computes them once at circuit-build time.
 * Restates: gadgets/nonnative.rs:245-388,491-536,584-596,626-872; gadgets/curve.rs:123-243;
 * gadgets/curve_fixed_base.rs:18-66; gadgets/curve_msm.rs:21-79; gadgets/glv.rs:53-142;
 * curve/glv.rs:39-77; gadgets/split_nonnative.rs:25-72; gadgets/ecdsa.rs:30-53.
 */
typedef struct { int nl; u32 l[9]; } nn;            /* NonNativeTarget wires */
typedef struct { nn x, y; } apt;                     /* AffinePointTarget */
typedef struct { u64* out; size_t n, i, slot; u32 err; } tapew;

static void emit(tapew* t, u64 v) { t->out[t->slot * t->n + t->i] = v; t->slot++; }

static void bn_from_nn(const nn* a, bn* r) { r->n = convert_base(a->l, a->nl, 29, 32, r->w); bn_trim(r); }
static void nn_from_bn(const bn* a, int nl, nn* r) {
  u32 t[24] = {0};
  convert_base(a->w, a->n, 32, 29, t);
  r->nl = nl;
  for (int i = 0; i < nl; i++) r->l[i] = t[i];
}
static void bn_add(const bn* a, const bn* b, bn* r) {
  int n = a->n > b->n ? a->n : b->n; u64 c = 0;
  for (int i = 0; i < n; i++) { u64 t = c + (i < a->n ? a->w[i] : 0) + (i < b->n ? b->w[i] : 0); r->w[i] = (u32)t; c = t >> 32; }
  r->w[n] = (u32)c; r->n = n + 1; bn_trim(r);
}
static void bn_sub(const bn* a, const bn* b, bn* r) {   /* a >= b */
  int64_t br = 0;
  for (int i = 0; i < a->n; i++) { int64_t t = (int64_t)a->w[i] - (i < b->n ? b->w[i] : 0) + br; r->w[i] = (u32)t; br = t >> 32; }
  r->n = a->n; bn_trim(r);
}
static void canon(bn* a, const bn* m) { if (bn_cmp(a, m) >= 0) { bn t; bn_sub(a, m, &t); *a = t; } }

static nn g_add(tapew* t, int mod, const nn* a, const nn* b) {      /* nonnative.rs:626-645 */
  bn m, x, y, s; bn_modulus(mod, &m);
  bn_from_nn(a, &x); bn_from_nn(b, &y); canon(&x, &m); canon(&y, &m);
  bn_add(&x, &y, &s);
  u32 ov = 0;
  if (bn_cmp(&s, &m) > 0) { bn d; bn_sub(&s, &m, &d); s = d; ov = 1; }
  nn r; nn_from_bn(&s, 9, &r);
  for (int i = 0; i < 9; i++) emit(t, r.l[i]);
  emit(t, ov);
  return r;
}
static nn g_sub(tapew* t, int mod, const nn* a, const nn* b) {      /* nonnative.rs:792-810 */
  bn m, x, y, d; bn_modulus(mod, &m);
  bn_from_nn(a, &x); bn_from_nn(b, &y); canon(&x, &m); canon(&y, &m);
  u32 ov = 0;
  if (bn_cmp(&x, &y) >= 0) bn_sub(&x, &y, &d);
  else { bn s; bn_add(&m, &x, &s); bn_sub(&s, &y, &d); ov = 1; }
  nn r; nn_from_bn(&d, 9, &r);
  for (int i = 0; i < 9; i++) emit(t, r.l[i]);
  emit(t, ov);
  return r;
}
static nn g_add_many4(tapew* t, int mod, const nn* a, const nn* b, const nn* c, const nn* d) {  /* nonnative.rs:696-728 */
  bn m, acc, x, s, q, r; bn_modulus(mod, &m);
  acc.n = 0;
  const nn* v[4] = {a, b, c, d};
  for (int k = 0; k < 4; k++) { bn_from_nn(v[k], &x); canon(&x, &m); bn_add(&acc, &x, &s); acc = s; }
  bn_divrem(&acc, &m, &q, &r);
  nn o; nn_from_bn(&r, 9, &o);
  for (int i = 0; i < 9; i++) emit(t, o.l[i]);
  emit(t, q.n ? q.w[0] : 0);
  return o;
}
static nn g_mul(tapew* t, int mod, const nn* a, const nn* b) {      /* nonnative.rs:390-464 */
  u64 xw[9] = {0}, yw[9] = {0}, o[51];
  for (int i = 0; i < a->nl; i++) xw[i] = a->l[i];
  for (int i = 0; i < b->nl; i++) yw[i] = b->l[i];
  t->err |= mul_nonnative_one(mod, xw, yw, o);
  nn r; r.nl = 9;
  for (int i = 0; i < 9; i++) r.l[i] = (u32)o[i];
  for (int i = 0; i < 51; i++) emit(t, o[i]);
  return r;
}
static nn g_inv(tapew* t, int mod, const nn* a) {                  /* nonnative.rs:857-872 */
  u64 xw[9] = {0}, o[18];
  for (int i = 0; i < a->nl; i++) xw[i] = a->l[i];
  t->err |= inv_nonnative_one(mod, xw, o);
  nn r; r.nl = 9;
  for (int i = 0; i < 9; i++) r.l[i] = (u32)o[9 + i];
  for (int i = 0; i < 18; i++) emit(t, o[i]);
  return r;
}
static nn nn_bool(const nn* a, u32 b) { nn r = *a; for (int i = 0; i < r.nl; i++) r.l[i] *= b; return r; }
static nn g_cond_neg(tapew* t, int mod, const nn* x, u32 b) {       /* nonnative.rs:584-596 */
  nn zero; zero.nl = 0;
  nn neg = g_sub(t, mod, &zero, x);
  nn xt = nn_bool(&neg, b), xf = nn_bool(x, 1 - b);
  return g_add(t, mod, &xt, &xf);
}
static apt c_add(tapew* t, const apt* p1, const apt* p2) {          /* curve.rs:202-223 */
  nn u = g_sub(t, 0, &p2->y, &p1->y);
  nn v = g_sub(t, 0, &p2->x, &p1->x);
  nn vi = g_inv(t, 0, &v);
  nn s = g_mul(t, 0, &u, &vi);
  nn s2 = g_mul(t, 0, &s, &s);
  nn xs = g_add(t, 0, &p2->x, &p1->x);
  apt r;
  r.x = g_sub(t, 0, &s2, &xs);
  nn xd = g_sub(t, 0, &p1->x, &r.x);
  nn pr = g_mul(t, 0, &s, &xd);
  r.y = g_sub(t, 0, &pr, &p1->y);
  return r;
}
static apt c_double(tapew* t, const apt* p) {                       /* curve.rs:160-185 */
  nn zero; zero.nl = 0;
  nn dy = g_add(t, 0, &p->y, &p->y);
  nn idy = g_inv(t, 0, &dy);
  nn xx = g_mul(t, 0, &p->x, &p->x);
  nn tr = g_add_many4(t, 0, &xx, &xx, &xx, &zero);
  nn lam = g_mul(t, 0, &tr, &idy);
  nn lam2 = g_mul(t, 0, &lam, &lam);
  nn xd = g_add(t, 0, &p->x, &p->x);
  apt r;
  r.x = g_sub(t, 0, &lam2, &xd);
  nn xdiff = g_sub(t, 0, &p->x, &r.x);
  nn lx = g_mul(t, 0, &lam, &xdiff);
  r.y = g_sub(t, 0, &lx, &p->y);
  return r;
}
static apt c_cond_add(tapew* t, const apt* p1, const apt* p2, u32 b) {   /* curve.rs:225-243 */
  apt s = c_add(t, p1, p2);
  nn xt = nn_bool(&s.x, b), yt = nn_bool(&s.y, b), xf = nn_bool(&p1->x, 1 - b), yf = nn_bool(&p1->y, 1 - b);
  apt r;
  r.x = g_add(t, 0, &xt, &xf);
  r.y = g_add(t, 0, &yt, &yf);
  return r;
}
static u32 digit_of(const nn* a, int pos, int bits) {               /* split_nonnative.rs:25-72 */
  u32 v = 0;
  for (int k = 0; k < bits; k++) { int bit = pos + k, l = bit / 29; if (l < a->nl) v |= ((a->l[l] >> (bit % 29)) & 1u) << k; }
  return v;
}
static nn nn_const(const u32* limbs9) { nn r; r.nl = 9; for (int i = 0; i < 9; i++) r.l[i] = limbs9[i]; while (r.nl > 0 && r.l[r.nl - 1] == 0) r.nl--; return r; }
static apt apt_const(const u32* p18) { apt r; r.x = nn_const(p18); r.y = nn_const(p18 + 9); return r; }

/* consts layout (u32 limbs): [0] rando(18) [18] neg_rando(18) [36] neg_spm(18) [54] beta(9) [63] S(9)
 * [72] 2*B2(9) [81] 2*MINUS_B1(9) [90] A1(9) [99] MINUS_B1(9) [108] A2(9); table: 66*16 points * 18 */
static void glv_gen(tapew* t, const u32* C, const nn* k, nn* k1, nn* k2, u32* n1, u32* n2) {   /* curve/glv.rs:39-77 */
  bn n, kk, x, num, q, r, two_n, c1, c2, t1, t2, a, b;
  bn_modulus(1, &n);
  bn_from_nn(k, &kk); canon(&kk, &n);
  bn_add(&n, &n, &two_n);
  nn cst;
  /* c = round(B*k/n) = (2*B*k + n) / (2n) */
  cst = nn_const(C + 72); bn_from_nn(&cst, &x); bn_mul(&x, &kk, &num); bn_add(&num, &n, &a); bn_divrem(&a, &two_n, &c1, &r);
  cst = nn_const(C + 81); bn_from_nn(&cst, &x); bn_mul(&x, &kk, &num); bn_add(&num, &n, &a); bn_divrem(&a, &two_n, &c2, &r);
  bn A1, MB1, A2;
  cst = nn_const(C + 90); bn_from_nn(&cst, &A1);
  cst = nn_const(C + 99); bn_from_nn(&cst, &MB1);
  cst = nn_const(C + 108); bn_from_nn(&cst, &A2);
  /* k1 = k - c1*A1 - c2*A2 ; k2 = c1*MINUS_B1 - c2*B2  (mod n), B2 == A1 */
  bn_mulmod(&c1, &A1, &n, &t1); bn_mulmod(&c2, &A2, &n, &t2);
  if (bn_cmp(&kk, &t1) >= 0) bn_sub(&kk, &t1, &a); else { bn_add(&kk, &n, &b); bn_sub(&b, &t1, &a); }
  bn k1r, k2r;
  if (bn_cmp(&a, &t2) >= 0) bn_sub(&a, &t2, &k1r); else { bn_add(&a, &n, &b); bn_sub(&b, &t2, &k1r); }
  bn_mulmod(&c1, &MB1, &n, &t1); bn_mulmod(&c2, &A1, &n, &t2);
  if (bn_cmp(&t1, &t2) >= 0) bn_sub(&t1, &t2, &k2r); else { bn_add(&t1, &n, &b); bn_sub(&b, &t2, &k2r); }
  bn half = n; for (int i = 0; i < 8; i++) half.w[i] = (n.w[i] >> 1) | (i < 7 ? n.w[i + 1] << 31 : 0);
  bn_trim(&half);
  *n1 = bn_cmp(&k1r, &half) > 0; if (*n1) { bn_sub(&n, &k1r, &a); k1r = a; }
  *n2 = bn_cmp(&k2r, &half) > 0; if (*n2) { bn_sub(&n, &k2r, &a); k2r = a; }
  nn_from_bn(&k1r, 5, k1); nn_from_bn(&k2r, 5, k2);
  for (int i = 0; i < 5; i++) emit(t, k1->l[i]);
  for (int i = 0; i < 5; i++) emit(t, k2->l[i]);
  emit(t, *n1); emit(t, *n2);
}

static int nn_eq(const nn* a, const nn* b) {
  for (int i = 0; i < 9; i++) if ((i < a->nl ? a->l[i] : 0) != (i < b->nl ? b->l[i] : 0)) return 0;
  return 1;
}

static void ecdsa_one(tapew* t, const u32* C, const u32* table, const nn* msg, const nn* r, const nn* s, const apt* pk) {
  /* curve_assert_valid, curve.rs:123-135 */
  nn zero; zero.nl = 0;
  nn seven; seven.nl = 1; seven.l[0] = 7;
  nn y2 = g_mul(t, 0, &pk->y, &pk->y);
  nn x2 = g_mul(t, 0, &pk->x, &pk->x);
  nn x3 = g_mul(t, 0, &x2, &pk->x);
  nn ax = g_mul(t, 0, &zero, &pk->x);
  nn axb = g_add(t, 0, &ax, &seven);
  nn rhs = g_add(t, 0, &x3, &axb);
  if (!nn_eq(&y2, &rhs)) t->err |= 128;
  /* ecdsa.rs:40-42 */
  nn c = g_inv(t, 1, s);
  nn u1 = g_mul(t, 1, msg, &c);
  nn u2 = g_mul(t, 1, r, &c);
  /* fixed_base_curve_mul_circuit, curve_fixed_base.rs:18-66 */
  apt rando = apt_const(C), neg_rando = apt_const(C + 18), neg_spm = apt_const(C + 36);
  apt res = rando;
  for (int w = 0; w < 66; w++) {
    u32 d = digit_of(&u1, 4 * w, 4);
    apt tp = apt_const(table + ((size_t)w * 16 + d) * 18);
    res = c_cond_add(t, &res, &tp, d != 0);
  }
  apt point1 = c_add(t, &res, &neg_rando);
  /* glv_mul, glv.rs:87-104 */
  nn k1, k2; u32 n1, n2;
  glv_gen(t, C, &u2, &k1, &k2, &n1, &n2);
  nn k1_raw = g_cond_neg(t, 1, &k1, n1);
  nn k2_raw = g_cond_neg(t, 1, &k2, n2);
  nn S = nn_const(C + 63);
  nn sbk = g_mul(t, 1, &S, &k2_raw);
  sbk = g_add(t, 1, &sbk, &k1_raw);
  if (!nn_eq(&sbk, &u2)) t->err |= 128;
  nn beta = nn_const(C + 54);
  apt sp; sp.x = g_mul(t, 0, &beta, &pk->x);
  apt p_neg; p_neg.x = pk->x; p_neg.y = g_cond_neg(t, 0, &pk->y, n1);
  sp.y = g_cond_neg(t, 0, &pk->y, n2);
  /* curve_msm_circuit, curve_msm.rs:21-79 */
  apt pre[16];
  for (int i = 0; i < 16; i++) pre[i] = p_neg;
  apt cur_p = rando, cur_q = rando;
  for (int i = 0; i < 4; i++) {
    pre[i] = cur_p; pre[4 * i] = cur_q;
    cur_p = c_add(t, &cur_p, &p_neg);
    cur_q = c_add(t, &cur_q, &sp);
  }
  for (int i = 1; i < 4; i++) { pre[i] = c_add(t, &pre[i], &neg_rando); pre[4 * i] = c_add(t, &pre[4 * i], &neg_rando); }
  for (int i = 1; i < 4; i++) for (int j = 1; j < 4; j++) pre[i + 4 * j] = c_add(t, &pre[i], &pre[4 * j]);
  res = rando;
  for (int dgt = 72; dgt >= 0; dgt--) {
    res = c_double(t, &res);
    res = c_double(t, &res);
    u32 idx = 4 * digit_of(&k2, 2 * dgt, 2) + digit_of(&k1, 2 * dgt, 2);
    res = c_cond_add(t, &res, &pre[idx], idx != 0);
  }
  apt point2 = c_add(t, &res, &neg_spm);
  apt point = c_add(t, &point1, &point2);
  if (!nn_eq(&point.x, r)) t->err |= 128;
}

/* msg, r, s, pkx, pky: [9][n]; out: [slots][n]; processes instances [i0, i1) */
void orc_ecdsa_witness(size_t n, size_t i0, size_t i1, const u64* msg, const u64* r, const u64* s, const u64* pkx,
                       const u64* pky, const u32* consts, const u32* table, u64* out, u32* err) {
  for (size_t i = i0; i < i1; i++) {
    nn m, rr, ss; apt pk;
    m.nl = rr.nl = ss.nl = pk.x.nl = pk.y.nl = 9;
    for (int l = 0; l < 9; l++) {
      m.l[l] = (u32)msg[l * n + i]; rr.l[l] = (u32)r[l * n + i]; ss.l[l] = (u32)s[l * n + i];
      pk.x.l[l] = (u32)pkx[l * n + i]; pk.y.l[l] = (u32)pky[l * n + i];
    }
    /* constant_nonnative drops leading zero limbs */
    nn* all[5] = {&m, &rr, &ss, &pk.x, &pk.y};
    for (int k = 0; k < 5; k++) while (all[k]->nl > 0 && all[k]->l[all[k]->nl - 1] == 0) all[k]->nl--;
    tapew t; t.out = out; t.n = n; t.i = i; t.slot = 0; t.err = 0;
    ecdsa_one(&t, consts, table, &m, &rr, &ss, &pk);
    err[i] = t.err;
  }
}
