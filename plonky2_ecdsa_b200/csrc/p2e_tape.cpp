// p2e_tape.cpp -- host-side mirror of the reference's gadget call graph; see p2e_tape.h.
// Gadget methods below keep the reference's names and call order; each cites the reference
// function it mirrors.  Circuit CONSTANTS (blinding point, fixed-base tables) are computed here
// with native affine arithmetic, as the reference does at circuit-build time
// (gadgets/curve_fixed_base.rs:24-52, gadgets/curve_msm.rs:33-39,74-76).
#include "p2e_tape.h"

#include <string.h>

#include <mutex>
#include <string>

#include "../../include/p2e.h"
#include "p2e_common.h"
#include "p2e_field.cuh"

namespace p2e {
namespace {

// ------------------------------------------------------------------ native field / curve (constants)
struct Fe {
  u32 l[NL];
};
static Fe fe_zero() { Fe r; memset(r.l, 0, sizeof r.l); return r; }
static Fe fe_small(u32 v) { Fe r = fe_zero(); r.l[0] = v; return r; }
static Fe fe_from_u64x4(const u64 (&w)[4]) {     // 256-bit little-endian words -> 29-bit limbs
  Fe r;
  for (int i = 0; i < NL; i++) {
    int bit = LB * i, wi = bit >> 6, sh = bit & 63;
    u64 v = w[wi] >> sh;
    if (sh + LB > 64 && wi + 1 < 4) v |= w[wi + 1] << (64 - sh);
    r.l[i] = (u32)v & LMASK;
  }
  return r;
}
static Fe fe_mul(const Fe& a, const Fe& b) { Fe r; mulmod<ModP>(a.l, b.l, r.l); return r; }
static Fe fe_add(const Fe& a, const Fe& b) { Fe r; addmod<ModP>(a.l, b.l, r.l); return r; }
static Fe fe_sub(const Fe& a, const Fe& b) { Fe r; submod<ModP>(a.l, b.l, r.l); return r; }
static Fe fe_inv(const Fe& a) { Fe r; inv_p(a.l, r.l); return r; }
static bool fe_eq(const Fe& a, const Fe& b) { return eq9(a.l, b.l); }
static bool fe_is_zero(const Fe& a) { return is_zero9(a.l); }

struct APoint {
  Fe x, y;
  bool inf;
};
static APoint ap_inf() { APoint p; p.x = fe_zero(); p.y = fe_zero(); p.inf = true; return p; }
static APoint ap_neg(const APoint& p) { APoint r = p; if (!p.inf) r.y = fe_sub(fe_zero(), p.y); return r; }
static APoint ap_double(const APoint& p) {                    // curve/curve_types.rs:83-102
  if (p.inf || fe_is_zero(p.y)) return ap_inf();
  Fe xx = fe_mul(p.x, p.x);
  Fe lam = fe_mul(fe_add(fe_add(xx, xx), xx), fe_inv(fe_add(p.y, p.y)));
  APoint r;
  r.inf = false;
  r.x = fe_sub(fe_mul(lam, lam), fe_add(p.x, p.x));
  r.y = fe_sub(fe_mul(lam, fe_sub(p.x, r.x)), p.y);
  return r;
}
static APoint ap_add(const APoint& p, const APoint& q) {      // curve/curve_adds.rs (affine + affine)
  if (p.inf) return q;
  if (q.inf) return p;
  if (fe_eq(p.x, q.x)) return fe_eq(p.y, q.y) ? ap_double(p) : ap_inf();
  Fe lam = fe_mul(fe_sub(q.y, p.y), fe_inv(fe_sub(q.x, p.x)));
  APoint r;
  r.inf = false;
  r.x = fe_sub(fe_sub(fe_mul(lam, lam), p.x), q.x);
  r.y = fe_sub(fe_mul(lam, fe_sub(p.x, r.x)), p.y);
  return r;
}
static APoint ap_mul(const u64 (&k)[4], const APoint& p) {    // double-and-add, LSB first
  APoint acc = ap_inf(), base = p;
  for (int bit = 0; bit < 256; bit++) {
    if ((k[bit >> 6] >> (bit & 63)) & 1) acc = ap_add(acc, base);
    base = ap_double(base);
  }
  return acc;
}

static const u64 GEN_X[4] = {0x59F2815B16F81798ull, 0x029BFCDB2DCE28D9ull, 0x55A06295CE870B07ull, 0x79BE667EF9DCBBACull};
static const u64 GEN_Y[4] = {0x9C47D08FFB10D4B8ull, 0xFD17B448A6855419ull, 0x5DA4FBFC0E1108A8ull, 0x483ADA7726A3C465ull};
static const u64 SCALAR_N[4] = {0xBFD25E8CD0364141ull, 0xBAAEDCE6AF48A03Bull, 0xFFFFFFFFFFFFFFFEull, 0xFFFFFFFFFFFFFFFFull};
// curve/glv.rs:11-23
static const u64 GLV_BETA_W[4] = {13923278643952681454ull, 11308619431505398165ull, 7954561588662645993ull, 8856726876819556112ull};
static const u64 GLV_S_W[4] = {16069571880186789234ull, 1310022930574435960ull, 11900229862571533402ull, 6008836872998760672ull};

static APoint generator() {
  APoint g;
  u64 gx[4], gy[4];
  memcpy(gx, GEN_X, 32);
  memcpy(gy, GEN_Y, 32);
  g.x = fe_from_u64x4(gx);
  g.y = fe_from_u64x4(gy);
  g.inf = false;
  return g;
}

// ------------------------------------------------------------------ Keccak-256 (original 0x01 padding)
static u64 rol64(u64 v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; }
static void keccak_f(u64 (&st)[25]) {
  static const u64 RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull,
      0x000000000000808Bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
      0x000000000000008Aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000Aull,
      0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull, 0x8000000000008003ull,
      0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  for (int rnd = 0; rnd < 24; rnd++) {
    u64 c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) st[i] ^= d[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(st[x + 5 * y], ROT[x + 5 * y]);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) st[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    st[0] ^= RC[rnd];
  }
}
// KeccakHash::<32>::hash_no_pad(&[F::ZERO]) -> BigUint::from_bytes_le -> from_noncanonical_biguint
// (gadgets/curve_msm.rs:33-36): Keccak-256 of the 8 little-endian bytes of the field element 0.
static void blinding_scalar(u64 (&k)[4]) {
  u64 st[25];
  memset(st, 0, sizeof st);
  // message: 8 zero bytes, pad10*1 with 0x01 domain byte, rate 136 bytes
  st[1] ^= 0x01ull;                 // byte 8
  st[16] ^= 0x8000000000000000ull;  // byte 135
  keccak_f(st);
  for (int i = 0; i < 4; i++) k[i] = st[i];
  // to_canonical: one conditional subtract of n
  bool ge = true;
  for (int i = 3; i >= 0; i--) {
    if (k[i] != SCALAR_N[i]) { ge = k[i] > SCALAR_N[i]; break; }
  }
  if (ge) {
    unsigned __int128 br = 0;
    for (int i = 0; i < 4; i++) {
      unsigned __int128 t = (unsigned __int128)k[i] - SCALAR_N[i] - br;
      k[i] = (u64)t;
      br = (t >> 64) & 1;
    }
  }
}

// ------------------------------------------------------------------ builder
constexpr int32_t PID_INPUT_BASE = 1 << 19;

struct NNT {      // NonNativeTarget (gadgets/nonnative.rs:34-38)
  Ref r;
};
struct BoolT {    // BoolTarget living in an output slot
  uint8_t space;
  uint32_t off;
  bool inv;
};
struct PtT {      // AffinePointTarget (gadgets/curve.rs:15-19) + the chain engine's point id
  int32_t pid;
  NNT x, y;
};

static Ref mkref(uint8_t space, uint32_t off, uint8_t nl) {
  Ref r;
  r.space = space;
  r.nl = nl;
  r.bspace = SP_NONE;
  r.binv = 0;
  r.off = off;
  r.boff = 0;
  return r;
}

class Builder {
 public:
  Tape t;
  APoint rando, neg_rando;
  int cur_chain = 0;
  int cur_round = 0;                 // round given to the prep / point ops emitted next
  // dynamic (C-API) use: value / point handles
  std::vector<NNT> vals;
  std::vector<uint8_t> val_round;    // round from which the value is available
  std::vector<uint8_t> pt_round;     // round of the op producing the point (or of its coordinate values)
  std::vector<PtT> pts;
  std::vector<int8_t> pt_chain;      // chain of the producing op, -1 for input / constant points
  int n_fixed = 0, n_msm = 0;
  std::string error;
  // coordinates of every point as targets (for the factor references of MulRec)
  std::vector<PtT> op_pts;           // result point of curve op k
  PtT cur_pt[N_CHAINS];              // the running point of a chain (PF_*_CUR)

  Builder() {
    u64 k[4];
    blinding_scalar(k);
    rando = ap_mul(k, generator());
    neg_rando = ap_neg(rando);
    for (int i = 0; i < 16; i++) t.pre_pid[i] = -1;
    t.fixed_scalar = mkref(SP_NONE, 0, 0);
    t.msm_k1 = mkref(SP_NONE, 0, 0);
    t.msm_k2 = mkref(SP_NONE, 0, 0);
  }

  uint32_t emit(GenKind k, uint32_t width) {
    uint32_t off = t.n_slots;
    t.gen_kind.push_back(k);
    t.gen_off.push_back(off);
    t.n_slots += width;
    return off;
  }

  NNT input(uint32_t k) {
    if (t.n_inputs < k + 1) t.n_inputs = k + 1;
    NNT v;
    v.r = mkref(SP_IN, 9 * k, 9);
    return v;
  }

  // constant_nonnative (nonnative.rs:205-208 / biguint.rs:165-175): convert_base drops leading
  // zero limbs, zero has NO limbs
  NNT constant_nonnative(const Fe& v) {
    int nl = NL;
    while (nl > 0 && v.l[nl - 1] == 0) nl--;
    NNT r;
    r.r = mkref(SP_CONST, (uint32_t)t.const_limbs.size(), (uint8_t)nl);
    for (int i = 0; i < nl; i++) t.const_limbs.push_back(v.l[i]);
    return r;
  }

  void prep2(PrepOpcode op, int mod, uint32_t slot, const NNT& a, const NNT& b) {
    PrepOp p;
    p.opcode = op;
    p.mod = (uint8_t)mod;
    p.pad = 0;
    p.round = (uint8_t)cur_round;
    p.slot = slot;
    p.a = a.r;
    p.b = b.r;
    t.prep.push_back(p);
    if ((uint32_t)cur_round + 1 > t.n_rounds) t.n_rounds = cur_round + 1;
  }

  NNT add_nonnative(const NNT& a, const NNT& b, int mod) {          // nonnative.rs:245-276
    uint32_t s = emit(G_ADD, 10);
    prep2(OP_ADD, mod, s, a, b);
    NNT r; r.r = mkref(SP_OUT, s, 9); return r;
  }
  NNT sub_nonnative(const NNT& a, const NNT& b, int mod) {          // nonnative.rs:356-388
    uint32_t s = emit(G_SUB, 10);
    prep2(OP_SUB, mod, s, a, b);
    NNT r; r.r = mkref(SP_OUT, s, 9); return r;
  }
  NNT mul_nonnative(const NNT& a, const NNT& b, int mod) {          // nonnative.rs:390-464
    uint32_t s = emit(G_MUL, 35);
    emit(G_CHECKSUM, 16);
    prep2(OP_MUL, mod, s, a, b);
    mul_rec(a.r, b.r, s, mod);
    NNT r; r.r = mkref(SP_OUT, s, 9); return r;                     // r occupies the first 9 slots
  }
  void mul_rec(const Ref& x, const Ref& y, uint32_t out, int mod) {
    MulRec m; m.x = x; m.y = y; m.out = out; m.mod = (uint32_t)mod;
    t.muls.push_back(m);
  }
  Ref out9(uint32_t off) { return mkref(SP_OUT, off, 9); }
  // the x coordinate of a point id as a target (the factor of curve_double's x_squared)
  Ref coord_x(int32_t pid) {
    if (pid >= PID_CONST_BASE) {
      // constant_affine_point -> constant_nonnative (curve.rs:99-105, biguint.rs:165-175): leading zero limbs are dropped
      const uint32_t base = (uint32_t)(pid - PID_CONST_BASE) * 18;
      int nl = NL;
      while (nl > 0 && t.const_points[base + nl - 1] == 0) nl--;
      return mkref(SP_CPOINT, base, (uint8_t)nl);
    }
    if (pid >= PID_INPUT_BASE) return t.in_points[pid - PID_INPUT_BASE].x;
    return op_pts[pid].x.r;
  }
  NNT inv_nonnative(const NNT& x, int mod) {                        // nonnative.rs:502-536
    uint32_t s = emit(G_INV, 18);
    prep2(OP_INV, mod, s, x, x);
    NNT r; r.r = mkref(SP_OUT, s + 9, 9); return r;                 // [div, inv]
  }
  NNT add_many_nonnative(const std::vector<NNT>& xs, int mod) {     // nonnative.rs:310-353
    if (xs.size() == 1) return xs[0];
    uint32_t s = emit(G_ADDMANY, 10);
    PrepOp p;
    p.opcode = OP_ADDMANY;
    p.mod = (uint8_t)mod;
    p.pad = (uint8_t)xs.size();
    p.round = (uint8_t)cur_round;
    if ((uint32_t)cur_round + 1 > t.n_rounds) t.n_rounds = cur_round + 1;
    p.slot = s;
    p.a = mkref(SP_NONE, 0, 0);
    p.a.boff = (uint32_t)t.extra_refs.size();
    p.b = p.a;
    for (const NNT& x : xs) t.extra_refs.push_back(x.r);
    t.prep.push_back(p);
    NNT r; r.r = mkref(SP_OUT, s, 9); return r;
  }
  NNT if_nonnative(const BoolT& b, const NNT& x, const NNT& y, int mod) {   // nonnative.rs:297-308
    BoolT nb = b; nb.inv = !b.inv;
    return add_nonnative(mul_nonnative_by_bool(x, b), mul_nonnative_by_bool(y, nb), mod);
  }
  NNT neg_nonnative(const NNT& x, int mod) {                        // nonnative.rs:491-500
    NNT zero; zero.r = mkref(SP_CONST, 0, 0);
    return sub_nonnative(zero, x, mod);
  }
  NNT mul_nonnative_by_bool(const NNT& a, const BoolT& b) {         // nonnative.rs:278-295
    NNT r = a;
    r.r.bspace = b.space;
    r.r.boff = b.off;
    r.r.binv = b.inv ? 1 : 0;
    return r;
  }
  NNT nonnative_conditional_neg(const NNT& x, const BoolT& b, int mod) {   // nonnative.rs:584-596
    BoolT nb = b; nb.inv = !b.inv;
    NNT neg = neg_nonnative(x, mod);
    NNT xt = mul_nonnative_by_bool(neg, b);
    NNT xf = mul_nonnative_by_bool(x, nb);
    return add_nonnative(xt, xf, mod);
  }
  void connect_nonnative(const NNT& a, const NNT& b) {              // nonnative.rs:215-221
    Check c; c.a = a.r; c.b = b.r;
    t.checks.push_back(c);
  }

  // ---- curve gadgets
  int32_t const_point(const APoint& p) {                            // constant_affine_point, curve.rs:99-105
    int32_t idx = (int32_t)(t.const_points.size() / 18);
    for (int i = 0; i < NL; i++) t.const_points.push_back(p.x.l[i]);
    for (int i = 0; i < NL; i++) t.const_points.push_back(p.y.l[i]);
    return PID_CONST_BASE + idx;
  }
  PtT input_point(const NNT& x, const NNT& y) {
    InputPoint ip; ip.x = x.r; ip.y = y.r; ip.round = (uint32_t)cur_round;
    PtT p; p.pid = PID_INPUT_BASE + (int32_t)t.in_points.size(); p.x = x; p.y = y;
    t.in_points.push_back(ip);
    return p;
  }

  void curve_assert_valid(const NNT& px, const NNT& py) {           // curve.rs:123-135
    NNT a = constant_nonnative(fe_zero());
    NNT b = constant_nonnative(fe_small(7));
    NNT y2 = mul_nonnative(py, py, 0);
    NNT x2 = mul_nonnative(px, px, 0);
    NNT x3 = mul_nonnative(x2, px, 0);
    NNT ax = mul_nonnative(a, px, 0);
    NNT axb = add_nonnative(ax, b, 0);
    NNT rhs = add_nonnative(x3, axb, 0);
    connect_nonnative(y2, rhs);
  }

  // generator emission inside curve_add (curve.rs:202-223):
  // u SUB, v SUB, v_inv INV, s MUL+CS, s^2 MUL+CS, x_sum ADD, x3 SUB, x_diff SUB, prod MUL+CS, y3 SUB
  uint32_t emit_curve_add() {
    uint32_t s0 = emit(G_SUB, 10);
    emit(G_SUB, 10);
    emit(G_INV, 18);
    emit(G_MUL, 35); emit(G_CHECKSUM, 16);
    emit(G_MUL, 35); emit(G_CHECKSUM, 16);
    emit(G_ADD, 10);
    emit(G_SUB, 10);
    emit(G_SUB, 10);
    emit(G_MUL, 35); emit(G_CHECKSUM, 16);
    emit(G_SUB, 10);
    mul_rec(out9(s0), out9(s0 + 29), s0 + 38, 0);                   // s = u * v_inv
    mul_rec(out9(s0 + 38), out9(s0 + 38), s0 + 89, 0);              // s^2
    mul_rec(out9(s0 + 38), out9(s0 + 160), s0 + 170, 0);            // prod = s * x_diff
    return s0;
  }
  PtT point_of_op(int32_t k, uint32_t xoff, uint32_t yoff) {
    PtT p; p.pid = k;
    p.x.r = mkref(SP_OUT, xoff, 9);
    p.y.r = mkref(SP_OUT, yoff, 9);
    return p;
  }
  PtT push_ptop(PtOpType type, uint8_t flags, uint16_t aux, int32_t p1, int32_t p2, uint32_t slot, uint32_t xo, uint32_t yo) {
    PtOp o;
    o.type = type; o.aux = aux; o.p1 = p1; o.p2 = p2; o.slot = slot; o.round = (uint32_t)cur_round;
    if ((uint32_t)cur_round + 1 > t.n_rounds) t.n_rounds = cur_round + 1;
    // operand 2 is statically affine (constant or per-signature input point): mixed addition
    bool p2_affine = (type == PT_CADD_FIXED) ||
                     ((type == PT_ADD || type == PT_CADD_BOOL) && !(flags & PF_P2_CUR) && p2 >= PID_INPUT_BASE);
    o.flags = (uint8_t)(flags | (p2_affine ? PF_P2_AFFINE : 0) | (cur_chain << PF_CHAIN_SHIFT));
    int32_t k = (int32_t)t.ptops.size();
    if (t.chain_end[cur_chain] == t.chain_begin[cur_chain]) t.chain_begin[cur_chain] = k;
    t.chain_end[cur_chain] = k + 1;
    t.ptops.push_back(o);
    PtT res = point_of_op(k, slot + xo, slot + yo);
    op_pts.push_back(res);
    if (flags & PF_SET_CUR) cur_pt[cur_chain] = res;
    return res;
  }
  PtT curve_add(int32_t p1, int32_t p2, uint8_t flags) {            // curve.rs:202-223
    uint32_t s = emit_curve_add();
    return push_ptop(PT_ADD, flags, 0, p1, p2, s, 150, 221);
  }
  PtT curve_double(int32_t p1, uint8_t flags) {                     // curve.rs:160-185
    uint32_t s = emit(G_ADD, 10);                                    // double_y
    emit(G_INV, 18);                                                 // inv_double_y
    emit(G_MUL, 35); emit(G_CHECKSUM, 16);                           // x_squared
    emit(G_ADDMANY, 10);                                             // triple_xx_a
    emit(G_MUL, 35); emit(G_CHECKSUM, 16);                           // lambda
    emit(G_MUL, 35); emit(G_CHECKSUM, 16);                           // lambda_squared
    emit(G_ADD, 10);                                                 // x_double
    emit(G_SUB, 10);                                                 // x3
    emit(G_SUB, 10);                                                 // x_diff
    emit(G_MUL, 35); emit(G_CHECKSUM, 16);                           // lambda_x_diff
    emit(G_SUB, 10);                                                 // y3
    const Ref x1 = (flags & PF_P1_CUR) ? cur_pt[cur_chain].x.r : coord_x(p1);
    mul_rec(x1, x1, s + 28, 0);                                      // x_squared
    mul_rec(out9(s + 79), out9(s + 19), s + 89, 0);                  // lambda = triple_xx_a * inv_double_y
    mul_rec(out9(s + 89), out9(s + 89), s + 140, 0);                 // lambda_squared
    mul_rec(out9(s + 89), out9(s + 211), s + 221, 0);                // lambda_x_diff
    return push_ptop(PT_DBL, flags, 0, p1, -1, s, 201, 272);
  }
  PtT curve_conditional_add(PtOpType type, uint8_t flags, uint16_t aux, int32_t p1, int32_t p2 = -1) {   // curve.rs:225-243
    uint32_t s = emit_curve_add();
    emit(G_ADD, 10);
    emit(G_ADD, 10);
    return push_ptop(type, flags, aux, p1, p2, s, 231, 241);
  }

  // fixed_base_curve_mul_circuit (curve_fixed_base.rs:18-66), base = G
  PtT fixed_base_curve_mul(const NNT& scalar) {
    const int n_limbs4 = (9 * LB + 3) / 4;                           // 66 4-bit limbs (split_nonnative.rs:25-50)
    cur_chain = 0;
    t.fixed_scalar = scalar.r;
    t.fixed_round = (uint32_t)cur_round;
    t.fixed_table_base = (uint32_t)(t.const_points.size() / 18);
    APoint base = generator();
    for (int w = 0; w < n_limbs4; w++) {
      APoint acc = ap_inf();
      APoint first = base;
      for (int e = 0; e < 16; e++) {                                 // muls_point[e] = e * P_w, entry 0 := entry 1
        if (e == 0) {
          const_point(first);
        } else {
          acc = ap_add(base, acc);
          const_point(acc);
        }
      }
      for (int d = 0; d < 4; d++) base = ap_double(base);
    }
    int32_t rando_pid = const_point(rando);
    int32_t neg_pid = const_point(neg_rando);
    for (int w = 0; w < n_limbs4; w++) {
      uint8_t fl = PF_SET_CUR | (w ? PF_P1_CUR : 0);
      curve_conditional_add(PT_CADD_FIXED, fl, (uint16_t)w, rando_pid);
    }
    return curve_add(-1, neg_pid, PF_P1_CUR | PF_SET_CUR);
  }

  // curve_msm_circuit (curve_msm.rs:21-79)
  PtT curve_msm(const PtT& p, const PtT& q, const NNT& n, const NNT& m) {
    const int num_limbs = (n.r.nl * LB + 1) / 2;                     // 2-bit limbs: 73 for the 5-limb GLV halves, 131 for 9 limbs
    cur_chain = 1;
    t.msm_k1 = n.r;
    t.msm_k2 = m.r;
    t.msm_round = (uint32_t)cur_round;
    int32_t rando_pid = const_point(rando);
    int32_t neg_pid = const_point(neg_rando);
    int32_t pre[16];
    for (int i = 0; i < 16; i++) pre[i] = p.pid;
    int32_t cur_p = rando_pid, cur_q = rando_pid;
    for (int i = 0; i < 4; i++) {
      pre[i] = cur_p;
      pre[4 * i] = cur_q;
      cur_p = curve_add(cur_p, p.pid, 0).pid;
      cur_q = curve_add(cur_q, q.pid, 0).pid;
    }
    for (int i = 1; i < 4; i++) {
      pre[i] = curve_add(pre[i], neg_pid, 0).pid;
      pre[4 * i] = curve_add(pre[4 * i], neg_pid, 0).pid;
    }
    for (int i = 1; i < 4; i++)
      for (int j = 1; j < 4; j++) pre[i + 4 * j] = curve_add(pre[i], pre[4 * j], 0).pid;
    for (int i = 0; i < 16; i++) t.pre_pid[i] = pre[i];
    for (int d = num_limbs - 1; d >= 0; d--) {                       // .rev(): most significant digit first
      bool first = (d == num_limbs - 1);
      curve_double(first ? rando_pid : -1, PF_SET_CUR | (first ? 0 : PF_P1_CUR));
      curve_double(-1, PF_SET_CUR | PF_P1_CUR);
      curve_conditional_add(PT_CADD_MSM, PF_SET_CUR | PF_P1_CUR, (uint16_t)d, -1);
    }
    APoint spm = rando;
    for (int i = 0; i < 2 * num_limbs; i++) spm = ap_double(spm);
    int32_t to_add = const_point(ap_neg(spm));
    return curve_add(-1, to_add, PF_P1_CUR | PF_SET_CUR);
  }

  // decompose_secp256k1_scalar + glv_mul (glv.rs:53-104)
  PtT glv_mul(const NNT& px, const NNT& py, const NNT& k) {
    uint32_t gs = emit(G_GLV, 12);
    {
      PrepOp p;
      p.opcode = OP_GLV; p.mod = 1; p.pad = 0; p.round = (uint8_t)cur_round; p.slot = gs; p.a = k.r; p.b = k.r;
      t.prep.push_back(p);
    }
    NNT k1; k1.r = mkref(SP_OUT, gs, 5);
    NNT k2; k2.r = mkref(SP_OUT, gs + 5, 5);
    BoolT k1_neg = {SP_OUT, gs + 10, false};
    BoolT k2_neg = {SP_OUT, gs + 11, false};
    NNT k1_raw = nonnative_conditional_neg(k1, k1_neg, 1);
    NNT k2_raw = nonnative_conditional_neg(k2, k2_neg, 1);
    u64 sw[4], bw[4];
    memcpy(sw, GLV_S_W, 32);
    memcpy(bw, GLV_BETA_W, 32);
    NNT s = constant_nonnative(fe_from_u64x4(sw));
    NNT sbk = mul_nonnative(s, k2_raw, 1);
    sbk = add_nonnative(sbk, k1_raw, 1);
    connect_nonnative(sbk, k);
    NNT beta = constant_nonnative(fe_from_u64x4(bw));
    NNT beta_px = mul_nonnative(beta, px, 0);
    NNT pnegy = nonnative_conditional_neg(py, k1_neg, 0);            // curve_conditional_neg(p, k1_neg)
    NNT spnegy = nonnative_conditional_neg(py, k2_neg, 0);           // curve_conditional_neg(sp, k2_neg)
    PtT p_neg = input_point(px, pnegy);
    PtT sp_neg = input_point(beta_px, spnegy);
    return curve_msm(p_neg, sp_neg, k1, k2);
  }

  // verify_secp256k1_message_circuit (ecdsa.rs:30-53)
  void verify_secp256k1_message() {
    NNT msg = input(0), r = input(1), s = input(2), pkx = input(3), pky = input(4);
    curve_assert_valid(pkx, pky);
    NNT c = inv_nonnative(s, 1);
    NNT u1 = mul_nonnative(msg, c, 1);
    NNT u2 = mul_nonnative(r, c, 1);
    PtT point1 = fixed_base_curve_mul(u1);
    PtT point2 = glv_mul(pkx, pky, u2);
    cur_chain = 2;                                                   // joins both chains: runs after them
    PtT point = curve_add(point1.pid, point2.pid, 0);
    connect_nonnative(r, point.x);
  }
};

}  // namespace

static std::vector<Tape*>& user_tapes() {
  static std::vector<Tape*> v;
  return v;
}
static std::mutex g_tape_mu;

bool tape_exists(int id) {
  if (id >= 0 && id < N_TAPES) return true;
  std::lock_guard<std::mutex> lock(g_tape_mu);
  return id >= TAPE_USER_BASE && (size_t)(id - TAPE_USER_BASE) < user_tapes().size();
}

const Tape& get_tape(int id) {
  static Tape* tapes[N_TAPES] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::lock_guard<std::mutex> lock(g_tape_mu);
  if (id >= TAPE_USER_BASE && (size_t)(id - TAPE_USER_BASE) < user_tapes().size()) return *user_tapes()[id - TAPE_USER_BASE];
  if (id < 0 || id >= N_TAPES) id = 0;
  if (!tapes[id]) {
    Builder b;
    b.cur_chain = 1;
    switch (id) {
      case TAPE_ECDSA:
        b.verify_secp256k1_message();
        break;
      case TAPE_GLV_MUL: {
        NNT k = b.input(0), px = b.input(1), py = b.input(2);
        b.glv_mul(px, py, k);
        break;
      }
      case TAPE_CURVE_ADD: {
        PtT p1 = b.input_point(b.input(0), b.input(1)), p2 = b.input_point(b.input(2), b.input(3));
        b.curve_add(p1.pid, p2.pid, 0);
        break;
      }
      case TAPE_CURVE_DOUBLE: {
        PtT p1 = b.input_point(b.input(0), b.input(1));
        b.curve_double(p1.pid, 0);
        break;
      }
      case TAPE_FIXED_BASE_MUL: {
        b.fixed_base_curve_mul(b.input(0));
        break;
      }
      case TAPE_CURVE_MSM: {
        PtT p1 = b.input_point(b.input(0), b.input(1)), p2 = b.input_point(b.input(2), b.input(3));
        b.curve_msm(p1, p2, b.input(4), b.input(5));
        break;
      }
      case TAPE_CURVE_COND_ADD: {
        PtT p1 = b.input_point(b.input(0), b.input(1)), p2 = b.input_point(b.input(2), b.input(3));
        b.input(4);
        b.curve_conditional_add(PT_CADD_BOOL, 0, 4, p1.pid, p2.pid);
        break;
      }
    }
    tapes[id] = new Tape(b.t);
  }
  return *tapes[id];
}

// ------------------------------------------------------------------ builder C-API (include/p2e.h)
static std::vector<Builder*>& builders() {
  static std::vector<Builder*> v;
  return v;
}
static Builder* get_builder(int h) {
  if (h < 0 || (size_t)h >= builders().size() || !builders()[h]) {
    set_error_msg("p2e_builder: bad builder handle");
    return nullptr;
  }
  return builders()[h];
}
static int berr(Builder* b, const char* msg) {
  b->error = msg;
  set_error_msg(msg);
  return -1;
}
constexpr int MAX_ROUNDS = 32;
static int push_val(Builder* b, const NNT& v, int round) {
  b->vals.push_back(v);
  b->val_round.push_back((uint8_t)round);
  return (int)b->vals.size() - 1;
}
static int push_pt(Builder* b, const PtT& p, int chain, int round) {
  b->pts.push_back(p);
  b->pt_chain.push_back((int8_t)chain);
  b->pt_round.push_back((uint8_t)round);
  return (int)b->pts.size() - 1;
}
static bool val_ok(Builder* b, int v) { return v >= 0 && (size_t)v < b->vals.size(); }
static int vround(Builder* b, int v) { return b->val_round[v]; }
static int max2(int a, int c) { return a > c ? a : c; }
static bool pt_ok(Builder* b, int p) { return p >= 0 && (size_t)p < b->pts.size(); }
static bool mod_ok(int m) { return m == 0 || m == 1; }
static Fe fe_from_limbs(const uint32_t* l) {
  Fe r;
  for (int i = 0; i < NL; i++) r.l[i] = l[i] & LMASK;
  return r;
}
// round from which the COORDINATE VALUES of a point can feed nonnative generators / new points
static int pt_value_round(Builder* b, int p) { return b->pt_round[p] + (b->pt_chain[p] >= 0 ? 1 : 0); }
// chain of a new curve op of round `round`: chain 1 unless an operand was produced IN THE SAME ROUND by
// the fixed-base chain (0) or the join chain (2); results of earlier rounds are plain data
static int op_chain(Builder* b, int round, int p1, int p2) {
  int c = 1;
  for (int p : {p1, p2})
    if (p >= 0 && b->pt_round[p] == round && (b->pt_chain[p] == 0 || b->pt_chain[p] == 2)) c = 2;
  return c;
}
static bool round_ok(Builder* b, int round) {
  if (round < MAX_ROUNDS) return true;
  b->error = "p2e_builder: too many dependency rounds";
  set_error_msg(b->error.c_str());
  return false;
}

}  // namespace p2e

using namespace p2e;
static std::mutex g_builder_mu;
#define P2E_B(h)                                    \
  std::lock_guard<std::mutex> lock__(g_builder_mu); \
  Builder* b = get_builder(h);                      \
  if (!b) return -1

extern "C" {

// the circuit constant keccak(0u64) as a scalar (gadgets/curve_msm.rs:33-36), 4 little-endian 64-bit words: hosts that
// build generator streams of the reference's circuits (plonky2_ecdsa_b200.genstream.EcdsaCircuitBuilder) need it
int p2e_blinding_scalar(uint64_t* words4) {
  if (!words4) return set_error_msg("p2e_blinding_scalar: NULL");
  u64 k[4];
  blinding_scalar(k);
  for (int i = 0; i < 4; i++) words4[i] = k[i];
  return 0;
}

int p2e_builder_new(void) {
  std::lock_guard<std::mutex> lock(g_builder_mu);
  Builder* b = new Builder();
  b->cur_chain = 1;
  builders().push_back(b);
  return (int)builders().size() - 1;
}

int p2e_builder_free(int h) {
  P2E_B(h);
  delete b;
  builders()[h] = nullptr;
  return 0;
}

int p2e_builder_input(int h) {
  P2E_B(h);
  if (b->t.n_inputs >= (uint32_t)MAX_TAPE_INPUTS) return berr(b, "p2e_builder: at most 16 inputs per tape");
  return push_val(b, b->input(b->t.n_inputs), 0);
}

int p2e_builder_constant_nonnative(int h, const uint32_t* limbs9) {
  P2E_B(h);
  return push_val(b, b->constant_nonnative(fe_from_limbs(limbs9)), 0);
}

static int binop(int h, int kind, int modulus, int x, int y) {
  P2E_B(h);
  if (!mod_ok(modulus)) return berr(b, "p2e_builder: unknown modulus");
  if (!val_ok(b, x) || !val_ok(b, y)) return berr(b, "p2e_builder: operand is not a value handle");
  const int round = max2(vround(b, x), vround(b, y));
  if (!round_ok(b, round)) return -1;
  b->cur_round = round;
  NNT r = kind == 0 ? b->add_nonnative(b->vals[x], b->vals[y], modulus)
        : kind == 1 ? b->sub_nonnative(b->vals[x], b->vals[y], modulus)
                    : b->mul_nonnative(b->vals[x], b->vals[y], modulus);
  return push_val(b, r, round);
}
int p2e_builder_add_nonnative(int h, int modulus, int x, int y) { return binop(h, 0, modulus, x, y); }
int p2e_builder_sub_nonnative(int h, int modulus, int x, int y) { return binop(h, 1, modulus, x, y); }
int p2e_builder_mul_nonnative(int h, int modulus, int x, int y) { return binop(h, 2, modulus, x, y); }

int p2e_builder_inv_nonnative(int h, int modulus, int x) {
  P2E_B(h);
  if (!mod_ok(modulus) || !val_ok(b, x)) return berr(b, "p2e_builder: bad operand for inv_nonnative");
  if (b->vals[x].r.nl != NL) return berr(b, "p2e_builder: inv_nonnative needs a 9-limb operand");
  b->cur_round = vround(b, x);
  return push_val(b, b->inv_nonnative(b->vals[x], modulus), b->cur_round);
}

int p2e_builder_neg_nonnative(int h, int modulus, int x) {
  P2E_B(h);
  if (!mod_ok(modulus) || !val_ok(b, x)) return berr(b, "p2e_builder: bad operand for neg_nonnative");
  b->cur_round = vround(b, x);
  return push_val(b, b->neg_nonnative(b->vals[x], modulus), b->cur_round);
}

int p2e_builder_add_many_nonnative(int h, int modulus, const int* xs, int k) {
  P2E_B(h);
  if (!mod_ok(modulus) || k < 1 || k > 16) return berr(b, "p2e_builder: add_many_nonnative takes 1..16 summands");
  std::vector<NNT> v;
  int round = 0;
  for (int j = 0; j < k; j++) {
    if (!val_ok(b, xs[j])) return berr(b, "p2e_builder: bad summand for add_many_nonnative");
    v.push_back(b->vals[xs[j]]);
    round = max2(round, vround(b, xs[j]));
  }
  if (k == 1) return xs[0];
  b->cur_round = round;
  return push_val(b, b->add_many_nonnative(v, modulus), round);
}

// BoolTarget operands must be input values (wire 0 holds 0 / 1)
static bool bool_of(Builder* b, int v, BoolT& out) {
  if (!val_ok(b, v) || b->vals[v].r.space != SP_IN) return false;
  out.space = SP_IN;
  out.off = b->vals[v].r.off;
  out.inv = false;
  return true;
}

int p2e_builder_nonnative_conditional_neg(int h, int modulus, int x, int bool_input) {
  P2E_B(h);
  BoolT bt;
  if (!mod_ok(modulus) || !val_ok(b, x) || !bool_of(b, bool_input, bt))
    return berr(b, "p2e_builder: bad operand for nonnative_conditional_neg (the BoolTarget must be an input)");
  b->cur_round = vround(b, x);
  return push_val(b, b->nonnative_conditional_neg(b->vals[x], bt, modulus), b->cur_round);
}

int p2e_builder_if_nonnative(int h, int modulus, int bool_input, int x, int y) {
  P2E_B(h);
  BoolT bt;
  if (!mod_ok(modulus) || !val_ok(b, x) || !val_ok(b, y) || !bool_of(b, bool_input, bt))
    return berr(b, "p2e_builder: bad operand for if_nonnative (the BoolTarget must be an input)");
  b->cur_round = max2(vround(b, x), vround(b, y));
  return push_val(b, b->if_nonnative(bt, b->vals[x], b->vals[y], modulus), b->cur_round);
}

int p2e_builder_connect_nonnative(int h, int x, int y) {
  P2E_B(h);
  if (!val_ok(b, x) || !val_ok(b, y)) return berr(b, "p2e_builder: bad operand for connect_nonnative");
  b->connect_nonnative(b->vals[x], b->vals[y]);
  return 0;
}

int p2e_builder_affine_point(int h, int x, int y) {
  P2E_B(h);
  if (!val_ok(b, x) || !val_ok(b, y)) return berr(b, "p2e_builder: bad coordinates for affine_point");
  const int round = max2(vround(b, x), vround(b, y));
  if (!round_ok(b, round)) return -1;
  b->cur_round = round;
  return push_pt(b, b->input_point(b->vals[x], b->vals[y]), -1, round);
}

int p2e_builder_constant_affine_point(int h, const uint32_t* x9, const uint32_t* y9) {
  P2E_B(h);
  APoint a;
  a.x = fe_from_limbs(x9);
  a.y = fe_from_limbs(y9);
  a.inf = false;
  PtT p;
  p.pid = b->const_point(a);
  p.x = b->constant_nonnative(a.x);
  p.y = b->constant_nonnative(a.y);
  return push_pt(b, p, -1, 0);
}

int p2e_builder_point_x(int h, int p) {
  P2E_B(h);
  if (!pt_ok(b, p)) return berr(b, "p2e_builder: bad point handle");
  return push_val(b, b->pts[p].x, pt_value_round(b, p));
}
int p2e_builder_point_y(int h, int p) {
  P2E_B(h);
  if (!pt_ok(b, p)) return berr(b, "p2e_builder: bad point handle");
  return push_val(b, b->pts[p].y, pt_value_round(b, p));
}

int p2e_builder_curve_add(int h, int p1, int p2) {
  P2E_B(h);
  if (!pt_ok(b, p1) || !pt_ok(b, p2)) return berr(b, "p2e_builder: bad point handle");
  const int round = max2(b->pt_round[p1], b->pt_round[p2]);
  b->cur_round = round;
  b->cur_chain = op_chain(b, round, p1, p2);
  return push_pt(b, b->curve_add(b->pts[p1].pid, b->pts[p2].pid, 0), b->cur_chain, round);
}

int p2e_builder_curve_double(int h, int p) {
  P2E_B(h);
  if (!pt_ok(b, p)) return berr(b, "p2e_builder: bad point handle");
  const int round = b->pt_round[p];
  b->cur_round = round;
  b->cur_chain = op_chain(b, round, p, -1);
  return push_pt(b, b->curve_double(b->pts[p].pid, 0), b->cur_chain, round);
}

int p2e_builder_curve_conditional_add(int h, int p1, int p2, int bool_input) {
  P2E_B(h);
  if (!pt_ok(b, p1) || !pt_ok(b, p2)) return berr(b, "p2e_builder: bad point handle");
  if (!val_ok(b, bool_input) || b->vals[bool_input].r.space != SP_IN)
    return berr(b, "p2e_builder: the BoolTarget of curve_conditional_add must be an input value (wire 0 holds 0 / 1)");
  const int round = max2(b->pt_round[p1], b->pt_round[p2]);
  b->cur_round = round;
  b->cur_chain = op_chain(b, round, p1, p2);
  uint16_t aux = (uint16_t)(b->vals[bool_input].r.off / 9);
  return push_pt(b, b->curve_conditional_add(PT_CADD_BOOL, 0, aux, b->pts[p1].pid, b->pts[p2].pid), b->cur_chain, round);
}

int p2e_builder_fixed_base_curve_mul(int h, int scalar) {
  P2E_B(h);
  if (!val_ok(b, scalar)) return berr(b, "p2e_builder: bad scalar handle");
  if (b->n_fixed++) return berr(b, "p2e_builder: only one fixed_base_curve_mul_circuit per tape in this version");
  b->cur_round = vround(b, scalar);
  PtT r = b->fixed_base_curve_mul(b->vals[scalar]);     // sets cur_chain = 0
  return push_pt(b, r, 0, b->cur_round);
}

// a point operand of a scalar-multiplication loop must be complete before the loop's chain starts:
// if it is produced in the same round by the fixed-base / join chain, the loop moves to the next round
static int loop_round(Builder* b, int round, int p) {
  if (b->pt_round[p] == round && (b->pt_chain[p] == 0 || b->pt_chain[p] == 2)) return round + 1;
  return round;
}

int p2e_builder_curve_msm(int h, int p, int q, int n, int m) {
  P2E_B(h);
  if (!pt_ok(b, p) || !pt_ok(b, q) || !val_ok(b, n) || !val_ok(b, m)) return berr(b, "p2e_builder: bad handle for curve_msm");
  if (b->vals[n].r.nl != b->vals[m].r.nl) return berr(b, "p2e_builder: curve_msm scalars must have the same limb count");
  if (b->n_msm++) return berr(b, "p2e_builder: only one curve_msm_circuit / glv_mul per tape in this version");
  int round = max2(max2(b->pt_round[p], b->pt_round[q]), max2(vround(b, n), vround(b, m)));
  round = max2(loop_round(b, round, p), loop_round(b, round, q));
  if (!round_ok(b, round)) return -1;
  b->cur_round = round;
  PtT r = b->curve_msm(b->pts[p], b->pts[q], b->vals[n], b->vals[m]);    // sets cur_chain = 1
  return push_pt(b, r, 1, round);
}

int p2e_builder_glv_mul(int h, int p, int k) {
  P2E_B(h);
  if (!pt_ok(b, p) || !val_ok(b, k)) return berr(b, "p2e_builder: bad handle for glv_mul");
  if (b->n_msm++) return berr(b, "p2e_builder: only one curve_msm_circuit / glv_mul per tape in this version");
  const int round = max2(pt_value_round(b, p), vround(b, k));     // glv_mul works on the coordinate VALUES of p
  if (!round_ok(b, round)) return -1;
  b->cur_round = round;
  PtT r = b->glv_mul(b->pts[p].x, b->pts[p].y, b->vals[k]);
  return push_pt(b, r, 1, round);
}

int p2e_builder_curve_assert_valid(int h, int p) {
  P2E_B(h);
  if (!pt_ok(b, p)) return berr(b, "p2e_builder: bad point handle");
  b->cur_round = pt_value_round(b, p);
  if (!round_ok(b, b->cur_round)) return -1;
  b->curve_assert_valid(b->pts[p].x, b->pts[p].y);
  return 0;
}

int p2e_builder_finish(int h) {
  P2E_B(h);
  if (!b->error.empty()) return berr(b, "p2e_builder: an earlier builder call failed");
  std::lock_guard<std::mutex> lock(g_tape_mu);
  user_tapes().push_back(new Tape(b->t));
  return TAPE_USER_BASE + (int)user_tapes().size() - 1;
}

// Host-side resolution of one mul_nonnative instance's two factor references for ONE signature (a layout query with
// values: loads only, no arithmetic) -- what the range-check rows of the factors contain.  inputs: [n_inputs] pointers to 9
// wires each; witness: [n_slots] plain values of that signature.
int p2e_tape_mul_factors_host(int tape, int j, const uint64_t* const* inputs, const uint64_t* witness, uint64_t* x9, uint64_t* y9) {
  if (!tape_exists(tape)) return set_error_msg("unknown tape id");
  const Tape& t = get_tape(tape);
  if (j < 0 || j >= (int)t.muls.size()) return set_error_msg("tape_mul_factors_host: instance out of range");
  if (!witness || !x9 || !y9 || (t.n_inputs && !inputs)) return set_error_msg("tape_mul_factors_host: NULL argument");
  auto load = [&](uint8_t space, uint32_t off) -> uint64_t {
    switch (space) {
      case SP_OUT: return off < t.n_slots ? witness[off] : 0;
      case SP_IN: return off / 9 < t.n_inputs ? inputs[off / 9][off % 9] : 0;
      case SP_CONST: return off < t.const_limbs.size() ? t.const_limbs[off] : 0;
      case SP_CPOINT: return off < t.const_points.size() ? t.const_points[off] : 0;
      default: return 0;
    }
  };
  for (int f = 0; f < 2; f++) {
    const Ref& r = f ? t.muls[j].y : t.muls[j].x;
    uint64_t bm = 1;
    if (r.bspace != SP_NONE) bm = (load(r.bspace, r.boff) ^ (uint64_t)r.binv) & 1;
    uint64_t* o = f ? y9 : x9;
    for (int l = 0; l < 9; l++) o[l] = l < r.nl ? load(r.space, r.off + l) * bm : 0;
  }
  return 0;
}

}  // extern "C"
