// p2e_chain.cu -- witness generation for whole gadget chains (verify_secp256k1_message_circuit,
// glv_mul, ...) from the host-built tape (p2e_tape.h).
//
// The reference runs 312 Fermat ladders per signature, one per curve_add / curve_double
// (gadgets/curve.rs:160-223 -> gadgets/nonnative.rs:857-872).  Every one of those inverses is
// unique mod p, so it can be obtained differently without changing a single witness bit:
//   K0 k_prep          per signature: interpreter over the non-curve generators (prep ops)
//   K1 k_chain_replay  per (signature, chain), two lanes: replays the curve ops in Jacobian coordinates
//   K2 k_chain_invert_up / _mid / _down: ONE field inversion per signature for all Z's (Montgomery's trick, tree)
//   K3 k_chain_affine  per (signature, point): affine x, y (29-bit limbs) and Z^3
//   K4 k_chain_expand  per (signature, curve op): v_inv = Z1^3 Z2^3 / Z3 (add) or Z1^4 / Z3 (double),
//                      then every generator of that curve op (fast cores of p2e_f8.cuh / p2e_field.cuh)
//   K5 k_chain_checks  per signature: the gadget's connect_* constraints -> ERR_CONNECT
// K1-K3 and the v_inv products of K4 run on 8 saturated 32-bit words (p2e_f8.cuh); 29-bit limbs appear only where
// the witness has them.  K0 / K4 write either the plain witness (u64 per slot) or the compact host format directly
// (u32 narrow slots, CheckSum carries as low words + packed tops, check_sum columns optional): 261 instead of 661 KB
// of HBM writes per signature.  Signatures are the fastest-varying index everywhere (coalesced wire-major accesses).
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/p2e.h"
#include "p2e_common.h"
#include "p2e_f8.cuh"
#include "p2e_f8inv.cuh"
#include "p2e_field.cuh"
#include "p2e_tape.h"

namespace p2e {

constexpr int32_t PID_INPUT_BASE = 1 << 19;
constexpr int INV_GROUPS = 16;        // K2 tree: groups per signature (measured: 8 -> 0.29 ms, 16 -> 0.22 ms, 32 -> 0.25 ms)
constexpr int NW = 8;                 // words per field element in scratch

struct DevTape {
  const PrepOp* prep;
  const PtOp* ptops;
  const Check* checks;
  const InputPoint* in_points;
  const Ref* extra_refs;
  const u32* const_limbs;
  const u32* const_points;      // 18 limbs (x[9], y[9]) per constant point: the witness side
  const u32* const_points8;     // 16 words (x[8], y[8]) per constant point: the replay side
  const u32* cmap;              // compact layout: slot -> narrow / wide row (p2e_tape_compact_layout)
  const u32* pt_nrow;           // per curve op: narrow row of its first slot
  const u32* pt_wrow;           // per curve op: wide row of its first check_sum slot
  const MulRec* muls;           // every mul_nonnative instance, gadget order
  int n_muls;
  int n_prep, n_ptops, n_checks, n_in_points;
  int32_t pre_pid[16];
  Ref fixed_scalar, msm_k1, msm_k2;
  u32 fixed_round, msm_round;
  u32 fixed_table_base;
  int32_t chain_begin[N_CHAINS], chain_end[N_CHAINS];
};

enum : int { FMT_PLAIN = 0, FMT_COMPACT = 1 };

struct ChainArgs {
  DevTape t;
  size_t n;                 // signatures in this chunk
  int round;                // dependency round being executed (Tape::n_rounds)
  const u64* in[MAX_TAPE_INPUTS];   // input nonnative values, [9][is] each
  size_t is;
  int fmt;
  u64* out;                 // FMT_PLAIN: [n_slots][os]
  u32* nar;                 // FMT_COMPACT: [narrow rows][os]
  u64* wide;                // FMT_COMPACT: [wide rows][os] or NULL (derived form: check_sum columns not stored)
  size_t os;
  u32* err;
  // scratch (stride = cap)
  size_t cap;
  u32 *JX, *JY, *JZ;        // [n_ptops + n_in_points][8][cap]  weak words
  u32* ZINV;                // [n_ptops][8][cap]
  u32 *AX, *AY;             // [n_ptops + n_in_points][9][cap]  canonical 29-bit limbs
  u32* ZC;                  // [n_ptops + n_in_points][8][cap]  Z^3, weak words
  int32_t *SRC1, *SRC2;     // [n_ptops][cap]
  uint8_t* FLG;             // [n_ptops][cap]
  u32 *GP, *GI;             // [INV_GROUPS][8][cap]: group products / their inverses (K2 tree)
  u32* BAD;                 // [cap]: bit g set = group g contains a zero Z
};

// ------------------------------------------------------------------------------------------
// witness access (both output formats).  Refs only ever name plain value / flag slots (checked at upload).
__device__ __forceinline__ u64 load_wire(const ChainArgs& A, uint8_t space, u32 off, size_t i) {
  if (space == SP_OUT) {
    if (A.fmt == FMT_PLAIN) return A.out[(size_t)off * A.os + i];
    return (u64)A.nar[(size_t)A.t.cmap[off] * A.os + i];
  }
  if (space == SP_IN) return A.in[off / 9][(size_t)(off % 9) * A.is + i];
  return A.t.const_limbs[off];
}

__device__ __forceinline__ void load_ref(const ChainArgs& A, const Ref& r, size_t i, u64 (&w)[NL]) {
  u64 bm = 1;
  if (r.bspace != SP_NONE) bm = (load_wire(A, r.bspace, r.boff, i) ^ (u64)r.binv) & 1;
#pragma unroll
  for (int l = 0; l < NL; l++) w[l] = (l < r.nl) ? load_wire(A, r.space, r.off + l, i) * bm : 0;
}

// plain value slots written by the prep interpreter
__device__ __forceinline__ void st1(const ChainArgs& A, u32 slot, size_t i, u64 v) {
  if (A.fmt == FMT_PLAIN) A.out[(size_t)slot * A.os + i] = v;
  else A.nar[(size_t)A.t.cmap[slot] * A.os + i] = (u32)v;
}
__device__ __forceinline__ void st9(const ChainArgs& A, u32 slot, size_t i, const u32 (&v)[NL]) {
  if (A.fmt == FMT_PLAIN) {
#pragma unroll
    for (int l = 0; l < NL; l++) A.out[(size_t)(slot + l) * A.os + i] = v[l];
  } else {
    const u32 row = A.t.cmap[slot];                 // the rows of one tuple are consecutive
#pragma unroll
    for (int l = 0; l < NL; l++) A.nar[(size_t)(row + l) * A.os + i] = v[l];
  }
}
// MulNonnative row tuple (r, q, check_sum) followed by the CheckSum tuple (b): 51 slots, from the generic core
__device__ u32 st_mul_generic(const ChainArgs& A, u32 slot, size_t i, const u32 (&r)[NL], const u32 (&q)[NL],
                              const u64 (&cs)[2 * NL - 1], const u64 (&b)[2 * NL - 2]) {
  if (A.fmt == FMT_PLAIN) {
    u64* o = A.out + (size_t)slot * A.os + i;
    for (int l = 0; l < NL; l++) { o[(size_t)l * A.os] = r[l]; o[(size_t)(9 + l) * A.os] = q[l]; }
    for (int l = 0; l < 2 * NL - 1; l++) o[(size_t)(18 + l) * A.os] = cs[l];
    for (int l = 0; l < 2 * NL - 2; l++) o[(size_t)(35 + l) * A.os] = b[l];
    return 0;
  }
  u32 e = 0;
  u32* o = A.nar + (size_t)A.t.cmap[slot] * A.os + i;
  for (int l = 0; l < NL; l++) { o[(size_t)l * A.os] = r[l]; o[(size_t)(9 + l) * A.os] = q[l]; }
  u32 tops = 0;
  for (int l = 0; l < 2 * NL - 2; l++) {
    o[(size_t)(18 + l) * A.os] = (u32)b[l];
    tops |= ((u32)(b[l] >> 32) & 3u) << (2 * l);
    if (b[l] >> 34) e |= ERR_CARRY_RANGE;              // does not fit the packed form (the reference's assert, :527)
  }
  o[(size_t)34 * A.os] = tops;
  if (A.wide) {
    u64* w = A.wide + (size_t)(A.t.cmap[slot + 18] & 0x7fffffffu) * A.os + i;
    for (int l = 0; l < 2 * NL - 1; l++) w[(size_t)l * A.os] = cs[l];
  } else {
    // derived form: the host rebuilds check_sum_j = 2^29 c_j - c_{j-1} from the carries (CheckSumGate's constraints,
    // gates/mul_nonnative.rs:411-427, solved for a); a row that is not of this form is flagged
    i64 prev = 0;
    for (int l = 0; l < 2 * NL - 1; l++) {
      const i64 cur = l < 2 * NL - 2 ? (i64)b[l] - ((i64)1 << 33) : 0;
      if (gl_from_i64(cur * ((i64)1 << LB) - prev) != cs[l]) e |= ERR_CARRY_RANGE;
      prev = cur;
    }
  }
  return e;
}

// ------------------------------------------------------------------------------------------
// Emit policies of the expansion kernel: `f` = slot offset inside the curve op (plain format), `c` = narrow-row
// offset inside the curve op (compact format), `m` = index of the MulNonnative row inside the curve op.
struct W9 { u32 v[NL]; };
__device__ __forceinline__ W9 w9(const u32 (&a)[NL]) {
  W9 r;
#pragma unroll
  for (int l = 0; l < NL; l++) r.v[l] = a[l];
  return r;
}
struct EmPlain {
  u64* o;          // -> out[op.slot][i]
  size_t os;
  __device__ __forceinline__ void val(int f, int, const u32 (&v)[NL], u32 flag) const {
#pragma unroll
    for (int l = 0; l < NL; l++) o[(size_t)(f + l) * os] = v[l];
    o[(size_t)(f + 9) * os] = flag;
  }
  __device__ __forceinline__ void inv(int f, int, const u32 (&div)[NL], const u32 (&iv)[NL]) const {
#pragma unroll
    for (int l = 0; l < NL; l++) { o[(size_t)(f + l) * os] = div[l]; o[(size_t)(f + 9 + l) * os] = iv[l]; }
  }
  __device__ __forceinline__ void mul(int f, int, int, const u32 (&r)[NL], const u32 (&q)[NL], const i64 (&v)[17],
                                      const u32 (&blo)[16], u32 tops) const {
#pragma unroll
    for (int l = 0; l < NL; l++) { o[(size_t)(f + l) * os] = r[l]; o[(size_t)(f + 9 + l) * os] = q[l]; }
#pragma unroll
    for (int l = 0; l < 17; l++) o[(size_t)(f + 18 + l) * os] = gl_from_i64(v[l]);
#pragma unroll
    for (int l = 0; l < 16; l++) o[(size_t)(f + 35 + l) * os] = (u64)blo[l] | ((u64)((tops >> (2 * l)) & 3u) << 32);
  }
};
struct EmCompact {
  u32* o;          // -> nar[pt_nrow][i]
  u64* w;          // -> wide[pt_wrow][i] or NULL
  size_t os;
  __device__ __forceinline__ void val(int, int c, const u32 (&v)[NL], u32 flag) const {
#pragma unroll
    for (int l = 0; l < NL; l++) o[(size_t)(c + l) * os] = v[l];
    o[(size_t)(c + 9) * os] = flag;
  }
  __device__ __forceinline__ void inv(int, int c, const u32 (&div)[NL], const u32 (&iv)[NL]) const {
#pragma unroll
    for (int l = 0; l < NL; l++) { o[(size_t)(c + l) * os] = div[l]; o[(size_t)(c + 9 + l) * os] = iv[l]; }
  }
  __device__ __forceinline__ void mul(int, int c, int m, const u32 (&r)[NL], const u32 (&q)[NL], const i64 (&v)[17],
                                      const u32 (&blo)[16], u32 tops) const {
#pragma unroll
    for (int l = 0; l < NL; l++) { o[(size_t)(c + l) * os] = r[l]; o[(size_t)(c + 9 + l) * os] = q[l]; }
#pragma unroll
    for (int l = 0; l < 16; l++) o[(size_t)(c + 18 + l) * os] = blo[l];
    o[(size_t)(c + 34) * os] = tops;
    if (w) {
#pragma unroll
      for (int l = 0; l < 17; l++) w[(size_t)(17 * m + l) * os] = gl_from_i64(v[l]);
    }
  }
};

// ------------------------------------------------------------------------------------------
// K0: prep-op interpreter (generic cores: operands are arbitrary wires, both moduli).  The heavy cores are real
// functions: fully inlined the interpreter was 31 000 instructions (496 KB) of straight-line code, and a kernel that
// runs one warp per SM next to every other batch's kernels then costs the whole mix its instruction cache
// (r02f: 0.136 ms of the 0.79 ms pipelined step for 4.7 % of the instructions).
struct MulTuple {
  u32 r[NL], q[NL];
  u64 cs[2 * NL - 1], cb[2 * NL - 2];
};
template <class MOD>
__device__ __noinline__ u32 nl_prep_mul(const ChainArgs* A, u32 slot, size_t i, W9 a, W9 b) {
  MulTuple t;
  u32 e = mul_witness_core<MOD>(a.v, b.v, t.r, t.q, t.cs, t.cb);
  const u32 e2 = st_mul_generic(*A, slot, i, t.r, t.q, t.cs, t.cb);
  return e ? e : e2;
}
template <class MOD>
__device__ __noinline__ W9 nl_prep_inv(W9 a) {
  W9 r;
  inv_mod<MOD>(a.v, r.v);
  return r;
}
template <class MOD>
__device__ __noinline__ W9 nl_prep_inv_div(W9 a, W9 inv, u32* e) {
  W9 d;
  *e = inv_div_core<MOD>(a.v, inv.v, d.v);
  return d;
}

template <class MOD>
__device__ u32 prep_exec(const ChainArgs& A, const PrepOp& op, size_t i) {
  u64 aw[NL], bw[NL];
  load_ref(A, op.a, i, aw);
  load_ref(A, op.b, i, bw);
  u32 a[NL], b[NL], r[NL];
  u32 e = 0;
  switch (op.opcode) {
    case OP_MUL: {
      e = load_mul_wires(aw, a) | load_mul_wires(bw, b);
      if (e == 0) {
        e = nl_prep_mul<MOD>(&A, op.slot, i, w9(a), w9(b));
      } else {
        u32 z[NL];
        u64 cs[2 * NL - 1], cb[2 * NL - 2];
        set_zero9(z);
        for (int l = 0; l < 2 * NL - 1; l++) cs[l] = 0;
        for (int l = 0; l < 2 * NL - 2; l++) cb[l] = 0;
        st_mul_generic(A, op.slot, i, z, z, cs, cb);
      }
      break;
    }
    case OP_ADD:
    case OP_SUB: {
      e = load_canon<MOD>(aw, a) | load_canon<MOD>(bw, b);
      if (e) { set_zero9(a); set_zero9(b); }
      const u32 ov = op.opcode == OP_ADD ? add_witness_core<MOD>(a, b, r) : sub_witness_core<MOD>(a, b, r);
      st9(A, op.slot, i, r);
      st1(A, op.slot + 9, i, ov);
      break;
    }
    case OP_INV: {
      e = load_canon<MOD>(aw, a);
      if (e == 0 && is_zero9(a)) e = ERR_INVERSE_ZERO;
      u32 inv[NL], div[NL];
      set_zero9(inv);
      set_zero9(div);
      if (e == 0) {
        const W9 iv = nl_prep_inv<MOD>(w9(a));
        u32 e2;
        const W9 dv = nl_prep_inv_div<MOD>(w9(a), iv, &e2);
        e |= e2;
        copy9(iv.v, inv);
        copy9(dv.v, div);
      }
      st9(A, op.slot, i, div);
      st9(A, op.slot + 9, i, inv);
      break;
    }
    default: break;
  }
  return e;
}

__device__ __noinline__ u32 nl_prep_glv(W9 kc, W9* k1, W9* k2, u32* n1, u32* n2) {
  glv_decompose_core(kc.v, k1->v, k2->v, *n1, *n2);
  u32 hi = 0;
  for (int l = 5; l < NL; l++) hi |= k1->v[l] | k2->v[l];
  return hi ? (u32)ERR_LIMBS_OVERFLOW : 0u;
}

// NonNativeMultipleAddsGenerator (nonnative.rs:696-728) over op.pad summands
template <class MOD>
__device__ u32 prep_add_many(const ChainArgs& A, const PrepOp& op, size_t i) {
  u32 acc[NL];
  set_zero9(acc);
  u32 ov = 0, e = 0;
  for (int t = 0; t < op.pad; t++) {
    u64 w[NL];
    load_ref(A, A.t.extra_refs[op.a.boff + t], i, w);
    u32 v[NL], sum[NL];
    e |= load_canon<MOD>(w, v);
    add9(acc, v, sum);
    if (geq_mod<MOD>(sum)) { sub_mod9<MOD>(sum, acc); ov++; }
    else copy9(sum, acc);
  }
  if (e) { set_zero9(acc); ov = 0; }
  st9(A, op.slot, i, acc);
  st1(A, op.slot + 9, i, ov);
  return e;
}

__global__ void __launch_bounds__(32) k_prep(ChainArgs A) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  u32 err = 0;
  for (int k = 0; k < A.t.n_prep; k++) {
    const PrepOp op = A.t.prep[k];
    if (op.round != A.round) continue;
    if (op.opcode == OP_GLV) {
      u64 w[NL];
      load_ref(A, op.a, i, w);
      u32 kc[NL], n1 = 0, n2 = 0;
      W9 k1w, k2w;
      u32 e = load_canon<ModN>(w, kc);
      set_zero9(k1w.v);
      set_zero9(k2w.v);
      if (e == 0) e |= nl_prep_glv(w9(kc), &k1w, &k2w, &n1, &n2);
      const u32 (&k1)[NL] = k1w.v;
      const u32 (&k2)[NL] = k2w.v;
      for (int l = 0; l < 5; l++) {
        st1(A, op.slot + l, i, k1[l]);
        st1(A, op.slot + 5 + l, i, k2[l]);
      }
      st1(A, op.slot + 10, i, n1);
      st1(A, op.slot + 11, i, n2);
      err |= e;
    } else if (op.opcode == OP_ADDMANY) {
      err |= (op.mod == 0) ? prep_add_many<ModP>(A, op, i) : prep_add_many<ModN>(A, op, i);
    } else if (op.mod == 0) {
      err |= prep_exec<ModP>(A, op, i);
    } else {
      err |= prep_exec<ModN>(A, op, i);
    }
  }
  if (A.round == 0) A.err[i] = err;
  else if (err) A.err[i] |= err;
}

// ------------------------------------------------------------------------------------------
// K1: Jacobian replay (secp256k1: a = 0).  x = X/Z^2, y = Y/Z^3.
struct JacF {
  u32 X[NW], Y[NW], Z[NW];
};

__device__ __forceinline__ int pt_index(const ChainArgs& A, int32_t pid) {
  return pid < PID_INPUT_BASE ? pid : A.t.n_ptops + (pid - PID_INPUT_BASE);
}

// An explicit reference to the result of a curve_conditional_add means the SELECTED point
// (b ? sum : p1, curve.rs:235-240): follow the recorded operands while the flag is clear.
__device__ __forceinline__ int32_t resolve_pid(const ChainArgs& A, int32_t pid, size_t i) {
  while (pid >= 0 && pid < PID_INPUT_BASE) {
    const uint8_t ty = A.t.ptops[pid].type;
    if (ty == PT_ADD || ty == PT_DBL) break;
    if (A.FLG[(size_t)pid * A.cap + i]) break;
    pid = A.SRC1[(size_t)pid * A.cap + i];
  }
  return pid;
}

template <int N>
__device__ __forceinline__ void ldw(const u32* __restrict__ base, int idx, size_t cap, size_t i, u32 (&v)[N]) {
#pragma unroll
  for (int l = 0; l < N; l++) v[l] = base[((size_t)idx * N + l) * cap + i];
}
template <int N>
__device__ __forceinline__ void stw(u32* __restrict__ base, int idx, size_t cap, size_t i, const u32 (&v)[N]) {
#pragma unroll
  for (int l = 0; l < N; l++) base[((size_t)idx * N + l) * cap + i] = v[l];
}

__device__ __forceinline__ void load_jac(const ChainArgs& A, int32_t pid, size_t i, JacF& P) {
  if (pid >= PID_CONST_BASE) {
    const u32* c = A.t.const_points8 + (size_t)(pid - PID_CONST_BASE) * 16;
#pragma unroll
    for (int l = 0; l < NW; l++) {
      P.X[l] = c[l];
      P.Y[l] = c[NW + l];
      P.Z[l] = (l == 0) ? 1u : 0u;
    }
  } else {
    int idx = pt_index(A, pid);
    ldw(A.JX, idx, A.cap, i, P.X);
    ldw(A.JY, idx, A.cap, i, P.Y);
    ldw(A.JZ, idx, A.cap, i, P.Z);
  }
}

// The latency-bound kernels run few warps per SM through thousands of multiplications: the multiply / square
// bodies (f8m_v / f8s_v, p2e_f8inv.cuh) are real functions so that the loops stay inside the instruction cache.
__device__ __forceinline__ F8 f8w(const u32 (&a)[NW]) {
  F8 r;
#pragma unroll
  for (int l = 0; l < NW; l++) r.v[l] = a[l];
  return r;
}
__device__ __forceinline__ void fm(const u32 (&a)[NW], const u32 (&b)[NW], u32 (&r)[NW]) {
  F8 t = f8m_v(f8w(a), f8w(b));
#pragma unroll
  for (int l = 0; l < NW; l++) r[l] = t.v[l];
}
__device__ __forceinline__ void fs(const u32 (&a)[NW], u32 (&r)[NW]) {
  F8 t = f8s_v(f8w(a));
#pragma unroll
  for (int l = 0; l < NW; l++) r[l] = t.v[l];
}

// digit of `bits` bits at bit position `pos` of a little-endian 29-bit-limb value (nl limbs)
__device__ __forceinline__ u32 digit_at(const u32* limbs, int nl, int pos, int bits) {
  int l = pos / LB, sh = pos % LB;
  u32 lo = l < nl ? limbs[l] : 0u;
  u32 hi = (l + 1) < nl ? limbs[l + 1] : 0u;
  u64 v = ((u64)hi << LB) | lo;
  return (u32)(v >> sh) & ((1u << bits) - 1u);
}

__device__ __forceinline__ u32 load_scalar_limbs(const ChainArgs& A, const Ref& r, bool live, size_t i, u32* limbs) {
  // split_le_base::<2>(limb, 29) asserts each limb fits 29 bits (split_nonnative.rs:29-37);
  // `live`: the scalar-multiplication loop runs in this round (its scalar may not exist earlier)
  u32 e = 0;
  for (int l = 0; l < NL; l++) {
    u64 w = (live && l < r.nl && r.space != SP_NONE) ? load_wire(A, r.space, r.off + l, i) : 0;
    if (w >> LB) e |= ERR_LIMB_RANGE;
    limbs[l] = (u32)w & LMASK;
  }
  return e;
}

// Two-lane form: lanes 2j and 2j+1 serve the same (signature, chain).  Both keep every value in registers; at
// each step of the point formulas the two lanes compute the two products of an INDEPENDENT pair and swap the
// results with an 8-word shfl_xor, so a general addition is 8 sequential multiplications instead of 16 (mixed:
// 6 vs 11, doubling: 4 vs 7).  Static schedule, no shared memory; linear steps are done redundantly by both lanes.
template <int LANES>
__device__ __forceinline__ void duo_mul2(bool odd, const u32 (&a)[NW], const u32 (&b)[NW], u32 (&r1)[NW],
                                         const u32 (&c)[NW], const u32 (&d)[NW], u32 (&r2)[NW]) {
  if (LANES == 1) {                                  // one lane per (signature, chain): both products here, no exchange
    fm(a, b, r1);
    fm(c, d, r2);
    return;
  }
  F8 x, y;
#pragma unroll
  for (int l = 0; l < NW; l++) { x.v[l] = odd ? c[l] : a[l]; y.v[l] = odd ? d[l] : b[l]; }
  F8 mine = f8m_v(x, y);
#pragma unroll
  for (int l = 0; l < NW; l++) {
    u32 other = __shfl_xor_sync(0xffffffffu, mine.v[l], 1);
    r1[l] = odd ? other : mine.v[l];
    r2[l] = odd ? mine.v[l] : other;
  }
}
template <int LANES>
__device__ __forceinline__ void duo_sqr2(bool odd, const u32 (&a)[NW], u32 (&r1)[NW], const u32 (&c)[NW], u32 (&r2)[NW]) {
  if (LANES == 1) {
    fs(a, r1);
    fs(c, r2);
    return;
  }
  F8 x;
#pragma unroll
  for (int l = 0; l < NW; l++) x.v[l] = odd ? c[l] : a[l];
  F8 mine = f8s_v(x);
#pragma unroll
  for (int l = 0; l < NW; l++) {
    u32 other = __shfl_xor_sync(0xffffffffu, mine.v[l], 1);
    r1[l] = odd ? other : mine.v[l];
    r2[l] = odd ? mine.v[l] : other;
  }
}

// R = P + Q with Z3 = Z1*Z2*H, H = X2*Z1^2 - X1*Z2^2  (so x2 - x1 = Z3 / (Z1^3 Z2^3)).  Q_AFFINE: Z2 == 1.
template <int LANES, bool Q_AFFINE>
__device__ __forceinline__ void duo_add(bool odd, const JacF& P, const JacF& Q, JacF& R) {
  u32 z1z1[NW], u2[NW], s2[NW], h[NW], r[NW], t[NW], t2[NW], hh[NW], rr[NW], hhh[NW], v[NW], w[NW], v2[NW];
  if (Q_AFFINE) {
    fs(P.Z, z1z1);
    duo_mul2<LANES>(odd, Q.Y, P.Z, t2, Q.X, z1z1, u2);
    f8_sub(u2, P.X, h);
    duo_mul2<LANES>(odd, t2, z1z1, s2, h, h, hh);
    f8_sub(s2, P.Y, r);
    duo_mul2<LANES>(odd, r, r, rr, h, hh, hhh);
    duo_mul2<LANES>(odd, P.X, hh, v, P.Z, h, R.Z);          // Z3 = Z1 h
    f8_sub(rr, hhh, t);
    f8_add(v, v, v2);
    f8_sub(t, v2, R.X);                              // X3 = r^2 - hhh - 2v
    f8_sub(v, R.X, w);
    duo_mul2<LANES>(odd, r, w, t, P.Y, hhh, t2);
    f8_sub(t, t2, R.Y);                              // Y3 = r (v - X3) - Y1 hhh
  } else {
    u32 z2z2[NW], zz[NW], u1[NW], s1[NW];
    duo_sqr2<LANES>(odd, P.Z, z1z1, Q.Z, z2z2);
    duo_mul2<LANES>(odd, P.Y, Q.Z, t, Q.Y, P.Z, t2);
    duo_mul2<LANES>(odd, P.X, z2z2, u1, Q.X, z1z1, u2);
    duo_mul2<LANES>(odd, t, z2z2, s1, t2, z1z1, s2);
    f8_sub(u2, u1, h);
    f8_sub(s2, s1, r);
    duo_sqr2<LANES>(odd, h, hh, r, rr);
    duo_mul2<LANES>(odd, P.Z, Q.Z, zz, h, hh, hhh);
    duo_mul2<LANES>(odd, u1, hh, v, zz, h, R.Z);            // Z3 = Z1 Z2 h
    f8_sub(rr, hhh, t);
    f8_add(v, v, v2);
    f8_sub(t, v2, R.X);
    f8_sub(v, R.X, w);
    duo_mul2<LANES>(odd, r, w, t, s1, hhh, t2);
    f8_sub(t, t2, R.Y);
  }
}

// R = 2P with Z3 = 2*Y1*Z1  (so 2y = Z3 / Z1^4)
template <int LANES>
__device__ __forceinline__ void duo_dbl(bool odd, const JacF& P, JacF& R) {
  u32 xx[NW], yy[NW], yyyy[NW], s[NW], s2[NW], m[NW], t[NW], t2[NW], yz[NW], mm[NW], y8[NW];
  duo_sqr2<LANES>(odd, P.X, xx, P.Y, yy);
  f8_mul_small<3>(xx, m);              // M = 3 XX
  duo_mul2<LANES>(odd, yy, yy, yyyy, P.X, yy, t);
  f8_mul_small<4>(t, s);               // S = 4 X YY
  duo_mul2<LANES>(odd, P.Y, P.Z, yz, m, m, mm);
  f8_mul_small<8>(t, s2);              // 2 S
  f8_sub(mm, s2, R.X);                 // X3 = M^2 - 2S
  f8_sub(s, R.X, t);
  fm(m, t, t2);                        // both lanes (redundant): no partner product left
  f8_mul_small<8>(yyyy, y8);
  f8_sub(t2, y8, R.Y);                 // Y3 = M (S - X3) - 8 YYYY
  f8_add(yz, yz, R.Z);                 // Z3 = 2 Y1 Z1
}

// LANES = 2: the two-lane form above (half the dependency chain: single batches, small shards).  LANES = 1: one lane
// per (signature, chain) -- the same formulas without operand selects / shuffles / redundant linear steps (~25 % fewer
// warp instructions for twice the latency): calls of 4096 signatures or more, where the pipelined rate is bound by
// issued instructions (option "chain_replay_lanes").
template <int LANES>
__global__ void __launch_bounds__(64) k_chain_replay(ChainArgs A, int chain0) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool odd = LANES == 2 && (threadIdx.x & 1);
  const size_t i_raw = LANES == 2 ? tid >> 1 : tid;
  const bool valid = i_raw < A.n;
  const size_t i = valid ? i_raw : A.n - 1;       // the pair of an out-of-range task shadows the last signature
  const bool writer = valid && !odd;
  const int chain = chain0 + blockIdx.y;
  const int k_begin = A.t.chain_begin[chain], k_end = A.t.chain_end[chain];
  if (k_begin == k_end) return;
  u32 err = 0;
  if (writer) {
    // per-signature affine input points (Z = 1); every chain writes the same values (benign)
    for (int j = 0; j < A.t.n_in_points; j++) {
      if ((int)A.t.in_points[j].round != A.round) continue;
      u64 w[NL];
      u32 x[NL], y[NL], x8[NW], y8[NW], one[NW];
      load_ref(A, A.t.in_points[j].x, i, w);
      err |= load_canon<ModP>(w, x);
      load_ref(A, A.t.in_points[j].y, i, w);
      err |= load_canon<ModP>(w, y);
      f8_from_l29(x, x8);
      f8_from_l29(y, y8);
      f8_one(one);
      int idx = A.t.n_ptops + j;
      stw(A.JX, idx, A.cap, i, x8);
      stw(A.JY, idx, A.cap, i, y8);
      stw(A.JZ, idx, A.cap, i, one);
      stw(A.AX, idx, A.cap, i, x);
      stw(A.AY, idx, A.cap, i, y);
      stw(A.ZC, idx, A.cap, i, one);
    }
  }
  __syncwarp();                                    // the partner lane reads the input points just written
  u32 fsc[NL], k1[NL], k2[NL];
  err |= load_scalar_limbs(A, A.t.fixed_scalar, (int)A.t.fixed_round == A.round, i, fsc);
  err |= load_scalar_limbs(A, A.t.msm_k1, (int)A.t.msm_round == A.round, i, k1);
  err |= load_scalar_limbs(A, A.t.msm_k2, (int)A.t.msm_round == A.round, i, k2);

  int32_t cur = -1, resident = -2;
  JacF R;
  for (int k = k_begin; k < k_end; k++) {
    const PtOp op = A.t.ptops[k];
    if (((op.flags >> PF_CHAIN_SHIFT) & 3) != chain || (int)op.round != A.round) continue;   // chains / rounds interleave in op order
    int32_t p1 = (op.flags & PF_P1_CUR) ? cur : resolve_pid(A, op.p1, i);
    int32_t p2 = (op.flags & PF_P2_CUR) ? cur : resolve_pid(A, op.p2, i);
    u32 flag = 1;
    if (op.type == PT_CADD_FIXED) {
      u32 d = digit_at(fsc, NL, 4 * op.aux, 4);
      p2 = PID_CONST_BASE + (int32_t)(A.t.fixed_table_base + 16u * op.aux + d);
      flag = d != 0;
    } else if (op.type == PT_CADD_MSM) {
      u32 d = 4u * digit_at(k2, NL, 2 * op.aux, 2) + digit_at(k1, NL, 2 * op.aux, 2);
      p2 = A.t.pre_pid[d];
      flag = d != 0;
    } else if (op.type == PT_CADD_BOOL) {
      flag = load_wire(A, SP_IN, 9u * op.aux, i) != 0;      // BoolTarget input (0 / 1)
    }
    // operands: the previous result is still in registers; anything else comes from global scratch,
    // which only the even lane writes (made visible to the odd lane by the __syncwarp below)
    JacF P, Q;
    if (p1 == resident) P = R; else load_jac(A, p1, i, P);
    if (op.type == PT_DBL) {
      duo_dbl<LANES>(odd, P, R);
      p2 = p1;
    } else {
      if (p2 == resident) Q = R; else load_jac(A, p2, i, Q);
      if (op.flags & PF_P2_AFFINE) duo_add<LANES, true>(odd, P, Q, R);
      else duo_add<LANES, false>(odd, P, Q, R);
    }
    if (writer) {
      stw(A.JX, k, A.cap, i, R.X);
      stw(A.JY, k, A.cap, i, R.Y);
      stw(A.JZ, k, A.cap, i, R.Z);
      A.SRC1[(size_t)k * A.cap + i] = p1;
      A.SRC2[(size_t)k * A.cap + i] = p2;
      A.FLG[(size_t)k * A.cap + i] = (uint8_t)flag;
    }
    __syncwarp();
    resident = k;
    if (op.flags & PF_SET_CUR) cur = flag ? k : p1;
  }
  if (err && writer) atomicOr(&A.err[i], err);
}

// ------------------------------------------------------------------------------------------
// K2, tree form: ONE field inversion per signature.
//   up   (signature, group): prefix products of the group's Z's -> ZINV, group product -> GP
//   mid  (signature)       : prefix over the group products, one inversion, inverse of every group -> GI
//   down (signature, group): back-substitution inside the group
// 311 + (G + 270 + 2G) + 622 multiplications per signature instead of 933 + 270 G.
__device__ __forceinline__ bool group_range(const ChainArgs& A, int& k0, int& k1) {
  const int n = A.t.n_ptops;
  const int per = (n + (int)gridDim.y - 1) / (int)gridDim.y;
  k0 = blockIdx.y * per;
  k1 = (k0 + per < n) ? k0 + per : n;
  bool any = false;
  for (int k = k0; k < k1; k++) any |= (int)A.t.ptops[k].round == A.round;
  return any;
}

__global__ void __launch_bounds__(32) k_chain_invert_up(ChainArgs A) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  int k0, k1;
  const bool any = group_range(A, k0, k1);
  F8 run, z;
  f8_one(run.v);
  if (any) {
    for (int k = k0; k < k1; k++) {
      if ((int)A.t.ptops[k].round != A.round) continue;
      stw(A.ZINV, k, A.cap, i, run.v);    // prefix product of this group's Z's
      ldw(A.JZ, k, A.cap, i, z.v);
      run = f8m_v(run, z);
    }
  }
  stw(A.GP, blockIdx.y, A.cap, i, run.v);
}

// inverse of a weak base-field value by division steps (p2e_modinv.cuh): canonical in, canonical out
static __device__ __noinline__ F8 nl_modinv_p(F8 a) {
  F8 r;
  f8_canon(a.v);
  modinv_words<ModP>(a.v, r.v);
  return r;
}

__global__ void __launch_bounds__(32) k_chain_invert_mid(ChainArgs A, int G) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  F8 run, gp;
  f8_one(run.v);
  for (int g = 0; g < G; g++) {
    stw(A.GI, g, A.cap, i, run.v);        // product of the groups before g
    ldw(A.GP, g, A.cap, i, gp.v);
    run = f8m_v(run, gp);
  }
  u32 bad = 0;
  if (f8_is_zero(run.v)) {
    // some curve op had x1 == x2 (or y == 0): the reference's inverse() panics.  Find the groups that
    // hold a zero, leave them out of the product (their ZINV rows are zeroed by the down pass).
    f8_one(run.v);
    for (int g = 0; g < G; g++) {
      stw(A.GI, g, A.cap, i, run.v);
      ldw(A.GP, g, A.cap, i, gp.v);
      if (f8_is_zero(gp.v)) { bad |= 1u << g; continue; }
      run = f8m_v(run, gp);
    }
    atomicOr(&A.err[i], (u32)ERR_INVERSE_ZERO);
  }
  A.BAD[i] = bad;
  F8 inv = nl_modinv_p(run);
  for (int g = G - 1; g >= 0; g--) {
    F8 pre;
    ldw(A.GI, g, A.cap, i, pre.v);
    F8 gi = f8m_v(inv, pre);
    stw(A.GI, g, A.cap, i, gi.v);
    if ((bad >> g) & 1) continue;
    ldw(A.GP, g, A.cap, i, gp.v);
    inv = f8m_v(inv, gp);
  }
}

__global__ void __launch_bounds__(32) k_chain_invert_down(ChainArgs A) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  int k0, k1;
  if (!group_range(A, k0, k1)) return;
  if ((A.BAD[i] >> blockIdx.y) & 1) {
    u32 zero[NW];
#pragma unroll
    for (int l = 0; l < NW; l++) zero[l] = 0;
    for (int k = k0; k < k1; k++)
      if ((int)A.t.ptops[k].round == A.round) stw(A.ZINV, k, A.cap, i, zero);
    return;
  }
  F8 inv, z;
  ldw(A.GI, blockIdx.y, A.cap, i, inv.v);
  for (int k = k1 - 1; k >= k0; k--) {
    if ((int)A.t.ptops[k].round != A.round) continue;
    F8 pre;
    ldw(A.ZINV, k, A.cap, i, pre.v);
    ldw(A.JZ, k, A.cap, i, z.v);
    F8 zi = f8m_v(inv, pre);
    stw(A.ZINV, k, A.cap, i, zi.v);
    inv = f8m_v(inv, z);
  }
}

// K3: canonical affine coordinates (29-bit limbs) of every computed point, and Z^3 (weak words; K4 folds it into
// its own product and canonicalises once)
__global__ void __launch_bounds__(128) k_chain_affine(ChainArgs A) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int k = blockIdx.y;
  if (i >= A.n) return;
  // all four loads are addressed by (k, i) alone: issue them together, ahead of the tape lookup
  u32 zi[NW], zi2[NW], zi3[NW], jx[NW], jy[NW], jz[NW], r[NW];
  u32 c[NL];
  ldw(A.ZINV, k, A.cap, i, zi);
  ldw(A.JX, k, A.cap, i, jx);
  ldw(A.JY, k, A.cap, i, jy);
  ldw(A.JZ, k, A.cap, i, jz);
  if ((int)A.t.ptops[k].round != A.round) return;
  fs(zi, zi2);
  fm(zi2, zi, zi3);
  fm(jx, zi2, r);
  f8_to_canon_l29(r, c);
  stw(A.AX, k, A.cap, i, c);
  fm(jy, zi3, r);
  f8_to_canon_l29(r, c);
  stw(A.AY, k, A.cap, i, c);
  fs(jz, zi2);
  fm(zi2, jz, r);
  stw(A.ZC, k, A.cap, i, r);
}

// ------------------------------------------------------------------------------------------
// K4: expand one curve op into all of its generators' output tuples.  Every operand here is a canonical generator
// output (< 2^256), so the fast cores apply.  The cores are kept OUT OF LINE: fully inlined, a curve op is tens of
// KB of straight-line code per thread and the kernel becomes instruction-fetch bound.
struct W9E {
  W9 r;
  u32 e;
};
// MulNonnative row + CheckSum row; the tuple is written from inside the call (only r and the error word return)
template <class EM>
__device__ __noinline__ W9E nl_mulw(EM em, int f, int c, int m, W9 x, W9 y) {
  u32 r[NL], q[NL], blo[16], tops;
  i64 v[17];
  W9E ret;
  ret.e = mulw_p_fast(x.v, y.v, r, q, v, blo, tops);
  em.mul(f, c, m, r, q, v, blo, tops);
#pragma unroll
  for (int l = 0; l < NL; l++) ret.r.v[l] = r[l];
  return ret;
}
// add / sub tuples are small (~80 instructions): inlined, the ~30 register moves of a by-value call disappear
// (r02b: 0.807 -> 0.789 ms per batch); 3 CTAs per SM: 4 (128 registers) is faster alone but not in the pipelined mix
template <class EM, bool IS_SUB>
__device__ __forceinline__ W9 nl_addsub(EM em, int f, int c, W9 a, W9 b) {
  W9 r;
  const u32 ov = IS_SUB ? sub_witness_core<ModP>(a.v, b.v, r.v) : add_witness_core<ModP>(a.v, b.v, r.v);
  em.val(f, c, r.v, ov);
  return r;
}
// NonNativeInverseGenerator tuple when the inverse is known: div = (x * inv - 1) / p, checked
template <class EM>
__device__ __noinline__ u32 nl_inv_known(EM em, int f, int c, W9 xc, F8 inv8, W9 inv) {
  u32 x8[NW], t[16], r8[NW], q8[NW], div[NL];
  f8_from_l29(xc.v, x8);
  f8_mul_wide(x8, inv8.v, t);
  u32 bad = f8_divmod16(t, r8, q8);
  bad |= r8[0] ^ 1u;
#pragma unroll
  for (int l = 1; l < NW; l++) bad |= r8[l];
  f8_to_l29(q8, div);
  em.inv(f, c, div, inv.v);
  return bad ? (u32)ERR_INTERNAL : 0u;
}

template <class EM>
__device__ __forceinline__ u32 do_mul(const EM& em, int f, int c, int m, const u32 (&x)[NL], const u32 (&y)[NL], u32 (&r)[NL]) {
  W9E t = nl_mulw<EM>(em, f, c, m, w9(x), w9(y));
#pragma unroll
  for (int l = 0; l < NL; l++) r[l] = t.r.v[l];
  return t.e;
}
template <class EM>
__device__ __forceinline__ void do_add(const EM& em, int f, int c, const u32 (&a)[NL], const u32 (&b)[NL], u32 (&r)[NL]) {
  W9 t = nl_addsub<EM, false>(em, f, c, w9(a), w9(b));
#pragma unroll
  for (int l = 0; l < NL; l++) r[l] = t.v[l];
}
template <class EM>
__device__ __forceinline__ void do_sub(const EM& em, int f, int c, const u32 (&a)[NL], const u32 (&b)[NL], u32 (&r)[NL]) {
  W9 t = nl_addsub<EM, true>(em, f, c, w9(a), w9(b));
#pragma unroll
  for (int l = 0; l < NL; l++) r[l] = t.v[l];
}
template <class EM>
__device__ __forceinline__ u32 do_inv_known(const EM& em, int f, int c, const u32 (&xc)[NL], const u32 (&vinv8)[NW],
                                            const u32 (&vinv)[NL]) {
  return nl_inv_known<EM>(em, f, c, w9(xc), f8w(vinv8), w9(vinv));
}

__device__ __forceinline__ void load_affine(const ChainArgs& A, int32_t pid, size_t i, u32 (&x)[NL], u32 (&y)[NL],
                                            u32 (&zc)[NW]) {
  if (pid >= PID_CONST_BASE) {
    const u32* c = A.t.const_points + (size_t)(pid - PID_CONST_BASE) * 18;
#pragma unroll
    for (int l = 0; l < NL; l++) {
      x[l] = c[l];
      y[l] = c[NL + l];
    }
    f8_one(zc);
  } else {
    int idx = pt_index(A, pid);
    ldw(A.AX, idx, A.cap, i, x);
    ldw(A.AY, idx, A.cap, i, y);
    ldw(A.ZC, idx, A.cap, i, zc);
  }
}

template <class EM>
__device__ __forceinline__ u32 expand_op(const ChainArgs& A, const EM& em, const PtOp& op, int k, size_t i) {
  // Loads are issued in two dependent rounds: everything addressed by (k, i) alone first ...
  const size_t ki = (size_t)k * A.cap + i;
  const int32_t p1 = A.SRC1[ki];
  const int32_t p2_any = A.SRC2[ki];              // meaningless for a doubling (never dereferenced then)
  const u32 flg = A.FLG[ki];
  u32 zinv[NW];
  ldw(A.ZINV, k, A.cap, i, zinv);
  u32 e = 0;
  u32 x1[NL], y1[NL], vinv[NL], x2[NL], y2[NL];
  u32 zc1[NW], zc2[NW], t[NW], vinv8[NW];
  // ... then both operands
  load_affine(A, p1, i, x1, y1, zc1);
  if (op.type != PT_DBL) load_affine(A, p2_any, i, x2, y2, zc2);
  if (op.type == PT_DBL) {
    // (2y)^-1 = Z1^4 / Z3
    u32 z1[NW];
    if (p1 >= PID_CONST_BASE) f8_one(z1);
    else ldw(A.JZ, pt_index(A, p1), A.cap, i, z1);
    fm(zc1, z1, t);
    fm(t, zinv, vinv8);
    f8_canon(vinv8);
    f8_to_l29(vinv8, vinv);
    u32 dy[NL], xx[NL], tr[NL], lam[NL], lam2[NL], xd[NL], x3[NL], xdiff[NL], lx[NL], y3[NL];
    do_add(em, 0, 0, y1, y1, dy);                                 // double_y
    canon9<ModP>(dy);
    if (is_zero9(dy)) e |= ERR_INVERSE_ZERO;
    e |= do_inv_known(em, 10, 10, dy, vinv8, vinv);               // inv_double_y
    e |= do_mul(em, 28, 28, 0, x1, x1, xx);                       // x_squared
    {                                                             // add_many([xx, xx, xx, A = 0])
      u32 acc[NL], sum[NL], ov = 0;
      copy9(xx, acc);
#pragma unroll
      for (int rep = 0; rep < 2; rep++) {
        add9(acc, xx, sum);
        if (geq_mod<ModP>(sum)) { sub_mod9<ModP>(sum, acc); ov++; }
        else copy9(sum, acc);
      }
      copy9(acc, tr);
      em.val(79, 63, tr, ov);
    }
    e |= do_mul(em, 89, 73, 1, tr, vinv, lam);                    // lambda
    e |= do_mul(em, 140, 108, 2, lam, lam, lam2);                 // lambda_squared
    do_add(em, 191, 143, x1, x1, xd);                             // x_double
    canon9<ModP>(xd);
    do_sub(em, 201, 153, lam2, xd, x3);                           // x3
    do_sub(em, 211, 163, x1, x3, xdiff);                          // x_diff
    e |= do_mul(em, 221, 173, 3, lam, xdiff, lx);                 // lambda_x_diff
    do_sub(em, 272, 208, lx, y1, y3);                             // y3
  } else {
    // (x2 - x1)^-1 = Z1^3 Z2^3 / Z3
    fm(zc1, zc2, t);
    fm(t, zinv, vinv8);
    f8_canon(vinv8);
    f8_to_l29(vinv8, vinv);
    u32 u[NL], v[NL], s[NL], s2[NL], xs[NL], x3[NL], xd[NL], pr[NL], y3[NL];
    do_sub(em, 0, 0, y2, y1, u);                                  // u
    do_sub(em, 10, 10, x2, x1, v);                                // v
    if (is_zero9(v)) e |= ERR_INVERSE_ZERO;
    e |= do_inv_known(em, 20, 20, v, vinv8, vinv);                // v_inv
    e |= do_mul(em, 38, 38, 0, u, vinv, s);                       // s
    e |= do_mul(em, 89, 73, 1, s, s, s2);                         // s_squared
    do_add(em, 140, 108, x2, x1, xs);                             // x_sum
    canon9<ModP>(xs);
    do_sub(em, 150, 118, s2, xs, x3);                             // x3
    do_sub(em, 160, 128, x1, x3, xd);                             // x_diff
    e |= do_mul(em, 170, 138, 2, s, xd, pr);                      // prod
    do_sub(em, 221, 173, pr, y1, y3);                             // y3
    if (op.type != PT_ADD) {
      // curve_conditional_add: x = sum.x * b + p1.x * !b (curve.rs:235-240)
      u32 z[NL], r[NL];
      set_zero9(z);
      if (flg) { do_add(em, 231, 183, x3, z, r); do_add(em, 241, 193, y3, z, r); }
      else { do_add(em, 231, 183, z, x1, r); do_add(em, 241, 193, z, y1, r); }
    }
  }
  return e;
}

template <int FMT>
#ifndef P2E_EXPAND_MINB
#define P2E_EXPAND_MINB 3
#endif
__global__ void __launch_bounds__(128, P2E_EXPAND_MINB) k_chain_expand(ChainArgs A) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int k = blockIdx.y;
  if (i >= A.n) return;
  const PtOp op = A.t.ptops[k];
  if ((int)op.round != A.round) return;
  u32 e;
  if (FMT == FMT_PLAIN) {
    EmPlain em{A.out + (size_t)op.slot * A.os + i, A.os};
    e = expand_op(A, em, op, k, i);
  } else {
    EmCompact em{A.nar + (size_t)A.t.pt_nrow[k] * A.os + i, A.wide ? A.wide + (size_t)A.t.pt_wrow[k] * A.os + i : nullptr, A.os};
    e = expand_op(A, em, op, k, i);
  }
  if (e) atomicOr(&A.err[i], e);
}

// ------------------------------------------------------------------------------------------
// The plonky2_ux range-check rows under the in-tree generators.  Every range_check_ux_circuit call of the crate sits in
// gadgets/nonnative.rs: add_many_nonnative (:330-331: the sum's limbs, then the overflow as one input), sub_nonnative
// (:373: the difference's limbs) and mul_nonnative (:454-460: x, y, r, q, r again -- 29-bit inputs, as many as the target
// has limbs -- then b[0..8], b[8..16] of the CheckSum row as 34-bit inputs).  Each call is ONE gate row.  The call list
// follows from the tape (generator list + Tape::muls); one thread per (signature, call) reads the values from the
// finished witness (either device format) and writes the gate row of the restated spec (p2e_ux.cu: inputs first, then
// ceil(bits / lb) aux limbs of lb bits per input; DESIGN.md section 9, PARITY UNPINNED like every plonky2_ux row).
enum RcKind : uint8_t { RC_REF = 0, RC_SLOTS = 1, RC_CARRIES = 2 };
struct RcCall {
  Ref r;               // RC_REF: a mul factor
  u32 slot;            // RC_SLOTS: first witness slot; RC_CARRIES: first carry slot of the CheckSum tuple (plain numbering)
  u32 tuple;           // RC_CARRIES: first slot of the MulNonnative tuple (compact rows: low words at +18, tops at +34)
  uint8_t kind, count, bits, first;   // first: index of the first carry (0 or 8)
};
struct RcArgs {
  const RcCall* calls;
  const u64* wire_off;  // [calls + 1] wire offset of each call's row for this aux-limb width
  int c0, lb;
  u64 base;             // wire_off[c0]
  u64* rows;
  size_t rs;
};
__global__ void __launch_bounds__(128) k_tape_range_check_rows(ChainArgs A, RcArgs M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const int c = M.c0 + blockIdx.y;
  const RcCall call = M.calls[c];
  const int count = call.count, n_aux = (call.bits + M.lb - 1) / M.lb;
  u64* row = M.rows + (size_t)(M.wire_off[c] - M.base) * M.rs + i;
  u64 v[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) v[l] = 0;
  if (call.kind == RC_REF) {
    const Ref& r = call.r;
    u64 bm = 1;
    if (r.bspace != SP_NONE) bm = (load_wire(A, r.bspace, r.boff, i) ^ (u64)r.binv) & 1;
    for (int l = 0; l < count; l++) v[l] = r.space == SP_CPOINT ? (u64)A.t.const_points[r.off + l] : load_wire(A, r.space, r.off + l, i) * bm;
  } else if (call.kind == RC_SLOTS) {
    for (int l = 0; l < count; l++) v[l] = load_wire(A, SP_OUT, call.slot + l, i);
  } else if (A.fmt == FMT_PLAIN) {
    for (int l = 0; l < count; l++) v[l] = A.out[(size_t)(call.slot + l) * A.os + i];
  } else {
    // compact: 16 low words after r and q, then one word with the 16 two-bit tops
    const u32 r0 = A.t.cmap[call.tuple];
    const u32 tops = A.nar[(size_t)(r0 + 34) * A.os + i];
    for (int l = 0; l < count; l++)
      v[l] = (u64)A.nar[(size_t)(r0 + 18 + call.first + l) * A.os + i] | ((u64)((tops >> (2 * (call.first + l))) & 3u) << 32);
  }
  u32 e = 0;
  const u64 mask = (1ull << M.lb) - 1;
  for (int u = 0; u < count; u++) row[(size_t)u * M.rs] = v[u];
  for (int u = 0; u < count; u++) {
    u64 t = v[u];
    for (int a = 0; a < n_aux; a++) {
      row[(size_t)(count + n_aux * u + a) * M.rs] = t & mask;
      t >>= M.lb;
    }
    if (t) e |= ERR_LIMB_RANGE;                      // does not fit its aux limbs: the row is unsatisfiable
  }
  if (e) atomicOr(&A.err[i], e);
}

// K5: connect_nonnative checks
__global__ void __launch_bounds__(32) k_chain_checks(ChainArgs A) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  u32 bad = 0;
  for (int c = 0; c < A.t.n_checks; c++) {
    u64 a[NL], b[NL];
    load_ref(A, A.t.checks[c].a, i, a);
    load_ref(A, A.t.checks[c].b, i, b);
    for (int l = 0; l < NL; l++) if (a[l] != b[l]) bad = 1;
  }
  if (bad) A.err[i] |= ERR_CONNECT;
}

// diagnostic runs that skip kernels (option chain_phases != 63) never return a usable witness: every row is flagged
__global__ void k_chain_poison(size_t n, u32* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) err[i] |= (u32)ERR_INTERNAL;
}

// test hook: the PTX bodies of p2e_f8.cuh on the device (op codes of tests/hostcheck hc_f8)
__global__ void k_debug_f8(int op, size_t n, const u32* __restrict__ a, const u32* __restrict__ b, u32* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 x[NW], y[NW], r[NW];
  for (int l = 0; l < NW; l++) { x[l] = a[9 * i + l]; y[l] = b[9 * i + l]; r[l] = 0; }
  if (op == 0) f8_mul(x, y, r);
  else if (op == 1) f8_sqr(x, r);
  else if (op == 2) f8_add(x, y, r);
  else if (op == 3) f8_sub(x, y, r);
  else if (op == 4) f8_mul_small<2>(x, r);
  else if (op == 5) f8_mul_small<3>(x, r);
  else if (op == 6) f8_mul_small<4>(x, r);
  else if (op == 7) f8_mul_small<8>(x, r);
  else if (op == 8) { f8_copy(x, r); f8_canon(r); }
  else if (op == 11) { F8 t = nl_modinv_p(f8w(x)); f8_copy(t.v, r); }       // division steps (what K2 runs)
  else if (op == 14) { F8 t = f8_inv_calls(f8w(x)); f8_copy(t.v, r); }      // the ladder, as a cross-check
  else if (op == 15) { u32 c[8]; f8_copy(x, c); modinv_words<ModN>(c, r); }   // division steps modulo the group order (x < n)
  else if (op == 12 || op == 13) {                     // divmod16 of the wide product: 12 -> remainder, 13 -> quotient
    u32 t[16], rr[NW], qq[NW];
    f8_mul_wide(x, y, t);
    f8_divmod16(t, rr, qq);
    f8_copy(op == 12 ? rr : qq, r);
  }
  for (int l = 0; l < NW; l++) out[9 * i + l] = r[l];
  out[9 * i + 8] = 0;
}

// ------------------------------------------------------------------------------------------
// host side: per-device state (device copies of the tapes, scratch arenas, launch sequence)
struct TapeOnDevice {
  bool ready = false;
  DevTape d;
  void* blob = nullptr;
  size_t n_slots = 0;
  u32 n_narrow = 0, n_wide = 0;
  // range_check_ux_circuit calls (built on first use): device call list, per aux-limb width the wire offsets
  RcCall* rc_calls = nullptr;
  int n_rc = 0;
  std::map<int, u64*> rc_off;
};

// One scratch arena per launching stream, so independent batches enqueued on different streams
// can overlap (the per-signature kernels K0-K2 are latency-bound and leave most SMs idle).
struct Arena {
  cudaStream_t stream = nullptr;
  void* ptr = nullptr;
  size_t bytes = 0;
  bool used = false;
  uint64_t last_use = 0;
  // internal streams: the latency-bound per-signature kernels (K0-K2) run at high priority so their
  // few small blocks are scheduled ahead of another batch's bandwidth-bound expansion (K3-K5)
  cudaStream_t hi = nullptr, lo = nullptr;
  cudaEvent_t ev_in = nullptr, ev_mid = nullptr, ev_out = nullptr;
  // the launch sequence of a repeated call (same tape, sizes and buffers: what a pipelined host does) as a CUDA graph:
  // one cudaGraphLaunch (~10 us) instead of ten launches (~47 us of host time per call -- with 512-signature shards the
  // host's enqueue rate is on the critical path of a short run).  A key seen twice in a row is captured, then replayed.
  cudaGraphExec_t gexec = nullptr;
  std::vector<uint64_t> gkey, seen_key;
};
constexpr int MAX_ARENAS = 64;

struct ChainState {
  std::mutex mu;
  std::map<int, TapeOnDevice> tapes;
  Arena arenas[MAX_ARENAS];
  uint64_t tick = 0;
  bool profiling = false;
  std::vector<cudaEvent_t> events;
  // diagnostic timeline (p2e_set_option("chain_trace", 1)): a timing event after every kernel of every call, on the
  // stream the kernel was launched on -- shows how the batches of a pipelined run interleave (p2e_chain_trace_dump)
  struct TraceEv { int call, kernel; cudaEvent_t ev; };
  std::vector<TraceEv> trace;
  int trace_calls = 0;
  // host-pointer compact path: copy stream + per-chunk events
  cudaStream_t cp = nullptr;
  cudaEvent_t ev_done[8] = {};
};

static ChainState& chain_state() {
  Context& c = ctx();
  static std::mutex create_mu;
  std::lock_guard<std::mutex> lock(create_mu);
  if (!c.chain) c.chain = new ChainState();
  return *(ChainState*)c.chain;
}

static void compact_layout(const Tape& t, uint32_t* map, uint32_t* narrow_slots, uint32_t* wide_slots) {
  uint32_t nn = 0, nw = 0;
  for (size_t g = 0; g < t.gen_kind.size(); g++) {
    const uint32_t kind = t.gen_kind[g], off = t.gen_off[g];
    const uint32_t end = (g + 1 < t.gen_kind.size()) ? t.gen_off[g + 1] : t.n_slots;
    for (uint32_t s = off; s < end; s++) {
      // MUL tuple: r[9], q[9] narrow, check_sum[17] wide; CHECKSUM tuple: b[16] as 16 low words + one
      // word of two-bit tops (b_j + 2^33 < 2^34, mul_nonnative.rs:527); everything else narrow
      const bool is_wide = (kind == G_MUL && s - off >= 18);
      if (kind == G_CHECKSUM) {
        if (map) map[s] = (1u << 30) | (s == off ? (1u << 29) : 0u) | nn;
        nn++;
        if (s + 1 == end) nn++;          // the tuple's row of tops
      } else {
        if (map) map[s] = is_wide ? ((1u << 31) | nw) : nn;
        if (is_wide) nw++; else nn++;
      }
    }
  }
  if (narrow_slots) *narrow_slots = nn;
  if (wide_slots) *wide_slots = nw;
}

static int upload_tape(const Tape& t, TapeOnDevice& o) {
  if (o.ready) return 0;
  // compact layout + per-curve-op rows; every Ref must name plain value slots (they are read back as u32 rows)
  std::vector<u32> cmap(t.n_slots + 1, 0), nrow(t.ptops.size() + 1, 0), wrow(t.ptops.size() + 1, 0);
  compact_layout(t, cmap.data(), &o.n_narrow, &o.n_wide);
  auto ref_ok = [&](const Ref& r) {
    if (r.space == SP_OUT)
      for (uint32_t l = 0; l < r.nl; l++) if (r.off + l >= t.n_slots || (cmap[r.off + l] >> 29)) return false;
    if (r.bspace == SP_OUT && (r.boff >= t.n_slots || (cmap[r.boff] >> 29))) return false;
    return true;
  };
  bool ok = ref_ok(t.fixed_scalar) && ref_ok(t.msm_k1) && ref_ok(t.msm_k2);
  for (auto& p : t.prep) ok = ok && ref_ok(p.a) && ref_ok(p.b);
  for (auto& r : t.extra_refs) ok = ok && ref_ok(r);
  for (auto& c : t.checks) ok = ok && ref_ok(c.a) && ref_ok(c.b);
  for (auto& p : t.in_points) ok = ok && ref_ok(p.x) && ref_ok(p.y);
  if (!ok) return set_error_msg("tape: a reference names a check_sum / carry slot (not representable in the compact format)");
  for (size_t k = 0; k < t.ptops.size(); k++) {
    const PtOp& op = t.ptops[k];
    nrow[k] = cmap[op.slot];
    const uint32_t first_cs = op.slot + (op.type == PT_DBL ? 28u : 38u) + 18u;
    wrow[k] = first_cs < t.n_slots ? (cmap[first_cs] & 0x7fffffffu) : 0u;
  }
  std::vector<u32> cp8(t.const_points.size() / 18 * 16 + 16, 0);
  for (size_t p = 0; p < t.const_points.size() / 18; p++) {
    u32 x[NL], y[NL], x8[NW], y8[NW];
    for (int l = 0; l < NL; l++) { x[l] = t.const_points[p * 18 + l]; y[l] = t.const_points[p * 18 + NL + l]; }
    f8_from_l29(x, x8);
    f8_from_l29(y, y8);
    for (int l = 0; l < NW; l++) { cp8[p * 16 + l] = x8[l]; cp8[p * 16 + NW + l] = y8[l]; }
  }
  size_t off = 0;
  auto reserve = [&](size_t bytes) { size_t at = (off + 15) / 16 * 16; off = at + bytes; return at; };
  size_t o_prep = reserve(t.prep.size() * sizeof(PrepOp));
  size_t o_pt = reserve(t.ptops.size() * sizeof(PtOp));
  size_t o_chk = reserve(t.checks.size() * sizeof(Check));
  size_t o_in = reserve(t.in_points.size() * sizeof(InputPoint));
  size_t o_xr = reserve((t.extra_refs.size() + 1) * sizeof(Ref));
  size_t o_cl = reserve((t.const_limbs.size() + 1) * sizeof(u32));
  size_t o_cp = reserve((t.const_points.size() + 1) * sizeof(u32));
  size_t o_c8 = reserve(cp8.size() * sizeof(u32));
  size_t o_cm = reserve(cmap.size() * sizeof(u32));
  size_t o_nr = reserve(nrow.size() * sizeof(u32));
  size_t o_wr = reserve(wrow.size() * sizeof(u32));
  size_t o_mu = reserve((t.muls.size() + 1) * sizeof(MulRec));
  std::vector<unsigned char> host(off + 16, 0);
  memcpy(&host[o_prep], t.prep.data(), t.prep.size() * sizeof(PrepOp));
  memcpy(&host[o_pt], t.ptops.data(), t.ptops.size() * sizeof(PtOp));
  memcpy(&host[o_chk], t.checks.data(), t.checks.size() * sizeof(Check));
  memcpy(&host[o_in], t.in_points.data(), t.in_points.size() * sizeof(InputPoint));
  memcpy(&host[o_xr], t.extra_refs.data(), t.extra_refs.size() * sizeof(Ref));
  memcpy(&host[o_cl], t.const_limbs.data(), t.const_limbs.size() * sizeof(u32));
  memcpy(&host[o_cp], t.const_points.data(), t.const_points.size() * sizeof(u32));
  memcpy(&host[o_c8], cp8.data(), cp8.size() * sizeof(u32));
  memcpy(&host[o_cm], cmap.data(), cmap.size() * sizeof(u32));
  memcpy(&host[o_nr], nrow.data(), nrow.size() * sizeof(u32));
  memcpy(&host[o_wr], wrow.data(), wrow.size() * sizeof(u32));
  memcpy(&host[o_mu], t.muls.data(), t.muls.size() * sizeof(MulRec));
  void* blob = nullptr;
  P2E_CUDA(cudaMalloc(&blob, host.size()));
  P2E_CUDA(cudaMemcpy(blob, host.data(), host.size(), cudaMemcpyHostToDevice));
  unsigned char* b = (unsigned char*)blob;
  o.d.prep = (const PrepOp*)(b + o_prep);
  o.d.ptops = (const PtOp*)(b + o_pt);
  o.d.checks = (const Check*)(b + o_chk);
  o.d.in_points = (const InputPoint*)(b + o_in);
  o.d.extra_refs = (const Ref*)(b + o_xr);
  o.d.const_limbs = (const u32*)(b + o_cl);
  o.d.const_points = (const u32*)(b + o_cp);
  o.d.const_points8 = (const u32*)(b + o_c8);
  o.d.cmap = (const u32*)(b + o_cm);
  o.d.pt_nrow = (const u32*)(b + o_nr);
  o.d.pt_wrow = (const u32*)(b + o_wr);
  o.d.muls = (const MulRec*)(b + o_mu);
  o.d.n_muls = (int)t.muls.size();
  o.d.n_prep = (int)t.prep.size();
  o.d.n_ptops = (int)t.ptops.size();
  o.d.n_checks = (int)t.checks.size();
  o.d.n_in_points = (int)t.in_points.size();
  for (int i = 0; i < 16; i++) o.d.pre_pid[i] = t.pre_pid[i];
  o.d.fixed_scalar = t.fixed_scalar;
  o.d.msm_k1 = t.msm_k1;
  o.d.msm_k2 = t.msm_k2;
  o.d.fixed_round = t.fixed_round;
  o.d.msm_round = t.msm_round;
  o.d.fixed_table_base = t.fixed_table_base;
  for (int c = 0; c < N_CHAINS; c++) { o.d.chain_begin[c] = t.chain_begin[c]; o.d.chain_end[c] = t.chain_end[c]; }
  o.blob = blob;
  o.n_slots = t.n_slots;
  o.ready = true;
  return 0;
}

static void arena_free(Arena& a) {
  if (a.gexec) cudaGraphExecDestroy(a.gexec);
  if (a.ptr) cudaFree(a.ptr);
  if (a.hi) cudaStreamDestroy(a.hi);
  if (a.lo) cudaStreamDestroy(a.lo);
  if (a.ev_in) cudaEventDestroy(a.ev_in);
  if (a.ev_mid) cudaEventDestroy(a.ev_mid);
  if (a.ev_out) cudaEventDestroy(a.ev_out);
  a = Arena();
}

// The arena of a caller stream; a stream never seen before takes a free entry or evicts the least recently used
// arena whose work has completed (a long-running host that creates a stream per job must not run out).
static int get_arena(ChainState& S, cudaStream_t st, size_t bytes, Arena** out) {
  Arena* a = nullptr;
  for (auto& x : S.arenas) if (x.used && x.stream == st) { a = &x; break; }
  if (!a) {
    for (auto& x : S.arenas) if (!x.used) { a = &x; break; }
    if (!a) {
      Arena* lru = nullptr;
      for (auto& x : S.arenas)
        if (cudaEventQuery(x.ev_out) == cudaSuccess && (!lru || x.last_use < lru->last_use)) lru = &x;
      if (!lru) return set_error_msg("chain: 64 streams have work in flight (one scratch arena each); synchronise one of them or call p2e_stream_release");
      arena_free(*lru);
      a = lru;
    }
    int lo_p = 0, hi_p = 0;
    P2E_CUDA(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));     // numerically lower = higher priority
    P2E_CUDA(cudaStreamCreateWithPriority(&a->hi, cudaStreamNonBlocking, hi_p));
    P2E_CUDA(cudaStreamCreateWithPriority(&a->lo, cudaStreamNonBlocking, lo_p));
    P2E_CUDA(cudaEventCreateWithFlags(&a->ev_in, cudaEventDisableTiming));
    P2E_CUDA(cudaEventCreateWithFlags(&a->ev_mid, cudaEventDisableTiming));
    P2E_CUDA(cudaEventCreateWithFlags(&a->ev_out, cudaEventDisableTiming));
    a->used = true;
    a->stream = st;
  }
  if (a->bytes < bytes) {
    if (a->ptr) {
      P2E_CUDA(cudaEventSynchronize(a->ev_out));   // the previous call on this stream may still read the old buffer
      P2E_CUDA(cudaFree(a->ptr));
    }
    a->ptr = nullptr;
    a->bytes = 0;
    P2E_CUDA(cudaMalloc(&a->ptr, bytes));
    a->bytes = bytes;
  }
  a->last_use = ++S.tick;
  *out = a;
  return 0;
}

void chain_release_device(Context& c) {
  if (!c.chain) return;
  ChainState* S = (ChainState*)c.chain;
  for (auto& kv : S->tapes) {
    if (kv.second.ready && kv.second.blob) cudaFree(kv.second.blob);
    if (kv.second.rc_calls) cudaFree(kv.second.rc_calls);
    for (auto& o : kv.second.rc_off) cudaFree(o.second);
  }
  for (auto& a : S->arenas) arena_free(a);
  for (auto e : S->events) cudaEventDestroy(e);
  for (auto& t : S->trace) cudaEventDestroy(t.ev);
  if (S->cp) cudaStreamDestroy(S->cp);
  for (auto& e : S->ev_done) if (e) cudaEventDestroy(e);
  delete S;
  c.chain = nullptr;
}

static void prof_mark(ChainState& S, cudaStream_t st) {
  if (!S.profiling) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, st);
  S.events.push_back(e);
}

static void trace_mark(ChainState& S, bool on, int call, int kernel, cudaStream_t st) {
  if (!on || S.trace.size() > 100000) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  S.trace.push_back({call, kernel, e});
}

static size_t chunk_capacity() {
  long v = option(OPT_CHAIN_CHUNK);
  return v > 0 ? (size_t)v : 65536;   // 75 KB of scratch per signature: 4.9 GB per arena at full chunks (180 GB HBM); larger chunks
                                      // give the one-warp-per-SM kernels more resident warps
}

struct OutSpec {
  int fmt;
  uint64_t* out;      // plain
  uint32_t* nar;      // compact
  uint64_t* wide;
  size_t os;
};

static int run_tape(int which, const Tape& t, size_t n, const uint64_t* const* ins, int n_in, size_t is, const OutSpec& os,
                    uint32_t* err, void* stream) {
  int rc = ensure_ready();
  if (rc) return rc;
  if (n == 0) return 0;
  if (is < n || os.os < n) return set_error_msg("stride < n");
  if (n_in > MAX_TAPE_INPUTS) return set_error_msg("tape has more than 16 inputs");
  ChainState& S = chain_state();
  std::lock_guard<std::mutex> lock(S.mu);
  TapeOnDevice& o = S.tapes[which];
  if ((rc = upload_tape(t, o))) return rc;
  cudaStream_t st = pick_stream(stream);
  const size_t cap = n < chunk_capacity() ? n : chunk_capacity();
  const size_t np = (size_t)o.d.n_ptops + o.d.n_in_points, no = (size_t)o.d.n_ptops;
  // scratch layout (u32 words)
  size_t words = (np * (4 * NW + 2 * NL) + no * NW) * cap;      // JX JY JZ ZC + AX AY + ZINV
  const unsigned G = INV_GROUPS;
  size_t bytes = words * 4 + no * cap * (4 + 4 + 1) + ((size_t)G * 2 * NW + 1) * cap * 4 + 64;
  Arena* arena = nullptr;
  if ((rc = get_arena(S, st, bytes, &arena))) return rc;
  void* scratch = arena->ptr;
  // profiling keeps everything on the caller's stream so the events bracket each kernel
  // internal high / low priority streams (option "chain_split", default off): they pay only for a host that keeps fewer
  // caller streams in flight than the device has hardware queues (16 streams x 4096 signatures, two-lane replay: 0.628
  // against 0.652 ms per call); with 32 caller streams they alias on the 32 queues (0.600 against 0.618 ms, one-lane
  // replay) and below 2048 signatures per call their event hops cost more than they hide (profiles/r02_notes.md)
  const bool split = !S.profiling && option(OPT_CHAIN_SPLIT) != 0;
  cudaStream_t s_hi = split ? arena->hi : st, s_lo = split ? arena->lo : st;
  if (split) {
    P2E_CUDA(cudaEventRecord(arena->ev_in, st));
    P2E_CUDA(cudaStreamWaitEvent(s_hi, arena->ev_in, 0));
    P2E_CUDA(cudaStreamWaitEvent(s_lo, arena->ev_in, 0));    // orders this call after the previous one's K3-K5 reads
  }
  ChainArgs A;
  A.t = o.d;
  A.is = is;
  A.fmt = os.fmt;
  A.os = os.os;
  A.cap = cap;
  u32* w = (u32*)scratch;
  A.JX = w; w += np * NW * cap;
  A.JY = w; w += np * NW * cap;
  A.JZ = w; w += np * NW * cap;
  A.ZC = w; w += np * NW * cap;
  A.AX = w; w += np * NL * cap;
  A.AY = w; w += np * NL * cap;
  A.ZINV = w; w += no * NW * cap;
  A.SRC1 = (int32_t*)w; w += no * cap;
  A.SRC2 = (int32_t*)w; w += no * cap;
  A.FLG = (uint8_t*)w;
  w += (no * cap + 3) / 4;
  A.GP = w; w += (size_t)G * NW * cap;
  A.GI = w; w += (size_t)G * NW * cap;
  A.BAD = w;
  const bool tr = option(OPT_CHAIN_TRACE) != 0 && !S.profiling;
  const int call_id = tr ? S.trace_calls++ : 0;
  trace_mark(S, tr, call_id, 0, s_hi);
  // graph replay of a repeated call (see Arena::gexec); never with the diagnostics or the internal priority streams
  const bool graphable = option(OPT_CHAIN_GRAPH) != 0 && !split && !S.profiling && !tr;
  std::vector<uint64_t> key;
  if (graphable) {
    key = {(uint64_t)which, (uint64_t)(uintptr_t)o.blob, (uint64_t)n, (uint64_t)is, (uint64_t)os.fmt, (uint64_t)os.os,
           (uint64_t)(uintptr_t)os.out, (uint64_t)(uintptr_t)os.nar, (uint64_t)(uintptr_t)os.wide, (uint64_t)(uintptr_t)err,
           (uint64_t)(uintptr_t)scratch, (uint64_t)cap, (uint64_t)option(OPT_CHAIN_REPLAY_LANES), (uint64_t)option(OPT_CHAIN_PHASES),
           (uint64_t)n_in};
    for (int k = 0; k < n_in; k++) key.push_back((uint64_t)(uintptr_t)ins[k]);
    if (arena->gexec && key == arena->gkey) {
      P2E_CUDA(cudaGraphLaunch(arena->gexec, st));
      P2E_CUDA(cudaEventRecord(arena->ev_out, st));
      return 0;
    }
  }
  auto enqueue = [&]() -> int {
  for (size_t c0 = 0; c0 < n; c0 += cap) {
    size_t cn = (n - c0 < cap) ? n - c0 : cap;
    A.n = cn;
    for (int k = 0; k < MAX_TAPE_INPUTS; k++) A.in[k] = (k < n_in) ? ins[k] + c0 : nullptr;
    A.out = os.out ? os.out + c0 : nullptr;
    A.nar = os.nar ? os.nar + c0 : nullptr;
    A.wide = os.wide ? os.wide + c0 : nullptr;
    A.err = err + c0;
    unsigned g32 = (unsigned)((cn + 31) / 32);
    const bool prof = S.profiling && c0 == 0;
    if (prof) {
      for (auto e : S.events) cudaEventDestroy(e);
      S.events.clear();
      prof_mark(S, st);
    }
    const int n_rounds = (int)t.n_rounds;
    // diagnostic (scripts/chain_phases.py): bit mask of the kernels to launch -- 1 prep, 2 replay + join, 4 invert, 8 affine,
    // 16 expand, 32 checks.  Anything but 63 skips work: every row of the call is then flagged ERR_INTERNAL.
    const unsigned pm = (unsigned)option(OPT_CHAIN_PHASES) & 63u;
    for (int round = 0; round < n_rounds; round++) {
      // Round r: generators of round r (K0), then the curve ops of round r (K1-K4).  A later round
      // consumes coordinates written by this round's expansion, so it waits for the low-priority stream.
      A.round = round;
      if (split && round > 0) {
        P2E_CUDA(cudaEventRecord(arena->ev_out, s_lo));
        P2E_CUDA(cudaStreamWaitEvent(s_hi, arena->ev_out, 0));
      }
      if (pm & 1) k_prep<<<g32, 32, 0, s_hi>>>(A);
      if (prof) prof_mark(S, st);
      trace_mark(S, tr, call_id, 1, s_hi);
      if (no) {
        const long lanes_opt = option(OPT_CHAIN_REPLAY_LANES);
        const bool duo = lanes_opt == 2 || (lanes_opt == 0 && n < 4096);
        const unsigned g_duo = (unsigned)(((duo ? 2 : 1) * cn + 63) / 64);
        // fixed-base and msm chains side by side
        if (pm & 2) {
          if (duo) k_chain_replay<2><<<dim3(g_duo, 2), 64, 0, s_hi>>>(A, 0);
          else k_chain_replay<1><<<dim3(g_duo, 2), 64, 0, s_hi>>>(A, 0);
        }
        if (prof) prof_mark(S, st);
        trace_mark(S, tr, call_id, 2, s_hi);
        if ((pm & 2) && o.d.chain_end[2] > o.d.chain_begin[2]) {                                             // join
          if (duo) k_chain_replay<2><<<dim3(g_duo, 1), 64, 0, s_hi>>>(A, 2);
          else k_chain_replay<1><<<dim3(g_duo, 1), 64, 0, s_hi>>>(A, 2);
        }
        if (prof) prof_mark(S, st);
        trace_mark(S, tr, call_id, 3, s_hi);
        if (pm & 4) {
          k_chain_invert_up<<<dim3(g32, G), 32, 0, s_hi>>>(A);
          k_chain_invert_mid<<<g32, 32, 0, s_hi>>>(A, (int)G);
          k_chain_invert_down<<<dim3(g32, G), 32, 0, s_hi>>>(A);
        }
        if (prof) prof_mark(S, st);
        trace_mark(S, tr, call_id, 4, s_hi);
        if (split) {
          P2E_CUDA(cudaEventRecord(arena->ev_mid, s_hi));
          P2E_CUDA(cudaStreamWaitEvent(s_lo, arena->ev_mid, 0));
        }
        const unsigned tpb = 128, gw = (unsigned)((cn + tpb - 1) / tpb);
        trace_mark(S, tr, call_id, 5, s_lo);        // s_lo has waited for ev_mid here: start of the full-GPU phase
        if (pm & 8) k_chain_affine<<<dim3(gw, (unsigned)no), tpb, 0, s_lo>>>(A);
        if (prof) prof_mark(S, st);
        trace_mark(S, tr, call_id, 6, s_lo);
        if (!(pm & 16)) {
        } else if (A.fmt == FMT_PLAIN) k_chain_expand<FMT_PLAIN><<<dim3(gw, (unsigned)no), tpb, 0, s_lo>>>(A);
        else k_chain_expand<FMT_COMPACT><<<dim3(gw, (unsigned)no), tpb, 0, s_lo>>>(A);
        if (prof) prof_mark(S, st);
        trace_mark(S, tr, call_id, 7, s_lo);
      } else if (split) {
        P2E_CUDA(cudaEventRecord(arena->ev_mid, s_hi));
        P2E_CUDA(cudaStreamWaitEvent(s_lo, arena->ev_mid, 0));
      }
    }
    if (pm & 32) k_chain_checks<<<g32, 32, 0, s_lo>>>(A);
    if (pm != 63u) k_chain_poison<<<g32, 32, 0, s_lo>>>(cn, A.err);
    if (prof) prof_mark(S, st);
    if (split && c0 + cap < n) {
      // next chunk reuses the scratch: its K0-K2 must not start before this chunk's K3-K5 are done
      P2E_CUDA(cudaEventRecord(arena->ev_out, s_lo));
      P2E_CUDA(cudaStreamWaitEvent(s_hi, arena->ev_out, 0));
    }
    P2E_CUDA(cudaGetLastError());
  }
  return 0;
  };
  if (graphable && key == arena->seen_key) {
    // second call in a row with this key: capture the sequence, instantiate, replay from now on
    if (arena->gexec) { cudaGraphExecDestroy(arena->gexec); arena->gexec = nullptr; arena->gkey.clear(); }
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      rc = enqueue();
      const cudaError_t ce = cudaStreamEndCapture(st, &g);
      if (rc == 0 && ce == cudaSuccess && g && cudaGraphInstantiate(&arena->gexec, g, 0) == cudaSuccess) {
        cudaGraphDestroy(g);
        arena->gkey = key;
        P2E_CUDA(cudaGraphLaunch(arena->gexec, st));
        P2E_CUDA(cudaEventRecord(arena->ev_out, st));
        return 0;
      }
      if (g) cudaGraphDestroy(g);
      arena->gexec = nullptr;
      cudaGetLastError();
      if (rc) return rc;
    } else {
      cudaGetLastError();         // the caller's stream cannot be captured (e.g. it is being captured already): plain launches
    }
  }
  if (graphable) arena->seen_key = key;
  if ((rc = enqueue())) return rc;
  if (split) {
    P2E_CUDA(cudaEventRecord(arena->ev_out, s_lo));
    P2E_CUDA(cudaStreamWaitEvent(st, arena->ev_out, 0));
  } else {
    P2E_CUDA(cudaEventRecord(arena->ev_out, st));
  }
  return 0;
}

// the range_check_ux_circuit calls of a tape in gadget order (see k_tape_range_check_rows)
static void range_check_calls(const Tape& t, std::vector<RcCall>& calls, std::vector<int>* gens) {
  size_t mi = 0;
  auto push = [&](uint8_t kind, const Ref& r, u32 slot, u32 tuple, int count, int bits, int first, int g) {
    RcCall c;
    c.r = r; c.slot = slot; c.tuple = tuple; c.kind = kind; c.count = (uint8_t)count; c.bits = (uint8_t)bits; c.first = (uint8_t)first;
    calls.push_back(c);
    if (gens) gens->push_back(g);
  };
  Ref none;
  memset(&none, 0, sizeof(none));
  for (size_t g = 0; g < t.gen_kind.size(); g++) {
    const u32 s = t.gen_off[g];
    switch (t.gen_kind[g]) {
      case G_ADDMANY:                                             // nonnative.rs:330-331
        push(RC_SLOTS, none, s, 0, NL, LB, 0, (int)g);
        push(RC_SLOTS, none, s + 9, 0, 1, LB, 0, (int)g);
        break;
      case G_SUB:                                                 // nonnative.rs:373
        push(RC_SLOTS, none, s, 0, NL, LB, 0, (int)g);
        break;
      case G_MUL: {                                               // nonnative.rs:454-460
        const MulRec& m = t.muls[mi++];
        push(RC_REF, m.x, 0, 0, m.x.nl, LB, 0, (int)g);
        push(RC_REF, m.y, 0, 0, m.y.nl, LB, 0, (int)g);
        push(RC_SLOTS, none, s, 0, NL, LB, 0, (int)g);            // r
        push(RC_SLOTS, none, s + 9, 0, NL, LB, 0, (int)g);        // q
        push(RC_SLOTS, none, s, 0, NL, LB, 0, (int)g);            // r again
        push(RC_CARRIES, none, s + 35, s, 8, 34, 0, (int)g);      // b[0..8]
        push(RC_CARRIES, none, s + 43, s, 8, 34, 8, (int)g);      // b[8..16]
        break;
      }
      default: break;
    }
  }
}

static size_t tape_layout(const Tape& t, uint32_t* kinds, uint32_t* offs) {
  if (kinds && offs) {
    for (size_t i = 0; i < t.gen_kind.size(); i++) {
      kinds[i] = t.gen_kind[i];
      offs[i] = t.gen_off[i];
    }
  }
  return t.gen_kind.size();
}

}  // namespace p2e

using namespace p2e;

extern "C" {

int p2e_set_profiling(int on) {
  int rc = ensure_ready();
  if (rc) return rc;
  chain_state().profiling = on != 0;
  return 0;
}

int p2e_last_kernel_times(float* ms, int cap) {
  int rc = ensure_ready();
  if (rc) return rc;
  ChainState& S = chain_state();
  std::lock_guard<std::mutex> lock(S.mu);
  if (S.events.size() < 2) return 0;
  P2E_CUDA(cudaEventSynchronize(S.events.back()));
  int k = 0;
  for (size_t j = 0; j + 1 < S.events.size() && k < cap; j++, k++) {
    float t = 0;
    P2E_CUDA(cudaEventElapsedTime(&t, S.events[j], S.events[j + 1]));
    ms[k] = t;
  }
  return k;
}

// diagnostic: events of the traced calls as (call, kernel id, ms since the first event); kernel ids: 0 call start,
// 1 k_prep done, 2 replay done, 3 join done, 4 invert done, 5 full-GPU phase may start, 6 affine done, 7 expand done.
// Synchronises the device; clears the trace.  Returns the number of entries written.
int p2e_chain_trace_dump(int* calls, int* kernels, float* ms, int cap) {
  int rc = ensure_ready();
  if (rc) return rc;
  ChainState& S = chain_state();
  std::lock_guard<std::mutex> lock(S.mu);
  P2E_CUDA(cudaDeviceSynchronize());
  int k = 0;
  for (size_t j = 0; j < S.trace.size(); j++) {
    if (k < cap) {
      float t = 0;
      cudaEventElapsedTime(&t, S.trace[0].ev, S.trace[j].ev);
      calls[k] = S.trace[j].call;
      kernels[k] = S.trace[j].kernel;
      ms[k] = t;
      k++;
    }
  }
  for (auto& t : S.trace) cudaEventDestroy(t.ev);
  S.trace.clear();
  S.trace_calls = 0;
  return k;
}

int p2e_stream_release(void* stream) {
  int rc = ensure_ready();
  if (rc) return rc;
  ChainState& S = chain_state();
  std::lock_guard<std::mutex> lock(S.mu);
  cudaStream_t st = pick_stream(stream);
  for (auto& a : S.arenas)
    if (a.used && a.stream == st) {
      P2E_CUDA(cudaEventSynchronize(a.ev_out));
      arena_free(a);
    }
  return 0;
}

static int bad_tape(int tape) {
  if (!tape_exists(tape)) return set_error_msg("unknown tape id");
  return 0;
}

size_t p2e_tape_slots(int tape) { return bad_tape(tape) ? 0 : get_tape(tape).n_slots; }
int p2e_tape_inputs(int tape) { return bad_tape(tape) ? -1 : (int)get_tape(tape).n_inputs; }
size_t p2e_tape_layout(int tape, uint32_t* kinds, uint32_t* offsets) {
  return bad_tape(tape) ? 0 : tape_layout(get_tape(tape), kinds, offsets);
}

// ---- the range-check rows of mul_nonnative (the plonky2_ux rows under the in-tree generators) ----
int p2e_tape_num_mul(int tape) { return bad_tape(tape) ? -1 : (int)get_tape(tape).muls.size(); }

int p2e_tape_mul_info(int tape, int j, int* x_limbs, int* y_limbs, uint32_t* out_slot, int* modulus) {
  if (bad_tape(tape)) return -1;
  const Tape& t = get_tape(tape);
  if (j < 0 || j >= (int)t.muls.size()) return set_error_msg("tape_mul_info: instance out of range");
  if (x_limbs) *x_limbs = t.muls[j].x.nl;
  if (y_limbs) *y_limbs = t.muls[j].y.nl;
  if (out_slot) *out_slot = t.muls[j].out;
  if (modulus) *modulus = (int)t.muls[j].mod;
  return 0;
}

int p2e_tape_num_range_checks(int tape) {
  if (bad_tape(tape)) return -1;
  std::vector<RcCall> calls;
  range_check_calls(get_tape(tape), calls, nullptr);
  return (int)calls.size();
}

int p2e_tape_range_check_info(int tape, int first, int count, int32_t* generator, int32_t* inputs, int32_t* bits) {
  if (bad_tape(tape)) return -1;
  std::vector<RcCall> calls;
  std::vector<int> gens;
  range_check_calls(get_tape(tape), calls, &gens);
  if (first < 0 || count < 0 || (size_t)first + count > calls.size()) return set_error_msg("tape_range_check_info: call range out of bounds");
  for (int k = 0; k < count; k++) {
    if (generator) generator[k] = gens[first + k];
    if (inputs) inputs[k] = calls[first + k].count;
    if (bits) bits[k] = calls[first + k].bits;
  }
  return 0;
}

int p2e_tape_range_check_layout(int tape, int limb_bits, uint64_t* wire_offsets) {
  if (bad_tape(tape)) return -1;
  if (limb_bits < 1 || limb_bits > 8) return set_error_msg("tape_range_check: limb_bits must be 1 .. 8");
  std::vector<RcCall> calls;
  range_check_calls(get_tape(tape), calls, nullptr);
  u64 off = 0;
  for (size_t c = 0; c < calls.size(); c++) {
    if (wire_offsets) wire_offsets[c] = off;
    off += (u64)calls[c].count * (1 + (calls[c].bits + limb_bits - 1) / limb_bits);
  }
  if (wire_offsets) wire_offsets[calls.size()] = off;
  return 0;
}

int p2e_tape_range_check_rows_dev(int tape, size_t n, const uint64_t* const* inputs, size_t is, const uint64_t* witness,
                                  const uint32_t* narrow, size_t ws, int c0, int c1, int limb_bits, uint64_t* rows, size_t rs,
                                  uint32_t* err, void* stream) {
  if (bad_tape(tape)) return -1;
  int rc = ensure_ready();
  if (rc) return rc;
  const Tape& t = get_tape(tape);
  if (limb_bits < 1 || limb_bits > 8) return set_error_msg("tape_range_check: limb_bits must be 1 .. 8");
  if ((witness == nullptr) == (narrow == nullptr)) return set_error_msg("tape_range_check_rows: give the plain witness OR the compact narrow matrix");
  if (n && (!rows || !err)) return set_error_msg("tape_range_check_rows: NULL output");
  if (is < n || ws < n || rs < n) return set_error_msg("stride < n");
  if (t.n_inputs && !inputs) return set_error_msg("tape_range_check_rows: NULL inputs");
  ChainState& S = chain_state();
  TapeOnDevice* o;
  u64* d_off = nullptr;
  u64 base = 0;
  {
    std::lock_guard<std::mutex> lock(S.mu);
    o = &S.tapes[tape];
    if ((rc = upload_tape(t, *o))) return rc;
    if (!o->rc_calls) {
      std::vector<RcCall> calls;
      range_check_calls(t, calls, nullptr);
      o->n_rc = (int)calls.size();
      P2E_CUDA(cudaMalloc((void**)&o->rc_calls, (calls.size() + 1) * sizeof(RcCall)));
      P2E_CUDA(cudaMemcpy(o->rc_calls, calls.data(), calls.size() * sizeof(RcCall), cudaMemcpyHostToDevice));
    }
    if (c0 < 0 || c1 < c0 || c1 > o->n_rc) return set_error_msg("tape_range_check_rows: call range out of bounds");
    if (c1 - c0 > 65535) return set_error_msg("tape_range_check_rows: at most 65535 calls per launch");
    std::vector<u64> off((size_t)o->n_rc + 1);
    if ((rc = p2e_tape_range_check_layout(tape, limb_bits, off.data()))) return rc;
    base = off[c0];
    auto it = o->rc_off.find(limb_bits);
    if (it == o->rc_off.end()) {
      u64* p = nullptr;
      P2E_CUDA(cudaMalloc((void**)&p, off.size() * 8));
      P2E_CUDA(cudaMemcpy(p, off.data(), off.size() * 8, cudaMemcpyHostToDevice));
      it = o->rc_off.emplace(limb_bits, p).first;
    }
    d_off = it->second;
  }
  if (n == 0 || c0 == c1) return 0;
  ChainArgs A;
  memset(&A, 0, sizeof(A));
  A.t = o->d;
  A.n = n;
  A.is = is;
  for (int k = 0; k < (int)t.n_inputs && k < MAX_TAPE_INPUTS; k++) A.in[k] = inputs[k];
  A.fmt = witness ? FMT_PLAIN : FMT_COMPACT;
  A.out = const_cast<u64*>(witness);
  A.nar = const_cast<u32*>(narrow);
  A.os = ws;
  A.err = err;
  RcArgs M;
  M.calls = o->rc_calls;
  M.wire_off = d_off;
  M.c0 = c0;
  M.lb = limb_bits;
  M.base = base;
  M.rows = rows;
  M.rs = rs;
  cudaStream_t st = pick_stream(stream);
  P2E_CUDA(cudaMemsetAsync(err, 0, n * sizeof(uint32_t), st));
  k_tape_range_check_rows<<<dim3((unsigned)((n + 127) / 128), (unsigned)(c1 - c0)), 128, 0, st>>>(A, M);
  P2E_CUDA(cudaGetLastError());
  return 0;
}

int p2e_tape_witness_dev(int tape, size_t n, const uint64_t* const* inputs, size_t is, uint64_t* out, size_t os,
                         uint32_t* err, void* stream) {
  if (bad_tape(tape)) return -1;
  const Tape& t = get_tape(tape);
  OutSpec o{FMT_PLAIN, out, nullptr, nullptr, os};
  return run_tape(tape, t, n, inputs, (int)t.n_inputs, is, o, err, stream);
}

int p2e_tape_witness_compact_dev(int tape, size_t n, const uint64_t* const* inputs, size_t is, uint32_t* narrow,
                                 uint64_t* wide, size_t os, uint32_t* err, void* stream) {
  if (bad_tape(tape)) return -1;
  if (n && !narrow) return set_error_msg("tape_witness_compact_dev: NULL narrow matrix");
  const Tape& t = get_tape(tape);
  OutSpec o{FMT_COMPACT, nullptr, narrow, wide, os};
  return run_tape(tape, t, n, inputs, (int)t.n_inputs, is, o, err, stream);
}

int p2e_tape_witness(int tape, size_t n, const uint64_t* const* inputs, uint64_t* out, uint32_t* err) {
  if (bad_tape(tape)) return -1;
  int rc = ensure_ready();
  if (rc) return rc;
  if (n == 0) return 0;
  P2E_HOST_LOCK();
  const Tape& t = get_tape(tape);
  const int n_in = (int)t.n_inputs;
  cudaStream_t st = ctx().stream;
  void *din = nullptr, *dout = nullptr, *derr = nullptr;
  if ((rc = get_scratch(0, (size_t)n_in * 9 * n * 8, &din))) return rc;
  if ((rc = get_scratch(1, (size_t)t.n_slots * n * 8, &dout))) return rc;
  if ((rc = get_scratch(2, n * 4, &derr))) return rc;
  const uint64_t* dptr[MAX_TAPE_INPUTS];
  for (int k = 0; k < n_in; k++) {
    dptr[k] = (const uint64_t*)din + (size_t)k * 9 * n;
    P2E_CUDA(cudaMemcpyAsync((uint64_t*)din + (size_t)k * 9 * n, inputs[k], 9 * n * 8, cudaMemcpyHostToDevice, st));
  }
  OutSpec o{FMT_PLAIN, (uint64_t*)dout, nullptr, nullptr, n};
  if ((rc = run_tape(tape, t, n, dptr, n_in, n, o, (uint32_t*)derr, (void*)st))) return rc;
  P2E_CUDA(cudaMemcpyAsync(out, dout, (size_t)t.n_slots * n * 8, cudaMemcpyDeviceToHost, st));
  if (err) P2E_CUDA(cudaMemcpyAsync(err, derr, n * 4, cudaMemcpyDeviceToHost, st));
  P2E_CUDA(cudaStreamSynchronize(st));
  return 0;
}

size_t p2e_tape_compact_layout(int tape, uint32_t* map, uint32_t* narrow_slots, uint32_t* wide_slots) {
  if (bad_tape(tape)) return 0;
  const Tape& t = get_tape(tape);
  compact_layout(t, map, narrow_slots, wide_slots);
  return t.n_slots;
}

// Host-side decoder of the compact formats (no device involved: a format conversion for hosts that want the plain
// [slots][n] uint64 layout back, e.g. to fill plonky2's witness vector): out[slot][i] from narrow (+ wide, or, when
// wide == NULL, the check_sum slots rebuilt from the carries -- the derived form described in include/p2e.h).
int p2e_tape_compact_expand(int tape, size_t n, const uint32_t* narrow, const uint64_t* wide, uint64_t* out) {
  if (bad_tape(tape)) return -1;
  if (n && (!narrow || !out)) return set_error_msg("compact_expand: NULL buffer");
  const Tape& t = get_tape(tape);
  std::vector<uint32_t> map(t.n_slots);
  uint32_t nn = 0, nw = 0;
  compact_layout(t, map.data(), &nn, &nw);
  const i64 off = (i64)1 << 33;
  for (size_t s = 0; s < t.n_slots; s++) {
    const uint32_t m = map[s];
    uint64_t* o = out + s * n;
    if (m >> 31) {
      const size_t row = m & 0x7fffffffu;
      if (wide) {
        for (size_t i = 0; i < n; i++) o[i] = wide[row * n + i];
      } else {
        // check_sum_j = c_j 2^29 - c_{j-1}, c_j = b_j - 2^33 (c_{-1} = c_16 = 0); the carries sit 17 slots further on
        // and are decoded below -- read them from the packed rows directly
        const int j = (int)(row % 17);
        auto carry = [&](size_t slot, size_t i) -> i64 {
          const uint32_t mc = map[slot];
          const size_t r = mc & 0x1fffffffu;
          const int k = (int)(slot - (s - j + 17));                    // index of the carry inside its tuple
          const uint64_t top = (narrow[(r - k + 16) * n + i] >> (2 * k)) & 3u;
          return (i64)((uint64_t)narrow[r * n + i] | (top << 32)) - off;
        };
        for (size_t i = 0; i < n; i++) {
          const i64 cur = j < 16 ? carry(s + 17, i) : 0;
          const i64 prev = j > 0 ? carry(s + 16, i) : 0;
          o[i] = gl_from_i64(cur * ((i64)1 << LB) - prev);
        }
      }
    } else if (m & (1u << 30)) {
      const size_t r = m & 0x1fffffffu;
      // position of this carry inside its tuple: the tuple's first slot carries bit 29
      size_t first = s;
      while (!(map[first] & (1u << 29))) first--;
      const int k = (int)(s - first);
      for (size_t i = 0; i < n; i++) {
        const uint64_t top = (narrow[(r - k + 16) * n + i] >> (2 * k)) & 3u;
        o[i] = (uint64_t)narrow[r * n + i] | (top << 32);
      }
    } else {
      for (size_t i = 0; i < n; i++) o[i] = narrow[(size_t)m * n + i];
    }
  }
  return 0;
}

int p2e_tape_witness_compact(int tape, size_t n, const uint64_t* const* inputs, uint32_t* narrow, uint64_t* wide,
                             uint32_t* err) {
  if (bad_tape(tape)) return -1;
  int rc = ensure_ready();
  if (rc) return rc;
  if (n == 0) return 0;
  P2E_HOST_LOCK();
  const Tape& t = get_tape(tape);
  const int n_in = (int)t.n_inputs;
  uint32_t nn = 0, nw = 0;
  compact_layout(t, nullptr, &nn, &nw);
  cudaStream_t st = ctx().stream;
  // The batch is processed in chunks of signatures so that the D2H copy of chunk c (on a second stream,
  // pitched copies into the caller's [rows][n] matrices) overlaps the kernels of chunk c + 1: the call
  // costs about one chunk of compute plus the PCIe time instead of their sum.  The kernels write the compact
  // rows themselves (no plain witness is materialised on this path).
  ChainState& S = chain_state();
  if (!S.cp) {
    P2E_CUDA(cudaStreamCreateWithFlags(&S.cp, cudaStreamNonBlocking));
    for (auto& e : S.ev_done) P2E_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  cudaStream_t cp = S.cp;
  size_t n_chunks = (size_t)option(OPT_HOST_CHUNKS);
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > 8) n_chunks = 8;
  if (n < 256 * n_chunks) n_chunks = 1;
  const size_t cn = ((n + n_chunks - 1) / n_chunks + 31) / 32 * 32;      // chunk stride: 128-byte aligned rows
  void *din = nullptr, *derr = nullptr, *dpack = nullptr;
  if ((rc = get_scratch(0, (size_t)n_in * 9 * n * 8, &din))) return rc;
  if ((rc = get_scratch(2, n * 4, &derr))) return rc;
  const size_t nbytes = (size_t)nn * cn * 4, wbytes = wide ? (size_t)nw * cn * 8 : 0;          // per chunk
  const size_t region = ((nbytes + 255) / 256) * 256 + ((wbytes + 255) / 256) * 256;
  if ((rc = get_scratch(3, region * n_chunks, &dpack))) return rc;
  const uint64_t* dptr[MAX_TAPE_INPUTS];
  for (int k = 0; k < n_in; k++)
    P2E_CUDA(cudaMemcpyAsync((uint64_t*)din + (size_t)k * 9 * n, inputs[k], 9 * n * 8, cudaMemcpyHostToDevice, st));
  for (size_t c = 0; c < n_chunks; c++) {
    const size_t c0 = c * cn;
    if (c0 >= n) break;
    const size_t cc = (n - c0 < cn) ? n - c0 : cn;
    u32* dn = (u32*)((char*)dpack + region * c);
    u64* dw = (u64*)((char*)dn + ((nbytes + 255) / 256) * 256);
    for (int k = 0; k < n_in; k++) dptr[k] = (const uint64_t*)din + (size_t)k * 9 * n + c0;
    OutSpec o{FMT_COMPACT, nullptr, dn, wide ? dw : nullptr, cn};
    if ((rc = run_tape(tape, t, cc, dptr, n_in, n, o, (uint32_t*)derr + c0, (void*)st))) return rc;
    P2E_CUDA(cudaEventRecord(S.ev_done[c], st));
    P2E_CUDA(cudaStreamWaitEvent(cp, S.ev_done[c], 0));
    P2E_CUDA(cudaMemcpy2DAsync(narrow + c0, n * 4, dn, cn * 4, cc * 4, nn, cudaMemcpyDeviceToHost, cp));
    if (wide) P2E_CUDA(cudaMemcpy2DAsync(wide + c0, n * 8, dw, cn * 8, cc * 8, nw, cudaMemcpyDeviceToHost, cp));
    if (err) P2E_CUDA(cudaMemcpyAsync(err + c0, (uint32_t*)derr + c0, cc * 4, cudaMemcpyDeviceToHost, cp));
  }
  P2E_CUDA(cudaStreamSynchronize(st));
  P2E_CUDA(cudaStreamSynchronize(cp));
  return 0;
}

// test hook (tests/test_gpu_f8.py): PTX bodies of the saturated field on the device; a, b, out: [n][9] u32 (8 words used)
int p2e_debug_f8(int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  int rc = ensure_ready();
  if (rc) return rc;
  if (n == 0) return 0;
  P2E_HOST_LOCK();
  cudaStream_t st = ctx().stream;
  void *da = nullptr, *db = nullptr, *dout = nullptr;
  if ((rc = get_scratch(0, n * 36, &da))) return rc;
  if ((rc = get_scratch(1, n * 36, &db))) return rc;
  if ((rc = get_scratch(2, n * 36, &dout))) return rc;
  P2E_CUDA(cudaMemcpyAsync(da, a, n * 36, cudaMemcpyHostToDevice, st));
  P2E_CUDA(cudaMemcpyAsync(db, b, n * 36, cudaMemcpyHostToDevice, st));
  k_debug_f8<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(op, n, (const u32*)da, (const u32*)db, (u32*)dout);
  P2E_CUDA(cudaGetLastError());
  P2E_CUDA(cudaMemcpyAsync(out, dout, n * 36, cudaMemcpyDeviceToHost, st));
  P2E_CUDA(cudaStreamSynchronize(st));
  return 0;
}

size_t p2e_ecdsa_tape_slots(void) { return p2e_tape_slots(TAPE_ECDSA); }
size_t p2e_ecdsa_tape_layout(uint32_t* kinds, uint32_t* offsets) { return p2e_tape_layout(TAPE_ECDSA, kinds, offsets); }
int p2e_ecdsa_witness_dev(size_t n, const uint64_t* msg, const uint64_t* r, const uint64_t* s, const uint64_t* pkx,
                          const uint64_t* pky, size_t is, uint64_t* out, size_t os, uint32_t* err, void* stream) {
  const uint64_t* ins[5] = {msg, r, s, pkx, pky};
  return p2e_tape_witness_dev(TAPE_ECDSA, n, ins, is, out, os, err, stream);
}
int p2e_ecdsa_witness_compact_dev(size_t n, const uint64_t* msg, const uint64_t* r, const uint64_t* s, const uint64_t* pkx,
                                  const uint64_t* pky, size_t is, uint32_t* narrow, uint64_t* wide, size_t os, uint32_t* err,
                                  void* stream) {
  const uint64_t* ins[5] = {msg, r, s, pkx, pky};
  return p2e_tape_witness_compact_dev(TAPE_ECDSA, n, ins, is, narrow, wide, os, err, stream);
}
int p2e_ecdsa_witness(size_t n, const uint64_t* msg, const uint64_t* r, const uint64_t* s, const uint64_t* pkx,
                      const uint64_t* pky, uint64_t* out, uint32_t* err) {
  const uint64_t* ins[5] = {msg, r, s, pkx, pky};
  return p2e_tape_witness(TAPE_ECDSA, n, ins, out, err);
}

size_t p2e_glv_mul_tape_slots(void) { return p2e_tape_slots(TAPE_GLV_MUL); }
size_t p2e_glv_mul_tape_layout(uint32_t* kinds, uint32_t* offsets) { return p2e_tape_layout(TAPE_GLV_MUL, kinds, offsets); }
int p2e_glv_mul_witness_dev(size_t n, const uint64_t* k, const uint64_t* px, const uint64_t* py, size_t is,
                            uint64_t* out, size_t os, uint32_t* err, void* stream) {
  const uint64_t* ins[3] = {k, px, py};
  return p2e_tape_witness_dev(TAPE_GLV_MUL, n, ins, is, out, os, err, stream);
}
int p2e_glv_mul_witness(size_t n, const uint64_t* k, const uint64_t* px, const uint64_t* py, uint64_t* out,
                        uint32_t* err) {
  const uint64_t* ins[3] = {k, px, py};
  return p2e_tape_witness(TAPE_GLV_MUL, n, ins, out, err);
}

}  // extern "C"
