// p2e_chain.cu -- witness generation for whole gadget chains (verify_secp256k1_message_circuit,
// glv_mul) from the host-built tape (p2e_tape.h).
//
// The reference runs 312 Fermat ladders per signature, one per curve_add / curve_double
// (gadgets/curve.rs:160-223 -> gadgets/nonnative.rs:857-872).  Every one of those inverses is
// unique mod p, so it can be obtained differently without changing a single witness bit:
//   K0 k_prep          per signature: interpreter over the non-curve generators (prep ops)
//   K1 k_chain_replay  per signature: replays all curve ops in Jacobian coordinates (no inversion)
//   K2 k_chain_invert  per signature: ONE field inversion for all Z's (Montgomery's trick)
//   K3 k_chain_affine  per (signature, point): affine x, y and Z^3
//   K4 k_chain_expand  per (signature, curve op): v_inv = Z1^3 Z2^3 / Z3 (add) or Z1^4 / Z3 (double),
//                      then every generator of that curve op with the cores of p2e_field.cuh
//   K5 k_chain_checks  per signature: the gadget's connect_* constraints -> ERR_CONNECT
// Signatures are the fastest-varying index everywhere (coalesced wire-major accesses).
#include <map>
#include <mutex>
#include <vector>

#include "../../include/p2e.h"
#include "p2e_common.h"
#include "p2e_field.cuh"
#include "p2e_tape.h"

#ifndef P2E_DEFAULT_REPLAY
#define P2E_DEFAULT_REPLAY 2
#endif

namespace p2e {

constexpr int32_t PID_INPUT_BASE = 1 << 19;

struct DevTape {
  const PrepOp* prep;
  const PtOp* ptops;
  const Check* checks;
  const InputPoint* in_points;
  const Ref* extra_refs;
  const u32* const_limbs;
  const u32* const_points;
  int n_prep, n_ptops, n_checks, n_in_points;
  int32_t pre_pid[16];
  Ref fixed_scalar, msm_k1, msm_k2;
  u32 fixed_round, msm_round;
  u32 fixed_table_base;
  int32_t chain_begin[N_CHAINS], chain_end[N_CHAINS];
};

struct ChainArgs {
  DevTape t;
  size_t n;                 // signatures in this chunk
  int round;                // dependency round being executed (Tape::n_rounds)
  const u64* in[MAX_TAPE_INPUTS];   // input nonnative values, [9][is] each
  size_t is;
  u64* out;                 // [n_slots][os]
  size_t os;
  u32* err;
  // scratch (stride = cap)
  size_t cap;
  u32 *JX, *JY, *JZ;        // [n_ptops + n_in_points][9][cap]
  u32* ZINV;                // [n_ptops][9][cap]
  u32 *AX, *AY, *ZC;        // [n_ptops + n_in_points][9][cap]
  int32_t *SRC1, *SRC2;     // [n_ptops][cap]
  uint8_t* FLG;             // [n_ptops][cap]
  u32 *GP, *GI;             // [inversion groups][9][cap]: group products / their inverses (K2 tree form)
  u32* BAD;                 // [cap]: bit g set = group g contains a zero Z
};

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 load_wire(const ChainArgs& A, uint8_t space, u32 off, size_t i) {
  if (space == SP_OUT) return A.out[(size_t)off * A.os + i];
  if (space == SP_IN) return A.in[off / 9][(size_t)(off % 9) * A.is + i];
  return A.t.const_limbs[off];
}

__device__ __forceinline__ void load_ref(const ChainArgs& A, const Ref& r, size_t i, u64 (&w)[NL]) {
  u64 bm = 1;
  if (r.bspace != SP_NONE) bm = (load_wire(A, r.bspace, r.boff, i) ^ (u64)r.binv) & 1;
#pragma unroll
  for (int l = 0; l < NL; l++) w[l] = (l < r.nl) ? load_wire(A, r.space, r.off + l, i) * bm : 0;
}

__device__ __forceinline__ void st9(const ChainArgs& A, u32 slot, size_t i, const u32 (&v)[NL]) {
#pragma unroll
  for (int l = 0; l < NL; l++) A.out[(size_t)(slot + l) * A.os + i] = v[l];
}
__device__ __forceinline__ void st1(const ChainArgs& A, u32 slot, size_t i, u64 v) { A.out[(size_t)slot * A.os + i] = v; }

// The generator cores are kept OUT OF LINE: a curve op expands to 3-4 full mul witnesses plus a dozen
// small generators, and fully inlined that is ~130 KB of straight-line code per thread -- the
// expansion kernel was instruction-fetch bound (ncu: stall_no_instruction 7.5 per issue).
// `o` points at out[slot][i]; consecutive tuple elements are `os` apart.

// Operands and results cross the call boundary BY VALUE (structs of 9 limbs): ptxas keeps them in
// registers, whereas pointer arguments forced every operand through local memory (ncu: the
// expansion kernel wrote 3.45 GB to DRAM for 2.63 GB of output).
struct W9 { u32 v[NL]; };
struct I9 { i32 v[NL]; };
struct W9E { W9 r; u32 e; };
__device__ __forceinline__ W9 w9(const u32 (&a)[NL]) { W9 r;
#pragma unroll
  for (int l = 0; l < NL; l++) r.v[l] = a[l];
  return r; }
__device__ __forceinline__ I9 i9(const i32 (&a)[NL]) { I9 r;
#pragma unroll
  for (int l = 0; l < NL; l++) r.v[l] = a[l];
  return r; }

// MulNonnative row tuple (r, q, check_sum) followed by the CheckSum tuple (b): 51 slots
template <class MOD>
__device__ __noinline__ W9E nl_mul(u64* __restrict__ o, size_t os, W9 x, W9 y) {
  u32 r[NL], q[NL];
  u64 cs[2 * NL - 1], b[2 * NL - 2];
  W9E ret;
  ret.e = mul_witness_core<MOD>(x.v, y.v, r, q, cs, b);
#pragma unroll
  for (int l = 0; l < NL; l++) {
    o[(size_t)l * os] = r[l];
    o[(size_t)(9 + l) * os] = q[l];
    ret.r.v[l] = r[l];
  }
#pragma unroll
  for (int l = 0; l < 2 * NL - 1; l++) o[(size_t)(18 + l) * os] = cs[l];
#pragma unroll
  for (int l = 0; l < 2 * NL - 2; l++) o[(size_t)(35 + l) * os] = b[l];
  return ret;
}
// ADD / SUB tuple: value[9], overflow
template <class MOD, bool IS_SUB>
__device__ __noinline__ W9 nl_addsub(u64* __restrict__ o, size_t os, W9 a, W9 b) {
  W9 r;
  u32 ov = IS_SUB ? sub_witness_core<MOD>(a.v, b.v, r.v) : add_witness_core<MOD>(a.v, b.v, r.v);
#pragma unroll
  for (int l = 0; l < NL; l++) o[(size_t)l * os] = r.v[l];
  o[(size_t)9 * os] = ov;
  return r;
}
// NonNativeInverseGenerator tuple (div, inv) when the inverse is already known
template <class MOD>
__device__ __noinline__ u32 nl_inv_known(u64* __restrict__ o, size_t os, W9 xc, W9 inv) {
  u32 div[NL];
  u32 e = inv_div_core<MOD>(xc.v, inv.v, div);
#pragma unroll
  for (int l = 0; l < NL; l++) { o[(size_t)l * os] = div[l]; o[(size_t)(9 + l) * os] = inv.v[l]; }
  return e;
}

template <class MOD>
__device__ __forceinline__ u32 do_mul(const ChainArgs& A, u32 slot, size_t i, const u32 (&x)[NL], const u32 (&y)[NL],
                                      u32 (&r)[NL]) {
  W9E t = nl_mul<MOD>(A.out + (size_t)slot * A.os + i, A.os, w9(x), w9(y));
#pragma unroll
  for (int l = 0; l < NL; l++) r[l] = t.r.v[l];
  return t.e;
}
template <class MOD>
__device__ __forceinline__ void do_add(const ChainArgs& A, u32 slot, size_t i, const u32 (&a)[NL], const u32 (&b)[NL],
                                       u32 (&r)[NL]) {
  W9 t = nl_addsub<MOD, false>(A.out + (size_t)slot * A.os + i, A.os, w9(a), w9(b));
#pragma unroll
  for (int l = 0; l < NL; l++) r[l] = t.v[l];
}
template <class MOD>
__device__ __forceinline__ void do_sub(const ChainArgs& A, u32 slot, size_t i, const u32 (&a)[NL], const u32 (&b)[NL],
                                       u32 (&r)[NL]) {
  W9 t = nl_addsub<MOD, true>(A.out + (size_t)slot * A.os + i, A.os, w9(a), w9(b));
#pragma unroll
  for (int l = 0; l < NL; l++) r[l] = t.v[l];
}
template <class MOD>
__device__ __forceinline__ u32 do_inv_known(const ChainArgs& A, u32 slot, size_t i, const u32 (&xc)[NL],
                                            const u32 (&inv)[NL]) {
  return nl_inv_known<MOD>(A.out + (size_t)slot * A.os + i, A.os, w9(xc), w9(inv));
}

// ------------------------------------------------------------------------------------------
// K0: prep-op interpreter
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
        e = do_mul<MOD>(A, op.slot, i, a, b, r);
      } else {
        for (int l = 0; l < 51; l++) st1(A, op.slot + l, i, 0);
      }
      break;
    }
    case OP_ADD:
    case OP_SUB: {
      e = load_canon<MOD>(aw, a) | load_canon<MOD>(bw, b);
      if (e) { set_zero9(a); set_zero9(b); }
      if (op.opcode == OP_ADD) do_add<MOD>(A, op.slot, i, a, b, r);
      else do_sub<MOD>(A, op.slot, i, a, b, r);
      break;
    }
    case OP_INV: {
      e = load_canon<MOD>(aw, a);
      if (e == 0 && is_zero9(a)) e = ERR_INVERSE_ZERO;
      u32 inv[NL];
      set_zero9(inv);
      if (e == 0) {
        inv_mod<MOD>(a, inv);
        e |= do_inv_known<MOD>(A, op.slot, i, a, inv);
      } else {
        st9(A, op.slot, i, inv);
        st9(A, op.slot + 9, i, inv);
      }
      break;
    }
    default: break;
  }
  return e;
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
      u32 kc[NL], k1[NL], k2[NL], n1 = 0, n2 = 0;
      u32 e = load_canon<ModN>(w, kc);
      set_zero9(k1);
      set_zero9(k2);
      if (e == 0) {
        glv_decompose_core(kc, k1, k2, n1, n2);
        u32 hi = 0;
        for (int l = 5; l < NL; l++) hi |= k1[l] | k2[l];
        if (hi) e |= ERR_LIMBS_OVERFLOW;
      }
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
// K1: Jacobian replay (secp256k1: a = 0) in lazy arithmetic.  x = X/Z^2, y = Y/Z^3.
// One thread per (signature, chain); chains are independent op ranges of the tape.
struct JacL {
  i32 X[NL], Y[NL], Z[NL];
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

template <class T>
__device__ __forceinline__ void ld9(const u32* __restrict__ base, int idx, size_t cap, size_t i, T (&v)[NL]) {
#pragma unroll
  for (int l = 0; l < NL; l++) v[l] = (T)base[((size_t)idx * NL + l) * cap + i];
}
template <class T>
__device__ __forceinline__ void sv9(u32* __restrict__ base, int idx, size_t cap, size_t i, const T (&v)[NL]) {
#pragma unroll
  for (int l = 0; l < NL; l++) base[((size_t)idx * NL + l) * cap + i] = (u32)v[l];
}

__device__ __forceinline__ void load_jac(const ChainArgs& A, int32_t pid, size_t i, JacL& P) {
  if (pid >= PID_CONST_BASE) {
    const u32* c = A.t.const_points + (size_t)(pid - PID_CONST_BASE) * 18;
#pragma unroll
    for (int l = 0; l < NL; l++) {
      P.X[l] = (i32)c[l];
      P.Y[l] = (i32)c[NL + l];
      P.Z[l] = (l == 0) ? 1 : 0;
    }
  } else {
    int idx = pt_index(A, pid);
    ld9(A.JX, idx, A.cap, i, P.X);
    ld9(A.JY, idx, A.cap, i, P.Y);
    ld9(A.JZ, idx, A.cap, i, P.Z);
  }
}

// The replay kernels run ~1 warp per SM through thousands of multiplications: keep the multiply /
// square bodies as real functions so the whole loop stays inside the instruction cache (fully
// inlined, one point addition was ~64 KB of straight-line code and instruction-fetch bound).
__device__ __noinline__ I9 lzm_v(I9 a, I9 b) {
  I9 r;
  lz_mul(a.v, b.v, r.v);
  return r;
}
__device__ __noinline__ I9 lzs_v(I9 a) {
  I9 r;
  lz_sqr(a.v, r.v);
  return r;
}
__device__ __forceinline__ void lzm(const i32 (&a)[NL], const i32 (&b)[NL], i32 (&r)[NL]) {
  I9 t = lzm_v(i9(a), i9(b));
#pragma unroll
  for (int l = 0; l < NL; l++) r[l] = t.v[l];
}
__device__ __forceinline__ void lzs(const i32 (&a)[NL], i32 (&r)[NL]) {
  I9 t = lzs_v(i9(a));
#pragma unroll
  for (int l = 0; l < NL; l++) r[l] = t.v[l];
}
#define lz_mul(a, b, r) lzm(a, b, r)
#define lz_sqr(a, r) lzs(a, r)

__device__ __forceinline__ void lzs_n(i32 (&x)[NL], int n) {
  I9 t = i9(x);
  for (int k = 0; k < n; k++) t = lzs_v(t);
#pragma unroll
  for (int l = 0; l < NL; l++) x[l] = t.v[l];
}
// x^(p-2) with the addition chain of p2e_field.cuh::lz_inv, through the out-of-line mul / sqr
__device__ void lz_inv_calls(const i32 (&a)[NL], i32 (&r)[NL]) {
  i32 x2[NL], x3[NL], x22[NL], x44[NL], x88[NL], t[NL], u[NL];
  lzs(a, t); lzm(t, a, x2);
  lzs(x2, t); lzm(t, a, x3);
  lz_copy(x3, t); lzs_n(t, 3); lzm(t, x3, u);
  lz_copy(u, t); lzs_n(t, 3); lzm(t, x3, u);
  lz_copy(u, t); lzs_n(t, 2); lzm(t, x2, u);
  lz_copy(u, t); lzs_n(t, 11); lzm(t, u, x22);
  lz_copy(x22, t); lzs_n(t, 22); lzm(t, x22, x44);
  lz_copy(x44, t); lzs_n(t, 44); lzm(t, x44, x88);
  lz_copy(x88, t); lzs_n(t, 88); lzm(t, x88, u);
  lz_copy(u, t); lzs_n(t, 44); lzm(t, x44, u);
  lz_copy(u, t); lzs_n(t, 3); lzm(t, x3, u);
  lz_copy(u, t); lzs_n(t, 23); lzm(t, x22, u);
  lz_copy(u, t); lzs_n(t, 5); lzm(t, a, u);
  lz_copy(u, t); lzs_n(t, 3); lzm(t, x2, u);
  lz_copy(u, t); lzs_n(t, 2); lzm(t, a, r);
}

// R = P + Q with Z3 = Z1*Z2*H, H = X2*Z1^2 - X1*Z2^2  (so x2 - x1 = Z3 / (Z1^3 Z2^3)).
// Q_AFFINE: Z2 == 1 (mixed addition).
template <bool Q_AFFINE>
__device__ __forceinline__ void jac_add(const JacL& P, const JacL& Q, JacL& R) {
  i32 z1z1[NL], u1[NL], u2[NL], s1[NL], s2[NL], h[NL], r[NL], t[NL], hh[NL], hhh[NL], v[NL];
  lz_sqr(P.Z, z1z1);
  if (Q_AFFINE) {
    lz_copy(P.X, u1);
    lz_copy(P.Y, s1);
  } else {
    i32 z2z2[NL];
    lz_sqr(Q.Z, z2z2);
    lz_mul(P.X, z2z2, u1);
    lz_mul(P.Y, Q.Z, t);
    lz_mul(t, z2z2, s1);
  }
  lz_mul(Q.X, z1z1, u2);
  lz_mul(Q.Y, P.Z, t);
  lz_mul(t, z1z1, s2);
  lz_sub(u2, u1, h);
  lz_sub(s2, s1, r);
  lz_sqr(h, hh);
  lz_mul(h, hh, hhh);
  lz_mul(u1, hh, v);
  lz_sqr(r, t);
  lz_sub(t, hhh, u2);
  lz_lin(u2, 1, v, 2, R.X);            // X3 = r^2 - hhh - 2v
  lz_sub(v, R.X, t);
  lz_mul(r, t, u2);
  lz_mul(s1, hhh, t);
  lz_sub(u2, t, R.Y);                  // Y3 = r (v - X3) - s1 hhh
  if (Q_AFFINE) {
    lz_mul(P.Z, h, R.Z);
  } else {
    lz_mul(P.Z, Q.Z, t);
    lz_mul(t, h, R.Z);
  }
}

// R = 2P with Z3 = 2*Y1*Z1  (so 2y = Z3 / Z1^4)
__device__ __forceinline__ void jac_dbl(const JacL& P, JacL& R) {
  i32 xx[NL], yy[NL], yyyy[NL], s[NL], m[NL], t[NL], t2[NL];
  lz_sqr(P.X, xx);
  lz_sqr(P.Y, yy);
  lz_sqr(yy, yyyy);
  lz_mul(P.X, yy, t);
  lz_lin(t, 4, t, 0, s);               // S = 4 X YY
  lz_lin(xx, 3, xx, 0, m);             // M = 3 XX
  lz_sqr(m, t);
  lz_lin(t, 1, s, 2, R.X);             // X3 = M^2 - 2S
  lz_sub(s, R.X, t);
  lz_mul(m, t, t2);
  lz_lin(t2, 1, yyyy, 8, R.Y);         // Y3 = M (S - X3) - 8 YYYY
  lz_mul(P.Y, P.Z, t);
  lz_lin(t, 2, t, 0, R.Z);             // Z3 = 2 Y1 Z1
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

__global__ void __launch_bounds__(32) k_chain_replay(ChainArgs A, int chain0) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int chain = chain0 + blockIdx.y;
  if (i >= A.n) return;
  const int k_begin = A.t.chain_begin[chain], k_end = A.t.chain_end[chain];
  if (k_begin == k_end) return;
  u32 err = 0;
  // per-signature affine input points (Z = 1); every chain writes the same values (benign)
  for (int j = 0; j < A.t.n_in_points; j++) {
    if ((int)A.t.in_points[j].round != A.round) continue;
    u64 w[NL];
    u32 x[NL], y[NL], one[NL];
    load_ref(A, A.t.in_points[j].x, i, w);
    err |= load_canon<ModP>(w, x);
    load_ref(A, A.t.in_points[j].y, i, w);
    err |= load_canon<ModP>(w, y);
    set_zero9(one);
    one[0] = 1;
    int idx = A.t.n_ptops + j;
    sv9(A.JX, idx, A.cap, i, x);
    sv9(A.JY, idx, A.cap, i, y);
    sv9(A.JZ, idx, A.cap, i, one);
    sv9(A.AX, idx, A.cap, i, x);
    sv9(A.AY, idx, A.cap, i, y);
    sv9(A.ZC, idx, A.cap, i, one);
  }
  u32 fs[NL], k1[NL], k2[NL];
  err |= load_scalar_limbs(A, A.t.fixed_scalar, (int)A.t.fixed_round == A.round, i, fs);
  err |= load_scalar_limbs(A, A.t.msm_k1, (int)A.t.msm_round == A.round, i, k1);
  err |= load_scalar_limbs(A, A.t.msm_k2, (int)A.t.msm_round == A.round, i, k2);

  int32_t cur = -1;
  for (int k = k_begin; k < k_end; k++) {
    const PtOp op = A.t.ptops[k];
    if (((op.flags >> PF_CHAIN_SHIFT) & 3) != chain || (int)op.round != A.round) continue;   // chains / rounds interleave in op order
    int32_t p1 = (op.flags & PF_P1_CUR) ? cur : resolve_pid(A, op.p1, i);
    int32_t p2 = (op.flags & PF_P2_CUR) ? cur : resolve_pid(A, op.p2, i);
    u32 flag = 1;
    if (op.type == PT_CADD_FIXED) {
      u32 d = digit_at(fs, NL, 4 * op.aux, 4);
      p2 = PID_CONST_BASE + (int32_t)(A.t.fixed_table_base + 16u * op.aux + d);
      flag = d != 0;
    } else if (op.type == PT_CADD_MSM) {
      u32 d = 4u * digit_at(k2, NL, 2 * op.aux, 2) + digit_at(k1, NL, 2 * op.aux, 2);
      p2 = A.t.pre_pid[d];
      flag = d != 0;
    } else if (op.type == PT_CADD_BOOL) {
      flag = load_wire(A, SP_IN, 9u * op.aux, i) != 0;      // BoolTarget input (0 / 1)
    }
    JacL P, R;
    load_jac(A, p1, i, P);
    if (op.type == PT_DBL) {
      jac_dbl(P, R);
      p2 = p1;
    } else {
      JacL Q;
      load_jac(A, p2, i, Q);
      if (op.flags & PF_P2_AFFINE) jac_add<true>(P, Q, R);
      else jac_add<false>(P, Q, R);
    }
    sv9(A.JX, k, A.cap, i, R.X);
    sv9(A.JY, k, A.cap, i, R.Y);
    sv9(A.JZ, k, A.cap, i, R.Z);
    A.SRC1[(size_t)k * A.cap + i] = p1;
    A.SRC2[(size_t)k * A.cap + i] = p2;
    A.FLG[(size_t)k * A.cap + i] = (uint8_t)flag;
    if (op.flags & PF_SET_CUR) cur = flag ? k : p1;
  }
  if (err) atomicOr(&A.err[i], err);
}

// ------------------------------------------------------------------------------------------
// K1 (two-lane form): lanes 2j and 2j+1 serve the same (signature, chain).  Both keep every value in
// registers; at each step of the point formulas the two lanes compute the two products of an
// INDEPENDENT pair and swap the results with a 9-word shfl_xor, so a general addition is 8
// sequential multiplications instead of 16 (mixed: 6 vs 11, doubling: 4 vs 7).  Static schedule,
// no shared memory; linear steps are done redundantly by both lanes.
__device__ __forceinline__ void duo_mul2(bool odd, const i32 (&a)[NL], const i32 (&b)[NL], i32 (&r1)[NL],
                                         const i32 (&c)[NL], const i32 (&d)[NL], i32 (&r2)[NL]) {
  I9 x, y;
#pragma unroll
  for (int l = 0; l < NL; l++) { x.v[l] = odd ? c[l] : a[l]; y.v[l] = odd ? d[l] : b[l]; }
  I9 mine = lzm_v(x, y);
#pragma unroll
  for (int l = 0; l < NL; l++) {
    i32 other = __shfl_xor_sync(0xffffffffu, mine.v[l], 1);
    r1[l] = odd ? other : mine.v[l];
    r2[l] = odd ? mine.v[l] : other;
  }
}

template <bool Q_AFFINE>
__device__ __forceinline__ void duo_add(bool odd, const JacL& P, const JacL& Q, JacL& R) {
  i32 z1z1[NL], u1[NL], u2[NL], s1[NL], s2[NL], h[NL], r[NL], t[NL], t2[NL], hh[NL], rr[NL], hhh[NL], v[NL], w[NL];
  if (Q_AFFINE) {
    lzs(P.Z, z1z1);
    duo_mul2(odd, Q.Y, P.Z, t2, Q.X, z1z1, u2);
    lz_sub(u2, P.X, h);
    duo_mul2(odd, t2, z1z1, s2, h, h, hh);
    lz_sub(s2, P.Y, r);
    duo_mul2(odd, r, r, rr, h, hh, hhh);
    duo_mul2(odd, P.X, hh, v, P.Z, h, R.Z);          // Z3 = Z1 h
    lz_sub(rr, hhh, t);
    lz_lin(t, 1, v, 2, R.X);                        // X3 = r^2 - hhh - 2v
    lz_sub(v, R.X, w);
    duo_mul2(odd, r, w, t, P.Y, hhh, t2);
    lz_sub(t, t2, R.Y);                             // Y3 = r (v - X3) - Y1 hhh
  } else {
    i32 z2z2[NL], zz[NL];
    duo_mul2(odd, P.Z, P.Z, z1z1, Q.Z, Q.Z, z2z2);
    duo_mul2(odd, P.Y, Q.Z, t, Q.Y, P.Z, t2);
    duo_mul2(odd, P.X, z2z2, u1, Q.X, z1z1, u2);
    duo_mul2(odd, t, z2z2, s1, t2, z1z1, s2);
    lz_sub(u2, u1, h);
    lz_sub(s2, s1, r);
    duo_mul2(odd, h, h, hh, r, r, rr);
    duo_mul2(odd, P.Z, Q.Z, zz, h, hh, hhh);
    duo_mul2(odd, u1, hh, v, zz, h, R.Z);            // Z3 = Z1 Z2 h
    lz_sub(rr, hhh, t);
    lz_lin(t, 1, v, 2, R.X);
    lz_sub(v, R.X, w);
    duo_mul2(odd, r, w, t, s1, hhh, t2);
    lz_sub(t, t2, R.Y);
  }
}

__device__ __forceinline__ void duo_dbl(bool odd, const JacL& P, JacL& R) {
  i32 xx[NL], yy[NL], yyyy[NL], s[NL], m[NL], t[NL], t2[NL], yz[NL], mm[NL];
  duo_mul2(odd, P.X, P.X, xx, P.Y, P.Y, yy);
  lz_lin(xx, 3, xx, 0, m);             // M = 3 XX
  duo_mul2(odd, yy, yy, yyyy, P.X, yy, t);
  lz_lin(t, 4, t, 0, s);               // S = 4 X YY
  duo_mul2(odd, P.Y, P.Z, yz, m, m, mm);
  lz_lin(mm, 1, s, 2, R.X);            // X3 = M^2 - 2S
  lz_sub(s, R.X, t);
  lzm(m, t, t2);                       // both lanes (redundant): no partner product left
  lz_lin(t2, 1, yyyy, 8, R.Y);         // Y3 = M (S - X3) - 8 YYYY
  lz_lin(yz, 2, yz, 0, R.Z);           // Z3 = 2 Y1 Z1
}

__global__ void __launch_bounds__(64) k_chain_replay_duo(ChainArgs A, int chain0) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool odd = threadIdx.x & 1;
  const size_t i_raw = tid >> 1;
  const bool valid = i_raw < A.n;
  const size_t i = valid ? i_raw : A.n - 1;       // the pair of an out-of-range task shadows the last signature
  const bool writer = valid && !odd;
  const int chain = chain0 + blockIdx.y;
  const int k_begin = A.t.chain_begin[chain], k_end = A.t.chain_end[chain];
  if (k_begin == k_end) return;
  u32 err = 0;
  if (writer) {
    for (int j = 0; j < A.t.n_in_points; j++) {
      if ((int)A.t.in_points[j].round != A.round) continue;
      u64 w[NL];
      u32 x[NL], y[NL], one[NL];
      load_ref(A, A.t.in_points[j].x, i, w);
      err |= load_canon<ModP>(w, x);
      load_ref(A, A.t.in_points[j].y, i, w);
      err |= load_canon<ModP>(w, y);
      set_zero9(one);
      one[0] = 1;
      int idx = A.t.n_ptops + j;
      sv9(A.JX, idx, A.cap, i, x);
      sv9(A.JY, idx, A.cap, i, y);
      sv9(A.JZ, idx, A.cap, i, one);
      sv9(A.AX, idx, A.cap, i, x);
      sv9(A.AY, idx, A.cap, i, y);
      sv9(A.ZC, idx, A.cap, i, one);
    }
  }
  __syncwarp();                                    // the partner lane reads the input points just written
  u32 fs[NL], k1[NL], k2[NL];
  err |= load_scalar_limbs(A, A.t.fixed_scalar, (int)A.t.fixed_round == A.round, i, fs);
  err |= load_scalar_limbs(A, A.t.msm_k1, (int)A.t.msm_round == A.round, i, k1);
  err |= load_scalar_limbs(A, A.t.msm_k2, (int)A.t.msm_round == A.round, i, k2);

  int32_t cur = -1, resident = -2;
  JacL R;
  for (int k = k_begin; k < k_end; k++) {
    const PtOp op = A.t.ptops[k];
    if (((op.flags >> PF_CHAIN_SHIFT) & 3) != chain || (int)op.round != A.round) continue;   // chains / rounds interleave in op order
    int32_t p1 = (op.flags & PF_P1_CUR) ? cur : resolve_pid(A, op.p1, i);
    int32_t p2 = (op.flags & PF_P2_CUR) ? cur : resolve_pid(A, op.p2, i);
    u32 flag = 1;
    if (op.type == PT_CADD_FIXED) {
      u32 d = digit_at(fs, NL, 4 * op.aux, 4);
      p2 = PID_CONST_BASE + (int32_t)(A.t.fixed_table_base + 16u * op.aux + d);
      flag = d != 0;
    } else if (op.type == PT_CADD_MSM) {
      u32 d = 4u * digit_at(k2, NL, 2 * op.aux, 2) + digit_at(k1, NL, 2 * op.aux, 2);
      p2 = A.t.pre_pid[d];
      flag = d != 0;
    } else if (op.type == PT_CADD_BOOL) {
      flag = load_wire(A, SP_IN, 9u * op.aux, i) != 0;
    }
    // operands: the previous result is still in registers; anything else comes from global scratch,
    // which only the even lane writes -> make those stores visible to the odd lane first
    JacL P, Q;
    if (p1 == resident) P = R; else load_jac(A, p1, i, P);
    if (op.type == PT_DBL) {
      duo_dbl(odd, P, R);
      p2 = p1;
    } else {
      if (p2 == resident) Q = R; else load_jac(A, p2, i, Q);
      if (op.flags & PF_P2_AFFINE) duo_add<true>(odd, P, Q, R);
      else duo_add<false>(odd, P, Q, R);
    }
    if (writer) {
      sv9(A.JX, k, A.cap, i, R.X);
      sv9(A.JY, k, A.cap, i, R.Y);
      sv9(A.JZ, k, A.cap, i, R.Z);
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
// K1 (cooperative form): 4 lanes per (signature, chain).  A Jacobian addition is 16 multiplications
// but only 5 dependency levels deep (doubling: 7 -> 4), so the four lanes of a task execute the
// independent multiplications of one level side by side; operands are staged in shared memory
// (value registers of 9 lazy limbs each) and every level ends with a __syncwarp().  Each stage is
// homogeneous (all MUL or all linear) so the lanes never diverge onto different code.
enum : uint8_t { MO_NOP = 0, MO_MUL = 1, MO_LIN = 2 };
struct MicroOp {
  uint8_t opc, dst, a, b, ka, kb;
};
constexpr int SP_STAGES = 10, SP_LANES = 4, SP_REGS = 28, SP_TASKS = 32, SP_PITCH = SP_TASKS + 1;
// value registers: 0-2 P (X1 Y1 Z1), 3-5 Q (X2 Y2 Z2), 6-8 result (X3 Y3 Z3), 9.. temporaries
#define MUL_(d, a, b) {MO_MUL, d, a, b, 0, 0}
#define LIN_(d, a, ka, b, kb) {MO_LIN, d, a, b, ka, kb}
#define NOP_ {MO_NOP, 0, 0, 0, 0, 0}
__constant__ MicroOp SP_TAB[3][SP_STAGES][SP_LANES] = {
    // 0: general addition   A=9 B=10 C=11 D=12 U1=13 U2=14 S1=15 S2=16 H=17 R=18 HH=19 RR=20 ZZ=21 HHH=22 V=23 T=24 W=25 RW=26 SH=27
    {{MUL_(9, 2, 2), MUL_(10, 5, 5), MUL_(11, 1, 5), MUL_(12, 4, 2)},
     {MUL_(13, 0, 10), MUL_(14, 3, 9), MUL_(15, 11, 10), MUL_(16, 12, 9)},
     {LIN_(17, 14, 1, 13, 1), LIN_(18, 16, 1, 15, 1), NOP_, NOP_},
     {MUL_(19, 17, 17), MUL_(20, 18, 18), MUL_(21, 2, 5), NOP_},
     {MUL_(22, 17, 19), MUL_(23, 13, 19), MUL_(8, 21, 17), NOP_},
     {LIN_(24, 20, 1, 22, 1), NOP_, NOP_, NOP_},
     {LIN_(6, 24, 1, 23, 2), NOP_, NOP_, NOP_},
     {LIN_(25, 23, 1, 6, 1), NOP_, NOP_, NOP_},
     {MUL_(26, 18, 25), MUL_(27, 15, 22), NOP_, NOP_},
     {LIN_(7, 26, 1, 27, 1), NOP_, NOP_, NOP_}},
    // 1: mixed addition (Z2 == 1): U1 = X1, S1 = Y1
    {{MUL_(9, 2, 2), NOP_, NOP_, MUL_(12, 4, 2)},
     {NOP_, MUL_(14, 3, 9), NOP_, MUL_(16, 12, 9)},
     {LIN_(17, 14, 1, 0, 1), LIN_(18, 16, 1, 1, 1), NOP_, NOP_},
     {MUL_(19, 17, 17), MUL_(20, 18, 18), NOP_, NOP_},
     {MUL_(22, 17, 19), MUL_(23, 0, 19), MUL_(8, 2, 17), NOP_},
     {LIN_(24, 20, 1, 22, 1), NOP_, NOP_, NOP_},
     {LIN_(6, 24, 1, 23, 2), NOP_, NOP_, NOP_},
     {LIN_(25, 23, 1, 6, 1), NOP_, NOP_, NOP_},
     {MUL_(26, 18, 25), MUL_(27, 1, 22), NOP_, NOP_},
     {LIN_(7, 26, 1, 27, 1), NOP_, NOP_, NOP_}},
    // 2: doubling   XX=9 YY=10 YZ=11 YYYY=12 TT=13 S=14 M=15 MM=16 W=17 MW=18
    {{MUL_(9, 0, 0), MUL_(10, 1, 1), MUL_(11, 1, 2), NOP_},
     {MUL_(12, 10, 10), MUL_(13, 0, 10), NOP_, NOP_},
     {LIN_(14, 13, 4, 13, 0), LIN_(15, 9, 3, 9, 0), LIN_(8, 11, 2, 11, 0), NOP_},
     {MUL_(16, 15, 15), NOP_, NOP_, NOP_},
     {LIN_(6, 16, 1, 14, 2), NOP_, NOP_, NOP_},
     {LIN_(17, 14, 1, 6, 1), NOP_, NOP_, NOP_},
     {MUL_(18, 15, 17), NOP_, NOP_, NOP_},
     {LIN_(7, 18, 1, 12, 8), NOP_, NOP_, NOP_},
     {NOP_, NOP_, NOP_, NOP_},
     {NOP_, NOP_, NOP_, NOP_}}};
__constant__ uint8_t SP_NSTAGE[3] = {10, 10, 8};
#undef MUL_
#undef LIN_
#undef NOP_

__device__ __forceinline__ void sp_ld(const i32* sh, int reg, int task, i32 (&v)[NL]) {
#pragma unroll
  for (int l = 0; l < NL; l++) v[l] = sh[(reg * NL + l) * SP_PITCH + task];
}
__device__ __forceinline__ void sp_st(i32* sh, int reg, int task, const i32 (&v)[NL]) {
#pragma unroll
  for (int l = 0; l < NL; l++) sh[(reg * NL + l) * SP_PITCH + task] = v[l];
}

// lane q in 0..2 loads coordinate q of point `pid` into value register base + q
__device__ __forceinline__ void sp_load_point(const ChainArgs& A, i32* sh, int base, int32_t pid, int32_t resident_pid,
                                              size_t i, int task, int q) {
  if (q > 2) return;
  i32 v[NL];
  if (pid == resident_pid) {
    sp_ld(sh, 6 + q, task, v);                         // the previous result is still in registers 6-8
  } else if (pid >= PID_CONST_BASE) {
    const u32* c = A.t.const_points + (size_t)(pid - PID_CONST_BASE) * 18;
#pragma unroll
    for (int l = 0; l < NL; l++) v[l] = (q == 0) ? (i32)c[l] : (q == 1) ? (i32)c[NL + l] : (l == 0);
  } else {
    int idx = pt_index(A, pid);
    const u32* src = (q == 0) ? A.JX : (q == 1) ? A.JY : A.JZ;
    ld9(src, idx, A.cap, i, v);
  }
  sp_st(sh, base + q, task, v);
}

__global__ void __launch_bounds__(128) k_chain_replay_coop(ChainArgs A, int chain0) {
  __shared__ i32 sh[SP_REGS * NL * SP_PITCH];
  const int q = threadIdx.x & 3;                       // lane role within the task
  const int task = threadIdx.x >> 2;                   // 0..31
  const size_t i_raw = (size_t)blockIdx.x * SP_TASKS + task;
  const bool valid = i_raw < A.n;
  const size_t i = valid ? i_raw : A.n - 1;            // out-of-range tasks shadow the last signature, store nothing
  const int chain = chain0 + blockIdx.y;
  const int k_begin = A.t.chain_begin[chain], k_end = A.t.chain_end[chain];
  if (k_begin == k_end) return;
  u32 err = 0;
  if (q == 0 && valid) {
    // per-signature affine input points (Z = 1); every chain writes the same values (benign)
    for (int j = 0; j < A.t.n_in_points; j++) {
      if ((int)A.t.in_points[j].round != A.round) continue;
      u64 w[NL];
      u32 x[NL], y[NL], one[NL];
      load_ref(A, A.t.in_points[j].x, i, w);
      err |= load_canon<ModP>(w, x);
      load_ref(A, A.t.in_points[j].y, i, w);
      err |= load_canon<ModP>(w, y);
      set_zero9(one);
      one[0] = 1;
      int idx = A.t.n_ptops + j;
      sv9(A.JX, idx, A.cap, i, x);
      sv9(A.JY, idx, A.cap, i, y);
      sv9(A.JZ, idx, A.cap, i, one);
      sv9(A.AX, idx, A.cap, i, x);
      sv9(A.AY, idx, A.cap, i, y);
      sv9(A.ZC, idx, A.cap, i, one);
    }
  }
  __syncwarp();
  u32 fs[NL], k1[NL], k2[NL];
  err |= load_scalar_limbs(A, A.t.fixed_scalar, (int)A.t.fixed_round == A.round, i, fs);
  err |= load_scalar_limbs(A, A.t.msm_k1, (int)A.t.msm_round == A.round, i, k1);
  err |= load_scalar_limbs(A, A.t.msm_k2, (int)A.t.msm_round == A.round, i, k2);

  int32_t cur = -1, resident = -2;
  for (int k = k_begin; k < k_end; k++) {
    const PtOp op = A.t.ptops[k];
    if (((op.flags >> PF_CHAIN_SHIFT) & 3) != chain || (int)op.round != A.round) continue;   // chains / rounds interleave in op order
    int32_t p1 = (op.flags & PF_P1_CUR) ? cur : resolve_pid(A, op.p1, i);
    int32_t p2 = (op.flags & PF_P2_CUR) ? cur : resolve_pid(A, op.p2, i);
    u32 flag = 1;
    if (op.type == PT_CADD_FIXED) {
      u32 d = digit_at(fs, NL, 4 * op.aux, 4);
      p2 = PID_CONST_BASE + (int32_t)(A.t.fixed_table_base + 16u * op.aux + d);
      flag = d != 0;
    } else if (op.type == PT_CADD_MSM) {
      u32 d = 4u * digit_at(k2, NL, 2 * op.aux, 2) + digit_at(k1, NL, 2 * op.aux, 2);
      p2 = A.t.pre_pid[d];
      flag = d != 0;
    } else if (op.type == PT_CADD_BOOL) {
      flag = load_wire(A, SP_IN, 9u * op.aux, i) != 0;
    }
    const int kind = (op.type == PT_DBL) ? 2 : (op.flags & PF_P2_AFFINE) ? 1 : 0;
    // stage operands (the just-computed result may be reused straight from shared memory)
    sp_load_point(A, sh, 0, p1, resident, i, task, q);
    if (kind != 2) sp_load_point(A, sh, 3, p2, resident, i, task, q);
    else p2 = p1;
    __syncwarp();
    const int nst = SP_NSTAGE[kind];
    for (int st = 0; st < nst; st++) {
      const MicroOp m = SP_TAB[kind][st][q];
      if (m.opc == MO_MUL) {
        i32 a[NL], b[NL];
        sp_ld(sh, m.a, task, a);
        sp_ld(sh, m.b, task, b);
        I9 r = lzm_v(i9(a), i9(b));
        sp_st(sh, m.dst, task, r.v);
      } else if (m.opc == MO_LIN) {
        i32 a[NL], b[NL], r[NL];
        sp_ld(sh, m.a, task, a);
        sp_ld(sh, m.b, task, b);
        lz_lin(a, m.ka, b, m.kb, r);
        sp_st(sh, m.dst, task, r);
      }
      __syncwarp();
    }
    if (valid && q < 3) {
      i32 v[NL];
      sp_ld(sh, 6 + q, task, v);
      u32* dst = (q == 0) ? A.JX : (q == 1) ? A.JY : A.JZ;
      sv9(dst, k, A.cap, i, v);
    }
    if (valid && q == 3) {
      A.SRC1[(size_t)k * A.cap + i] = p1;
      A.SRC2[(size_t)k * A.cap + i] = p2;
      A.FLG[(size_t)k * A.cap + i] = (uint8_t)flag;
    }
    resident = k;
    if (op.flags & PF_SET_CUR) cur = flag ? k : p1;
    __syncwarp();
  }
  if (err && valid && q == 0) atomicOr(&A.err[i], err);
}

// K2: Montgomery's trick over the Z3's of one signature, split in INV_GROUPS independent groups
// (one field inversion per group) to shorten the dependency chain.
__global__ void __launch_bounds__(32) k_chain_invert(ChainArgs A) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const int n = A.t.n_ptops;
  const int per = (n + (int)gridDim.y - 1) / (int)gridDim.y;
  const int k0 = blockIdx.y * per, k1 = (k0 + per < n) ? k0 + per : n;
  if (k0 >= k1) return;
  i32 run[NL], z[NL], t[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) run[l] = (l == 0);
  bool any = false;
  for (int k = k0; k < k1; k++) any |= (int)A.t.ptops[k].round == A.round;
  if (!any) return;                     // no curve op of this group in this round
  for (int k = k0; k < k1; k++) {
    if ((int)A.t.ptops[k].round != A.round) continue;
    sv9(A.ZINV, k, A.cap, i, run);      // prefix product of this group's Z's
    ld9(A.JZ, k, A.cap, i, z);
    lz_mul(run, z, t);
    lz_copy(t, run);
  }
  u32 rc[NL];
  lz_canon(run, rc);
  if (is_zero9(rc)) {
    // some curve op had x1 == x2 (or y == 0): the reference's inverse() panics
    atomicOr(&A.err[i], (u32)ERR_INVERSE_ZERO);
    set_zero9(rc);
    for (int k = k0; k < k1; k++)
      if ((int)A.t.ptops[k].round == A.round) sv9(A.ZINV, k, A.cap, i, rc);
    return;
  }
  i32 inv[NL];
  lz_inv_calls(run, inv);
  for (int k = k1 - 1; k >= k0; k--) {
    if ((int)A.t.ptops[k].round != A.round) continue;
    i32 pre[NL], zi[NL];
    ld9(A.ZINV, k, A.cap, i, pre);
    ld9(A.JZ, k, A.cap, i, z);
    lz_mul(inv, pre, zi);
    sv9(A.ZINV, k, A.cap, i, zi);
    lz_mul(inv, z, t);
    lz_copy(t, inv);
  }
}

// K2, tree form (default): ONE field inversion per signature instead of one per group.
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
  i32 run[NL], z[NL], t[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) run[l] = (l == 0);
  if (any) {
    for (int k = k0; k < k1; k++) {
      if ((int)A.t.ptops[k].round != A.round) continue;
      sv9(A.ZINV, k, A.cap, i, run);      // prefix product of this group's Z's
      ld9(A.JZ, k, A.cap, i, z);
      lz_mul(run, z, t);
      lz_copy(t, run);
    }
  }
  sv9(A.GP, blockIdx.y, A.cap, i, run);
}

__global__ void __launch_bounds__(32) k_chain_invert_mid(ChainArgs A, int G) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  i32 run[NL], gp[NL], t[NL];
  u32 rc[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) run[l] = (l == 0);
  for (int g = 0; g < G; g++) {
    sv9(A.GI, g, A.cap, i, run);          // product of the groups before g
    ld9(A.GP, g, A.cap, i, gp);
    lz_mul(run, gp, t);
    lz_copy(t, run);
  }
  u32 bad = 0;
  lz_canon(run, rc);
  if (is_zero9(rc)) {
    // some curve op had x1 == x2 (or y == 0): the reference's inverse() panics.  Find the groups that
    // hold a zero, leave them out of the product (their ZINV rows are zeroed by the down pass).
#pragma unroll
    for (int l = 0; l < NL; l++) run[l] = (l == 0);
    for (int g = 0; g < G; g++) {
      sv9(A.GI, g, A.cap, i, run);
      ld9(A.GP, g, A.cap, i, gp);
      lz_canon(gp, rc);
      if (is_zero9(rc)) { bad |= 1u << g; continue; }
      lz_mul(run, gp, t);
      lz_copy(t, run);
    }
    atomicOr(&A.err[i], (u32)ERR_INVERSE_ZERO);
  }
  A.BAD[i] = bad;
  i32 inv[NL];
  lz_inv_calls(run, inv);
  for (int g = G - 1; g >= 0; g--) {
    i32 pre[NL], gi[NL];
    ld9(A.GI, g, A.cap, i, pre);
    lz_mul(inv, pre, gi);
    sv9(A.GI, g, A.cap, i, gi);
    if ((bad >> g) & 1) continue;
    ld9(A.GP, g, A.cap, i, gp);
    lz_mul(inv, gp, t);
    lz_copy(t, inv);
  }
}

__global__ void __launch_bounds__(32) k_chain_invert_down(ChainArgs A) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  int k0, k1;
  if (!group_range(A, k0, k1)) return;
  if ((A.BAD[i] >> blockIdx.y) & 1) {
    u32 zero[NL];
    set_zero9(zero);
    for (int k = k0; k < k1; k++)
      if ((int)A.t.ptops[k].round == A.round) sv9(A.ZINV, k, A.cap, i, zero);
    return;
  }
  i32 inv[NL], z[NL], t[NL];
  ld9(A.GI, blockIdx.y, A.cap, i, inv);
  for (int k = k1 - 1; k >= k0; k--) {
    if ((int)A.t.ptops[k].round != A.round) continue;
    i32 pre[NL], zi[NL];
    ld9(A.ZINV, k, A.cap, i, pre);
    ld9(A.JZ, k, A.cap, i, z);
    lz_mul(inv, pre, zi);
    sv9(A.ZINV, k, A.cap, i, zi);
    lz_mul(inv, z, t);
    lz_copy(t, inv);
  }
}

// K3: canonical affine coordinates of every computed point; Z^3 stays in lazy limbs (K4 folds it
// into its own lazy product and canonicalises once)
__global__ void __launch_bounds__(128) k_chain_affine(ChainArgs A) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int k = blockIdx.y;
  if (i >= A.n) return;
  // all four loads are addressed by (k, i) alone: issue them together, ahead of the tape lookup
  i32 zi[NL], zi2[NL], zi3[NL], jx[NL], jy[NL], jz[NL], r[NL];
  u32 c[NL];
  ld9(A.ZINV, k, A.cap, i, zi);
  ld9(A.JX, k, A.cap, i, jx);
  ld9(A.JY, k, A.cap, i, jy);
  ld9(A.JZ, k, A.cap, i, jz);
  if ((int)A.t.ptops[k].round != A.round) return;
  lz_sqr(zi, zi2);
  lz_mul(zi2, zi, zi3);
  lz_mul(jx, zi2, r);
  lz_canon(r, c);
  sv9(A.AX, k, A.cap, i, c);
  lz_mul(jy, zi3, r);
  lz_canon(r, c);
  sv9(A.AY, k, A.cap, i, c);
  lz_sqr(jz, zi2);
  lz_mul(zi2, jz, r);
  sv9(A.ZC, k, A.cap, i, r);
}

#undef lz_mul
#undef lz_sqr

// K4: expand one curve op into all of its generators' output tuples
__device__ __noinline__ I9 nl_lzmul_v(I9 a, I9 b) {
  I9 r;
  lz_mul(a.v, b.v, r.v);
  return r;
}
__device__ __noinline__ W9 nl_lzcanon_v(I9 a) {
  W9 r;
  lz_canon(a.v, r.v);
  return r;
}
__device__ __forceinline__ void nl_lzmul(const i32 (&a)[NL], const i32 (&b)[NL], i32 (&r)[NL]) {
  I9 t = nl_lzmul_v(i9(a), i9(b));
#pragma unroll
  for (int l = 0; l < NL; l++) r[l] = t.v[l];
}
__device__ __forceinline__ void nl_lzcanon(const i32 (&a)[NL], u32 (&r)[NL]) {
  W9 t = nl_lzcanon_v(i9(a));
#pragma unroll
  for (int l = 0; l < NL; l++) r[l] = t.v[l];
}

__device__ __forceinline__ void load_affine(const ChainArgs& A, int32_t pid, size_t i, u32 (&x)[NL], u32 (&y)[NL],
                                            i32 (&zc)[NL]) {
  if (pid >= PID_CONST_BASE) {
    const u32* c = A.t.const_points + (size_t)(pid - PID_CONST_BASE) * 18;
#pragma unroll
    for (int l = 0; l < NL; l++) {
      x[l] = c[l];
      y[l] = c[NL + l];
      zc[l] = (l == 0) ? 1 : 0;
    }
  } else {
    int idx = pt_index(A, pid);
    ld9(A.AX, idx, A.cap, i, x);
    ld9(A.AY, idx, A.cap, i, y);
    ld9(A.ZC, idx, A.cap, i, zc);
  }
}

// 3 CTAs / SM (<= 168 registers): measured best of {none, 2, 3, 4} on B200 (0.60 ms vs 0.66 ms for 4096 signatures)
__global__ void __launch_bounds__(128, 3) k_chain_expand(ChainArgs A) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int k = blockIdx.y;
  if (i >= A.n) return;
  // Loads are issued in two dependent rounds, not five (ncu: a third of the stall samples sat on the
  // chain tape -> SRC1 -> operand 1 -> SRC2 -> operand 2): everything addressed by (k, i) alone first ...
  const size_t ki = (size_t)k * A.cap + i;
  const PtOp op = A.t.ptops[k];
  const int32_t p1 = A.SRC1[ki];
  const int32_t p2_any = A.SRC2[ki];              // meaningless for a doubling (never dereferenced then)
  const u32 flg = A.FLG[ki];
  i32 zinv[NL];
  ld9(A.ZINV, k, A.cap, i, zinv);
  if ((int)op.round != A.round) return;
  const u32 s0 = op.slot;
  u32 e = 0;
  u32 x1[NL], y1[NL], vinv[NL], x2[NL], y2[NL];
  i32 zc1[NL], zc2[NL], t[NL], t2[NL];
  // ... then both operands
  load_affine(A, p1, i, x1, y1, zc1);
  if (op.type != PT_DBL) load_affine(A, p2_any, i, x2, y2, zc2);
  if (op.type == PT_DBL) {
    // (2y)^-1 = Z1^4 / Z3
    i32 z1[NL];
    if (p1 >= PID_CONST_BASE) {
#pragma unroll
      for (int l = 0; l < NL; l++) z1[l] = (l == 0);
    } else {
      ld9(A.JZ, pt_index(A, p1), A.cap, i, z1);
    }
    nl_lzmul(zc1, z1, t);
    nl_lzmul(t, zinv, t2);
    nl_lzcanon(t2, vinv);
    u32 dy[NL], xx[NL], tr[NL], lam[NL], lam2[NL], xd[NL], x3[NL], xdiff[NL], lx[NL], y3[NL];
    do_add<ModP>(A, s0 + 0, i, y1, y1, dy);                       // double_y
    canon9<ModP>(dy);
    if (is_zero9(dy)) e |= ERR_INVERSE_ZERO;
    e |= do_inv_known<ModP>(A, s0 + 10, i, dy, vinv);             // inv_double_y
    e |= do_mul<ModP>(A, s0 + 28, i, x1, x1, xx);                 // x_squared
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
      st9(A, s0 + 79, i, tr);
      st1(A, s0 + 88, i, ov);
    }
    e |= do_mul<ModP>(A, s0 + 89, i, tr, vinv, lam);              // lambda
    e |= do_mul<ModP>(A, s0 + 140, i, lam, lam, lam2);            // lambda_squared
    do_add<ModP>(A, s0 + 191, i, x1, x1, xd);                     // x_double
    canon9<ModP>(xd);
    do_sub<ModP>(A, s0 + 201, i, lam2, xd, x3);                   // x3
    do_sub<ModP>(A, s0 + 211, i, x1, x3, xdiff);                  // x_diff
    e |= do_mul<ModP>(A, s0 + 221, i, lam, xdiff, lx);            // lambda_x_diff
    do_sub<ModP>(A, s0 + 272, i, lx, y1, y3);                     // y3
  } else {
    // (x2 - x1)^-1 = Z1^3 Z2^3 / Z3
    nl_lzmul(zc1, zc2, t);
    nl_lzmul(t, zinv, t2);
    nl_lzcanon(t2, vinv);
    u32 u[NL], v[NL], s[NL], s2[NL], xs[NL], x3[NL], xd[NL], pr[NL], y3[NL];
    do_sub<ModP>(A, s0 + 0, i, y2, y1, u);                        // u
    do_sub<ModP>(A, s0 + 10, i, x2, x1, v);                       // v
    if (is_zero9(v)) e |= ERR_INVERSE_ZERO;
    e |= do_inv_known<ModP>(A, s0 + 20, i, v, vinv);              // v_inv
    e |= do_mul<ModP>(A, s0 + 38, i, u, vinv, s);                 // s
    e |= do_mul<ModP>(A, s0 + 89, i, s, s, s2);                   // s_squared
    do_add<ModP>(A, s0 + 140, i, x2, x1, xs);                     // x_sum
    canon9<ModP>(xs);
    do_sub<ModP>(A, s0 + 150, i, s2, xs, x3);                     // x3
    do_sub<ModP>(A, s0 + 160, i, x1, x3, xd);                     // x_diff
    e |= do_mul<ModP>(A, s0 + 170, i, s, xd, pr);                 // prod
    do_sub<ModP>(A, s0 + 221, i, pr, y1, y3);                     // y3
    if (op.type != PT_ADD) {
      // curve_conditional_add: x = sum.x * b + p1.x * !b (curve.rs:235-240)
      const u32 b = flg;
      u32 z[NL], r[NL];
      set_zero9(z);
      if (b) { do_add<ModP>(A, s0 + 231, i, x3, z, r); do_add<ModP>(A, s0 + 241, i, y3, z, r); }
      else { do_add<ModP>(A, s0 + 231, i, z, x1, r); do_add<ModP>(A, s0 + 241, i, z, y1, r); }
    }
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

// ------------------------------------------------------------------------------------------
// compact host format: split the [slots][n] u64 witness into a u32 matrix (narrow slots) and a u64
// matrix (wide slots); map[slot] = (1u << 31 | wide row), (1u << 30 | narrow row) for a 34-bit carry
// (bit 29 marks the first carry of a CheckSum tuple) or the narrow row
__global__ void __launch_bounds__(256) k_pack_compact(size_t n, size_t n_slots, const u64* __restrict__ wit, size_t ws,
                                                      const u32* __restrict__ map, u32* __restrict__ narrow,
                                                      u64* __restrict__ wide, u32* __restrict__ err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (size_t s = blockIdx.y; s < n_slots; s += gridDim.y) {
    const u32 m = map[s];
    const u64 v = wit[s * ws + i];
    if (m >> 31) {
      if (wide) {
        wide[(size_t)(m & 0x7fffffffu) * n + i] = v;
      } else {
        // derived form: check_sum column j of a MUL tuple is not shipped, the host rebuilds it from the carries of the
        // CheckSum tuple that follows (a_j + c_{j-1} = 2^29 c_j exactly, c_j = b_j - 2^33; the last column is -c_15:
        // gates/mul_nonnative.rs:411-427 with every constraint zero).  Verified here; a row that is not of this form
        // (only rows whose generator already raised an error) is flagged.
        const int j = (int)((m & 0x7fffffffu) % 17u);
        const i64 cprev = j ? (i64)wit[(s + 16) * ws + i] - ((i64)1 << 33) : 0;
        const i64 ccur = j < 16 ? (i64)wit[(s + 17) * ws + i] - ((i64)1 << 33) : 0;
        const i64 expect = ccur * ((i64)1 << LB) - cprev;
        if (gl_from_i64(expect) != v) atomicOr(&err[i], (u32)ERR_CARRY_RANGE);
      }
    } else if (m & (1u << 30)) {
      // CheckSum carry b_j < 2^34: low word in its own narrow row; the 16 two-bit tops of the tuple share
      // the narrow row after them, written by the thread that handles the tuple's first slot
      const u32 row = m & 0x1fffffffu;
      narrow[(size_t)row * n + i] = (u32)v;
      if (m & (1u << 29)) {
        u32 tops = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) tops |= ((u32)(wit[(s + j) * ws + i] >> 32) & 3u) << (2 * j);
        narrow[(size_t)(row + 16) * n + i] = tops;
      }
    } else {
      narrow[(size_t)m * n + i] = (u32)v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side: device copies of the tapes, scratch, launch sequence
struct TapeOnDevice {
  bool ready = false;
  int device = -1;
  DevTape d;
  void* blob = nullptr;
  size_t n_slots = 0;
};

static std::map<int, TapeOnDevice> g_dev_tapes;
static std::mutex g_mu;

template <class T>
static size_t align_up(size_t v) { return (v + alignof(T) - 1) / alignof(T) * alignof(T); }

static int upload_tape(const Tape& t, TapeOnDevice& o) {
  Context& c = ctx();
  if (o.ready && o.device == c.device) return 0;
  size_t off = 0;
  auto reserve = [&](size_t bytes) { size_t at = (off + 15) / 16 * 16; off = at + bytes; return at; };
  size_t o_prep = reserve(t.prep.size() * sizeof(PrepOp));
  size_t o_pt = reserve(t.ptops.size() * sizeof(PtOp));
  size_t o_chk = reserve(t.checks.size() * sizeof(Check));
  size_t o_in = reserve(t.in_points.size() * sizeof(InputPoint));
  size_t o_xr = reserve((t.extra_refs.size() + 1) * sizeof(Ref));
  size_t o_cl = reserve((t.const_limbs.size() + 1) * sizeof(u32));
  size_t o_cp = reserve((t.const_points.size() + 1) * sizeof(u32));
  std::vector<unsigned char> host(off + 16, 0);
  memcpy(&host[o_prep], t.prep.data(), t.prep.size() * sizeof(PrepOp));
  memcpy(&host[o_pt], t.ptops.data(), t.ptops.size() * sizeof(PtOp));
  memcpy(&host[o_chk], t.checks.data(), t.checks.size() * sizeof(Check));
  memcpy(&host[o_in], t.in_points.data(), t.in_points.size() * sizeof(InputPoint));
  memcpy(&host[o_xr], t.extra_refs.data(), t.extra_refs.size() * sizeof(Ref));
  memcpy(&host[o_cl], t.const_limbs.data(), t.const_limbs.size() * sizeof(u32));
  memcpy(&host[o_cp], t.const_points.data(), t.const_points.size() * sizeof(u32));
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
  o.device = c.device;
  o.ready = true;
  return 0;
}

// One scratch arena per launching stream, so independent batches enqueued on different streams
// can overlap (the per-signature kernels K0-K2 are latency-bound and leave most SMs idle).
struct Arena {
  cudaStream_t stream = nullptr;
  void* ptr = nullptr;
  size_t bytes = 0;
  bool used = false;
  // internal streams: the latency-bound per-signature kernels (K0-K2) run at high priority so their
  // few small blocks are scheduled ahead of another batch's bandwidth-bound expansion (K3-K5)
  cudaStream_t hi = nullptr, lo = nullptr;
  cudaEvent_t ev_in = nullptr, ev_mid = nullptr, ev_out = nullptr;
};
constexpr int MAX_ARENAS = 16;
static Arena g_arenas[MAX_ARENAS];

static int get_arena(cudaStream_t st, size_t bytes, Arena** out) {
  Arena* a = nullptr;
  for (auto& x : g_arenas) if (x.used && x.stream == st) { a = &x; break; }
  if (!a) {
    for (auto& x : g_arenas) if (!x.used) { a = &x; break; }
    if (!a) return set_error_msg("too many distinct streams in flight (max 16 chain arenas)");
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
    if (a->ptr) P2E_CUDA(cudaFree(a->ptr));
    a->ptr = nullptr;
    a->bytes = 0;
    P2E_CUDA(cudaMalloc(&a->ptr, bytes));
    a->bytes = bytes;
  }
  *out = a;
  return 0;
}

void chain_release_device_tapes() {
  for (auto& kv : g_dev_tapes) {
    TapeOnDevice& o = kv.second;
    if (o.ready && o.blob) cudaFree(o.blob);
    o.ready = false;
    o.blob = nullptr;
  }
  for (auto& a : g_arenas) {
    if (a.ptr) cudaFree(a.ptr);
    if (a.hi) cudaStreamDestroy(a.hi);
    if (a.lo) cudaStreamDestroy(a.lo);
    if (a.ev_in) cudaEventDestroy(a.ev_in);
    if (a.ev_mid) cudaEventDestroy(a.ev_mid);
    if (a.ev_out) cudaEventDestroy(a.ev_out);
    a = Arena();
  }
}

static bool g_profiling = false;
static std::vector<cudaEvent_t> g_events;
static cudaStream_t g_event_stream = nullptr;
static void prof_mark(cudaStream_t st) {
  if (!g_profiling) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, st);
  g_events.push_back(e);
  g_event_stream = st;
}

// replay kernel: 0 one thread per (signature, chain); 1 four lanes + shared memory (measured slower,
// profiles/r01_notes.md); 2 two lanes + register shuffles.  P2E_CHAIN_REPLAY = serial | coop | duo.
static int replay_mode() {
  const char* e = getenv("P2E_CHAIN_REPLAY");
  if (!e) return P2E_DEFAULT_REPLAY;
  return e[0] == 'c' ? 1 : e[0] == 'd' ? 2 : 0;
}

// K2 form: tree (one inversion per signature, default) or groups (one inversion per group).
// P2E_CHAIN_INVERT = tree | groups
static bool invert_tree() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("P2E_CHAIN_INVERT");
    v = (e && e[0] == 'g') ? 0 : 1;
  }
  return v == 1;
}

static unsigned inv_groups() {
  static int g = 0;
  if (!g) {
    const char* e = getenv("P2E_CHAIN_INV_GROUPS");
    g = e ? atoi(e) : (invert_tree() ? 16 : 8);
    if (g < 1) g = 1;
    if (g > 32) g = 32;
  }
  return (unsigned)g;
}

// threads per block of the two full-GPU kernels (K3, K4): smaller blocks pack better next to the
// resident one-warp blocks of other batches' K0-K2 (register file granularity)
static unsigned wide_tpb() {
  static unsigned v = 0;
  if (!v) {
    const char* e = getenv("P2E_CHAIN_WIDE_TPB");
    v = e ? (unsigned)atoi(e) : 128u;
    if (v != 32 && v != 64 && v != 128) v = 128;
  }
  return v;
}

// dynamic shared memory (bytes, unused by the kernels) requested for K3 / K4 to cap their CTAs per SM; 0 = no cap
__global__ void k_chain_affine(ChainArgs A);
__global__ void k_chain_expand(ChainArgs A);
static size_t wide_smem() {
  static long v = -1;
  if (v < 0) {
    const char* e = getenv("P2E_CHAIN_WIDE_SMEM");
    v = e ? atol(e) : 0;
    if (v < 0 || v > 200 * 1024) v = 0;
    if (v > 48 * 1024) {
      cudaFuncSetAttribute(k_chain_affine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v);
      cudaFuncSetAttribute(k_chain_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v);
    }
  }
  return (size_t)v;
}

// diagnostic only (wrong results when not 63): bit mask of the kernels to launch
// 1 prep, 2 replay + join, 4 invert, 8 affine, 16 expand, 32 checks
static unsigned phase_mask() {
  const char* e = getenv("P2E_CHAIN_PHASES");     // read per call: the diagnostic script flips it in-process
  return e ? (unsigned)atoi(e) : 63u;
}

static size_t chunk_capacity() {
  const char* e = getenv("P2E_CHAIN_CHUNK");
  size_t v = e ? (size_t)atoll(e) : 0;
  return v ? v : 65536;       // 81 KB of scratch per signature: 5.3 GB per arena at full chunks (180 GB HBM); larger chunks
                              // give the one-warp-per-SM kernels more resident warps (glv, 2^16 chains: 21.1 -> 14.3 ms)
}

static int run_tape(int which, const Tape& t, size_t n, const uint64_t* const* ins, int n_in, size_t is, uint64_t* out,
                    size_t os, uint32_t* err, void* stream) {
  int rc = ensure_ready();
  if (rc) return rc;
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  if (n_in > MAX_TAPE_INPUTS) return set_error_msg("tape has more than 16 inputs");
  std::lock_guard<std::mutex> lock(g_mu);
  TapeOnDevice& o = g_dev_tapes[which];
  if ((rc = upload_tape(t, o))) return rc;
  cudaStream_t st = pick_stream(stream);
  const size_t cap = n < chunk_capacity() ? n : chunk_capacity();
  const size_t np = (size_t)o.d.n_ptops + o.d.n_in_points, no = (size_t)o.d.n_ptops;
  // scratch layout (u32 words)
  size_t words = (np * 27 * 2 + no * 9) * cap;      // JX JY JZ AX AY ZC + ZINV
  const unsigned G = inv_groups();
  size_t bytes = words * 4 + no * cap * (4 + 4 + 1) + ((size_t)G * 18 + 1) * cap * 4 + 64;
  Arena* arena = nullptr;
  if ((rc = get_arena(st, bytes, &arena))) return rc;
  void* scratch = arena->ptr;
  // profiling keeps everything on the caller's stream so the events bracket each kernel
  const bool split = !g_profiling && getenv("P2E_CHAIN_NO_PRIORITY") == nullptr;
  cudaStream_t s_hi = split ? arena->hi : st, s_lo = split ? arena->lo : st;
  if (split) {
    P2E_CUDA(cudaEventRecord(arena->ev_in, st));
    P2E_CUDA(cudaStreamWaitEvent(s_hi, arena->ev_in, 0));
    P2E_CUDA(cudaStreamWaitEvent(s_lo, arena->ev_in, 0));    // orders this call after the previous one's K3-K5 reads
  }
  ChainArgs A;
  A.t = o.d;
  A.is = is;
  A.os = os;
  A.cap = cap;
  u32* w = (u32*)scratch;
  A.JX = w; w += np * 9 * cap;
  A.JY = w; w += np * 9 * cap;
  A.JZ = w; w += np * 9 * cap;
  A.AX = w; w += np * 9 * cap;
  A.AY = w; w += np * 9 * cap;
  A.ZC = w; w += np * 9 * cap;
  A.ZINV = w; w += no * 9 * cap;
  A.SRC1 = (int32_t*)w; w += no * cap;
  A.SRC2 = (int32_t*)w; w += no * cap;
  A.FLG = (uint8_t*)w;
  w += (no * cap + 3) / 4;
  A.GP = w; w += (size_t)G * 9 * cap;
  A.GI = w; w += (size_t)G * 9 * cap;
  A.BAD = w;
  for (size_t c0 = 0; c0 < n; c0 += cap) {
    size_t cn = (n - c0 < cap) ? n - c0 : cap;
    A.n = cn;
    for (int k = 0; k < MAX_TAPE_INPUTS; k++) A.in[k] = (k < n_in) ? ins[k] + c0 : nullptr;
    A.out = out + c0;
    A.err = err + c0;
    unsigned g32 = (unsigned)((cn + 31) / 32);
    if (g_profiling && c0 == 0) {
      for (auto e : g_events) cudaEventDestroy(e);
      g_events.clear();
    }
    const bool prof = g_profiling && c0 == 0;
    if (prof) prof_mark(st);
    const int n_rounds = (int)t.n_rounds;
    const unsigned pm = phase_mask();
    for (int round = 0; round < n_rounds; round++) {
      // Round r: generators of round r (K0), then the curve ops of round r (K1-K4).  A later round
      // consumes coordinates written by this round's expansion, so it waits for the low-priority stream.
      A.round = round;
      if (split && round > 0) {
        P2E_CUDA(cudaEventRecord(arena->ev_out, s_lo));
        P2E_CUDA(cudaStreamWaitEvent(s_hi, arena->ev_out, 0));
      }
      if (pm & 1) k_prep<<<g32, 32, 0, s_hi>>>(A);
      if (prof) prof_mark(st);
      if (no) {
        if (!(pm & 2)) {
        } else if (replay_mode() == 2) {
          const unsigned g_duo = (unsigned)((2 * cn + 63) / 64);
          k_chain_replay_duo<<<dim3(g_duo, 2), 64, 0, s_hi>>>(A, 0);      // 2 lanes per (signature, chain)
          if (prof) prof_mark(st);
          if (o.d.chain_end[2] > o.d.chain_begin[2]) k_chain_replay_duo<<<dim3(g_duo, 1), 64, 0, s_hi>>>(A, 2);   // join
        } else if (replay_mode() == 1) {
          k_chain_replay_coop<<<dim3(g32, 2), 128, 0, s_hi>>>(A, 0);    // 4 lanes per (signature, chain)
          if (prof) prof_mark(st);
          if (o.d.chain_end[2] > o.d.chain_begin[2]) k_chain_replay_coop<<<dim3(g32, 1), 128, 0, s_hi>>>(A, 2);   // join
        } else {
          k_chain_replay<<<dim3(g32, 2), 32, 0, s_hi>>>(A, 0);    // fixed-base and msm chains side by side
          if (prof) prof_mark(st);
          if (o.d.chain_end[2] > o.d.chain_begin[2]) k_chain_replay<<<dim3(g32, 1), 32, 0, s_hi>>>(A, 2);   // join
        }
        if (prof) prof_mark(st);
        if (!(pm & 4)) {
        } else if (invert_tree()) {
          k_chain_invert_up<<<dim3(g32, G), 32, 0, s_hi>>>(A);
          k_chain_invert_mid<<<g32, 32, 0, s_hi>>>(A, (int)G);
          k_chain_invert_down<<<dim3(g32, G), 32, 0, s_hi>>>(A);
        } else {
          k_chain_invert<<<dim3(g32, G), 32, 0, s_hi>>>(A);
        }
        if (prof) prof_mark(st);
        if (split) {
          P2E_CUDA(cudaEventRecord(arena->ev_mid, s_hi));
          P2E_CUDA(cudaStreamWaitEvent(s_lo, arena->ev_mid, 0));
        }
        const unsigned tpb = wide_tpb(), gw = (unsigned)((cn + tpb - 1) / tpb);
        // Occupancy throttle for the two full-GPU kernels: unused dynamic shared memory caps their CTAs per SM so
        // that the latency-bound kernels of OTHER batches (high-priority streams) find registers on every SM and
        // run next to them instead of after them (K4 at 3 CTAs / SM owns the whole register file).
        const size_t wsm = wide_smem();
        if (pm & 8) k_chain_affine<<<dim3(gw, (unsigned)no), tpb, wsm, s_lo>>>(A);
        if (prof) prof_mark(st);
        if (pm & 16) k_chain_expand<<<dim3(gw, (unsigned)no), tpb, wsm, s_lo>>>(A);
        if (prof) prof_mark(st);
      } else if (split) {
        P2E_CUDA(cudaEventRecord(arena->ev_mid, s_hi));
        P2E_CUDA(cudaStreamWaitEvent(s_lo, arena->ev_mid, 0));
      }
    }
    if (pm & 32) k_chain_checks<<<g32, 32, 0, s_lo>>>(A);
    if (prof) prof_mark(st);
    if (split && c0 + cap < n) {
      // next chunk reuses the scratch: its K0-K2 must not start before this chunk's K3-K5 are done
      P2E_CUDA(cudaEventRecord(arena->ev_out, s_lo));
      P2E_CUDA(cudaStreamWaitEvent(s_hi, arena->ev_out, 0));
    }
    P2E_CUDA(cudaGetLastError());
  }
  if (split) {
    P2E_CUDA(cudaEventRecord(arena->ev_out, s_lo));
    P2E_CUDA(cudaStreamWaitEvent(st, arena->ev_out, 0));
  }
  return 0;
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

template <class Launch>
static int run_tape_host(const Tape& t, size_t n, const uint64_t* const* ins, int n_in, uint64_t* out, uint32_t* err,
                         Launch launch) {
  int rc = ensure_ready();
  if (rc) return rc;
  if (n == 0) return 0;
  cudaStream_t st = ctx().stream;
  void *din = nullptr, *dout = nullptr, *derr = nullptr;
  if ((rc = get_scratch(0, (size_t)n_in * 9 * n * 8, &din))) return rc;
  if ((rc = get_scratch(1, (size_t)t.n_slots * n * 8, &dout))) return rc;
  if ((rc = get_scratch(2, n * 4, &derr))) return rc;
  const uint64_t* dptr[MAX_TAPE_INPUTS];
  for (int k = 0; k < n_in; k++) {
    dptr[k] = (const uint64_t*)din + (size_t)k * 9 * n;
    P2E_CUDA(cudaMemcpyAsync((uint64_t*)din + (size_t)k * 9 * n, ins[k], 9 * n * 8, cudaMemcpyHostToDevice, st));
  }
  if ((rc = launch(dptr, (uint64_t*)dout, (uint32_t*)derr, (void*)st))) return rc;
  P2E_CUDA(cudaMemcpyAsync(out, dout, (size_t)t.n_slots * n * 8, cudaMemcpyDeviceToHost, st));
  if (err) P2E_CUDA(cudaMemcpyAsync(err, derr, n * 4, cudaMemcpyDeviceToHost, st));
  P2E_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace p2e

using namespace p2e;

extern "C" {

int p2e_set_profiling(int on) {
  g_profiling = on != 0;
  return 0;
}

int p2e_last_kernel_times(float* ms, int cap) {
  int rc = ensure_ready();
  if (rc) return rc;
  if (g_events.size() < 2) return 0;
  P2E_CUDA(cudaEventSynchronize(g_events.back()));
  int k = 0;
  for (size_t j = 0; j + 1 < g_events.size() && k < cap; j++, k++) {
    float t = 0;
    P2E_CUDA(cudaEventElapsedTime(&t, g_events[j], g_events[j + 1]));
    ms[k] = t;
  }
  return k;
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

int p2e_tape_witness_dev(int tape, size_t n, const uint64_t* const* inputs, size_t is, uint64_t* out, size_t os,
                         uint32_t* err, void* stream) {
  if (bad_tape(tape)) return -1;
  const Tape& t = get_tape(tape);
  return run_tape(tape, t, n, inputs, (int)t.n_inputs, is, out, os, err, stream);
}

int p2e_tape_witness(int tape, size_t n, const uint64_t* const* inputs, uint64_t* out, uint32_t* err) {
  if (bad_tape(tape)) return -1;
  const Tape& t = get_tape(tape);
  const int n_in = (int)t.n_inputs;
  return run_tape_host(t, n, inputs, n_in, out, err, [&](const uint64_t** d, uint64_t* o, uint32_t* e, void* st) {
    return run_tape(tape, t, n, d, n_in, n, o, n, e, st);
  });
}

size_t p2e_tape_compact_layout(int tape, uint32_t* map, uint32_t* narrow_slots, uint32_t* wide_slots) {
  if (bad_tape(tape)) return 0;
  const Tape& t = get_tape(tape);
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
  p2e_tape_compact_layout(tape, map.data(), &nn, &nw);
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
  const Tape& t = get_tape(tape);
  const int n_in = (int)t.n_inputs;
  static std::map<int, std::vector<uint32_t>> maps;
  static std::map<int, void*> dmaps;
  static std::map<int, int> dmap_device;
  uint32_t nn = 0, nw = 0;
  if (maps[tape].empty()) maps[tape].resize(t.n_slots);
  p2e_tape_compact_layout(tape, maps[tape].data(), &nn, &nw);
  if (!dmaps[tape] || dmap_device[tape] != ctx().device) {
    P2E_CUDA(cudaMalloc(&dmaps[tape], (size_t)t.n_slots * 4));
    P2E_CUDA(cudaMemcpy(dmaps[tape], maps[tape].data(), (size_t)t.n_slots * 4, cudaMemcpyHostToDevice));
    dmap_device[tape] = ctx().device;
  }
  cudaStream_t st = ctx().stream;
  // The batch is processed in chunks of signatures so that the D2H copy of chunk c (on a second stream,
  // pitched copies into the caller's [rows][n] matrices) overlaps the kernels of chunk c + 1: the call
  // costs about one chunk of compute plus the PCIe time instead of their sum.
  static cudaStream_t cp = nullptr;
  static int cp_device = -1;
  static cudaEvent_t ev_done[8];
  if (!cp || cp_device != ctx().device) {
    P2E_CUDA(cudaStreamCreateWithFlags(&cp, cudaStreamNonBlocking));
    for (auto& e : ev_done) P2E_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cp_device = ctx().device;
  }
  size_t n_chunks = 1;
  {
    const char* e = getenv("P2E_HOST_CHUNKS");
    n_chunks = e ? (size_t)atoi(e) : 4;
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > 8) n_chunks = 8;
    if (n < 256 * n_chunks) n_chunks = 1;
  }
  const size_t cn = (n + n_chunks - 1) / n_chunks;
  void *din = nullptr, *dout = nullptr, *derr = nullptr, *dpack = nullptr;
  if ((rc = get_scratch(0, (size_t)n_in * 9 * n * 8, &din))) return rc;
  if ((rc = get_scratch(1, (size_t)t.n_slots * cn * 8, &dout))) return rc;
  if ((rc = get_scratch(2, n * 4, &derr))) return rc;
  const size_t nbytes = (size_t)nn * cn * 4, wbytes = wide ? (size_t)nw * cn * 8 : 0;          // per chunk
  const size_t region = ((nbytes + 15) / 16) * 16 + ((wbytes + 15) / 16) * 16;
  if ((rc = get_scratch(3, region * n_chunks, &dpack))) return rc;
  const uint64_t* dptr[MAX_TAPE_INPUTS];
  for (int k = 0; k < n_in; k++)
    P2E_CUDA(cudaMemcpyAsync((uint64_t*)din + (size_t)k * 9 * n, inputs[k], 9 * n * 8, cudaMemcpyHostToDevice, st));
  for (size_t c = 0; c < n_chunks; c++) {
    const size_t c0 = c * cn;
    if (c0 >= n) break;
    const size_t cc = (n - c0 < cn) ? n - c0 : cn;
    u32* dn = (u32*)((char*)dpack + region * c);
    u64* dw = (u64*)((char*)dn + ((nbytes + 15) / 16) * 16);
    for (int k = 0; k < n_in; k++) dptr[k] = (const uint64_t*)din + (size_t)k * 9 * n + c0;
    if ((rc = run_tape(tape, t, cc, dptr, n_in, n, (uint64_t*)dout, cn, (uint32_t*)derr + c0, (void*)st))) return rc;
    k_pack_compact<<<dim3((unsigned)((cc + 255) / 256), 1024), 256, 0, st>>>(cc, t.n_slots, (const u64*)dout, cn,
                                                                            (const u32*)dmaps[tape], dn, wide ? dw : nullptr,
                                                                            (uint32_t*)derr + c0);
    P2E_CUDA(cudaGetLastError());
    P2E_CUDA(cudaEventRecord(ev_done[c], st));
    P2E_CUDA(cudaStreamWaitEvent(cp, ev_done[c], 0));
    P2E_CUDA(cudaMemcpy2DAsync(narrow + c0, n * 4, dn, cc * 4, cc * 4, nn, cudaMemcpyDeviceToHost, cp));
    if (wide) P2E_CUDA(cudaMemcpy2DAsync(wide + c0, n * 8, dw, cc * 8, cc * 8, nw, cudaMemcpyDeviceToHost, cp));
    if (err) P2E_CUDA(cudaMemcpyAsync(err + c0, (uint32_t*)derr + c0, cc * 4, cudaMemcpyDeviceToHost, cp));
  }
  P2E_CUDA(cudaStreamSynchronize(st));
  P2E_CUDA(cudaStreamSynchronize(cp));
  return 0;
}

size_t p2e_ecdsa_tape_slots(void) { return p2e_tape_slots(TAPE_ECDSA); }
size_t p2e_ecdsa_tape_layout(uint32_t* kinds, uint32_t* offsets) { return p2e_tape_layout(TAPE_ECDSA, kinds, offsets); }
int p2e_ecdsa_witness_dev(size_t n, const uint64_t* msg, const uint64_t* r, const uint64_t* s, const uint64_t* pkx,
                          const uint64_t* pky, size_t is, uint64_t* out, size_t os, uint32_t* err, void* stream) {
  const uint64_t* ins[5] = {msg, r, s, pkx, pky};
  return p2e_tape_witness_dev(TAPE_ECDSA, n, ins, is, out, os, err, stream);
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
