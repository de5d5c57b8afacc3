// p2e_ux.cu -- the plonky2_ux gates the reference's gadgets call (range check, comparison, add-many,
// subtraction, arithmetic): batched witness rows (the gates' generators) and batched
// constraint evaluation (Gate::eval_unfiltered_base_batch) over wire-major Goldilocks matrices.
//
// PARITY UNPINNED: plonky2_ux is an un-vendored, un-pinned git dependency of the reference
// (Cargo.toml:19) and no reference test fixes any of its wire maps.  These kernels implement the spec
// restated in DESIGN.md section 9 (upstream plonky2_u32 generalised to `bits`-wide limbs); every
// geometry parameter (limb width, aux-limb width, ops per row, addends) is a run-time argument.
// Reference call sites: range_check_ux_circuit gadgets/nonnative.rs:330-331,373,454-460;
// list_le_ux_circuit gadgets/biguint.rs:222-229; add_many_ux :261; sub_ux :284; mul_ux :305,335.
//
// All five evaluators are HBM-streaming kernels: one thread per (LDE row, input limb | op), rows the
// fastest index so every wire access of a warp is one 256-byte segment.  The range products
// prod_{x<2^a}(v - x) are evaluated pairwise, (v - x)(v - (2^a-1-x)) = u + x(2^a-1-x) with
// u = v (v - (2^a - 1)): 2^(a-1) Goldilocks products instead of 2^a - 1.
#include "../../include/p2e.h"
#include "p2e_common.h"
#include "p2e_field.cuh"
#include "p2e_fold.cuh"

namespace p2e {

constexpr int UX_TPB = 128;

struct UxShape {
  int gate, bits, lb, count, na;     // lb: aux / limb / chunk bits; count: input limbs or ops; na: addends
  int num_wires, num_constraints, num_inputs;
  int n_aux;                         // limbs per range-checked value
  int n_carry;                       // add-many: carry limbs
  int units;                         // threads per row (grid.y)
};

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

static int ux_shape(int gate, int bits, int lb, int count, int na, UxShape* s) {
  if (lb < 1 || lb > 4) return set_error_msg("ux gate: limb / aux / chunk bits must be 1..4");
  if (bits < 1 || bits > 48) return set_error_msg("ux gate: bits must be 1..48");
  if (count < 1 || count > 64) return set_error_msg("ux gate: count must be 1..64");
  s->gate = gate; s->bits = bits; s->lb = lb; s->count = count; s->na = na;
  s->n_aux = ceil_div(bits, lb);
  s->n_carry = 0;
  s->units = count;
  switch (gate) {
    case P2E_UX_RANGE_CHECK:
      s->num_wires = s->num_constraints = count * (1 + s->n_aux);
      s->num_inputs = count;
      break;
    case P2E_UX_COMPARISON:
      if (count != 1) return set_error_msg("ux comparison gate: count must be 1 (one comparison per row)");
      if (s->n_aux > 32) return set_error_msg("ux comparison gate: too many chunks");
      if (ceil_div(bits, s->n_aux) != lb)      // ComparisonGate::chunk_bits() = ceil(num_bits / num_chunks)
        return set_error_msg("ux comparison gate: chunk_bits must equal ceil(num_bits / ceil(num_bits / chunk_bits))");
      s->num_wires = 4 + 5 * s->n_aux + lb + 1;
      s->num_constraints = 6 + 5 * s->n_aux + lb;
      s->num_inputs = 2;
      break;
    case P2E_UX_ADD_MANY:
      if (na < 1 || na > 16) return set_error_msg("ux add-many gate: 1..16 addends");
      s->n_carry = ceil_div(4, lb);    // LOG2_MAX_NUM_ADDENDS = 4
      s->num_wires = (na + 3 + s->n_aux + s->n_carry) * count;
      s->num_constraints = (3 + s->n_aux + s->n_carry) * count;
      s->num_inputs = (na + 1) * count;
      break;
    case P2E_UX_SUBTRACTION:
      s->num_wires = (5 + s->n_aux) * count;
      s->num_constraints = (3 + s->n_aux) * count;
      s->num_inputs = 3 * count;
      break;
    case P2E_UX_ARITHMETIC:
      if (bits > 32) return set_error_msg("ux arithmetic gate: bits must be <= 32");
      s->num_wires = (6 + 2 * s->n_aux) * count;
      s->num_constraints = (4 + 2 * s->n_aux) * count;
      s->num_inputs = 3 * count;
      break;
    default:
      return set_error_msg("unknown ux gate kind");
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// Lazy Goldilocks for the evaluators.  Wires are canonical; intermediates are ANY u64 congruent to
// the value ("any"); every stored constraint is canonicalised once (glc).  The corrections use the
// carry flag (sub.cc / subc, add.cc / addc) instead of 64-bit compares + selects, which is what the
// range products were spending most of their ALU slots on.
//   glm(any, any) -> any        gla(any, canonical) -> any        gls(any, canonical) -> any
template <int A>
__device__ __forceinline__ u64 gl_shl(u64 v) { return gl_red_lazy(v << A, v >> (64 - A)); }     // any * 2^A -> any
__device__ __forceinline__ u64 gl_shl_rt(u64 v, int a) { return gl_red_lazy(v << a, v >> (64 - a)); }

// Horner recomposition sum_j a_j 2^(A j) (most significant limb first) on an un-reduced 128-bit accumulator: a constant
// 128-bit shift and a 128-bit add per limb instead of a shift-reduce-add in the field (A n_aux < bits + A <= 52, so the
// value stays below 2^116); reduced once per recomposition
template <int A>
struct Recomp {
  u64 lo = 0, hi = 0;
  __device__ __forceinline__ void push(u64 a) {
    hi = (hi << A) | (lo >> (64 - A));
    lo <<= A;
    asm("add.cc.u64 %0, %0, %2;\n\taddc.u64 %1, %1, 0;" : "+l"(lo), "+l"(hi) : "l"(a));
  }
  __device__ __forceinline__ u64 value() const { return gl_red_lazy(lo, hi); }          // any
};

// prod_{x < 2^A} (v - x) for a canonical wire v -> any
template <int A>
__device__ __forceinline__ u64 range_product(u64 v) {
  constexpr u64 B = 1ull << A;
  const u64 u = glm(v, gls(v, B - 1));            // (v - x)(v - (B-1-x)) = u + x (B-1-x)
  if (A == 1) return u;
  if (A == 3) {
    // u (u+6)(u+10)(u+12) = (z^2 - 36)(z^2 + 4 z), z = u + 6: three products for the degree-8 polynomial
    const u64 z = gla(u, 6);
    const u64 zz = glm(z, z);
    return glm(gls(zz, 36), gla(zz, glc(gl_shl<2>(z))));
  }
  u64 acc = u;
#pragma unroll
  for (u64 x = 1; x < B / 2; x++) acc = glm(acc, gla(u, x * (B - 1 - x)));
  return acc;
}

__device__ u64 gl_inv(u64 x) {       // x^(P-2), P - 2 = 0xFFFFFFFEFFFFFFFF
  u64 r = 1, b = x;
  u64 e = GL_P - 2;
  for (int k = 0; k < 64; k++) {
    if ((e >> k) & 1) r = gl_mul(r, b);
    b = gl_mul(b, b);
  }
  return r;
}

#define UX_ROW()                                             \
  const size_t i = (size_t)blockIdx.x * UX_TPB + threadIdx.x; \
  const int u = blockIdx.y;                                   \
  if (i >= n) return
#define W(k) wires[(size_t)(k) * is + i]
#define O(k) out[(size_t)(k) * os + i]

// ------------------------------------------------------------------------------------------
// evaluators
template <int A>
__global__ void __launch_bounds__(UX_TPB)
k_eval_ux_range_check(UxShape s, size_t n, const u64* __restrict__ wires, size_t is, u64* __restrict__ out, size_t os) {
  UX_ROW();
  const int aux0 = s.count + s.n_aux * u, c0 = u * (1 + s.n_aux);
  Recomp<A> acc;
#pragma unroll 4
  for (int j = s.n_aux - 1; j >= 0; j--) {
    const u64 a = W(aux0 + j);
    O(c0 + 1 + j) = glc(range_product<A>(a));
    acc.push(a);
  }
  O(c0) = glc(gls(acc.value(), W(u)));
}

template <int A>
__global__ void __launch_bounds__(UX_TPB)
k_eval_ux_comparison(UxShape s, size_t n, const u64* __restrict__ wires, size_t is, u64* __restrict__ out, size_t os) {
  UX_ROW();
  (void)u;
  const int nc = s.n_aux;
  Recomp<A> fa, sa;
  for (int c = nc - 1; c >= 0; c--) {
    fa.push(W(4 + c));
    sa.push(W(4 + nc + c));
  }
  O(0) = glc(gls(fa.value(), W(0)));
  O(1) = glc(gls(sa.value(), W(1)));
  u64 msd = 0;                                     // canonical throughout
#pragma unroll 2
  for (int c = 0; c < nc; c++) {
    const u64 f = W(4 + c), sc = W(4 + nc + c);
    const u64 dummy = W(4 + 2 * nc + c), eq = W(4 + 3 * nc + c), inter = W(4 + 4 * nc + c);
    const u64 diff = gls(sc, f);                   // canonical - canonical -> canonical
    const u64 neq = gls(1, eq);
    O(2 + 5 * c) = glc(range_product<A>(f));
    O(3 + 5 * c) = glc(range_product<A>(sc));
    O(4 + 5 * c) = glc(gls(glm(diff, dummy), neq));
    O(5 + 5 * c) = glc(glm(eq, diff));
    O(6 + 5 * c) = gls(inter, glc(glm(eq, msd)));
    msd = glc(gla(glm(neq, diff), inter));
  }
  const u64 msd_w = W(3);
  O(2 + 5 * nc) = gls(msd_w, msd);
  u64 comb = 0, top = 0;
  for (int b = A; b >= 0; b--) {
    const u64 bit = W(4 + 5 * nc + b);
    if (b == A) top = bit;
    O(3 + 5 * nc + b) = glc(glm(bit, gls(1, bit)));
    comb = gla(gl_shl<1>(comb), bit);
  }
  O(4 + 5 * nc + A) = glc(gls(gla(1ull << A, msd_w), glc(comb)));
  O(5 + 5 * nc + A) = gls(W(2), top);
}

template <int A>
__global__ void __launch_bounds__(UX_TPB)
k_eval_ux_add_many(UxShape s, size_t n, const u64* __restrict__ wires, size_t is, u64* __restrict__ out, size_t os) {
  UX_ROW();
  const int r0 = (s.na + 3) * u, nl = s.n_aux + s.n_carry;
  const int l0 = (s.na + 3) * s.count + nl * u, c0 = (3 + nl) * u;
  u64 computed = W(r0 + s.na);
  for (int j = 0; j < s.na; j++) computed = gla(computed, W(r0 + j));
  const u64 res = W(r0 + s.na + 1), car = W(r0 + s.na + 2);
  O(c0) = glc(gls(gla(gl_shl_rt(car, s.bits), res), glc(computed)));
  Recomp<A> cres, ccar;
#pragma unroll 4
  for (int j = nl - 1; j >= 0; j--) {
    const u64 limb = W(l0 + j);
    O(c0 + 1 + (nl - 1 - j)) = glc(range_product<A>(limb));
    if (j < s.n_aux) cres.push(limb);
    else ccar.push(limb);
  }
  O(c0 + 1 + nl) = glc(gls(cres.value(), res));
  O(c0 + 2 + nl) = glc(gls(ccar.value(), car));
}

template <int A>
__global__ void __launch_bounds__(UX_TPB)
k_eval_ux_subtraction(UxShape s, size_t n, const u64* __restrict__ wires, size_t is, u64* __restrict__ out, size_t os) {
  UX_ROW();
  const int nl = s.n_aux, l0 = 5 * s.count + nl * u, c0 = (3 + nl) * u;
  const u64 x = W(5 * u), y = W(5 * u + 1), bin = W(5 * u + 2), res = W(5 * u + 3), bout = W(5 * u + 4);
  const u64 init = gls(gls(x, y), bin);            // canonical
  O(c0) = gls(res, glc(gla(gl_shl_rt(bout, s.bits), init)));
  Recomp<A> comb;
#pragma unroll 4
  for (int j = nl - 1; j >= 0; j--) {
    const u64 limb = W(l0 + j);
    O(c0 + 1 + (nl - 1 - j)) = glc(range_product<A>(limb));
    comb.push(limb);
  }
  O(c0 + 1 + nl) = glc(gls(comb.value(), res));
  O(c0 + 2 + nl) = glc(glm(bout, gls(1, bout)));
}

template <int A>
__global__ void __launch_bounds__(UX_TPB)
k_eval_ux_arithmetic(UxShape s, size_t n, const u64* __restrict__ wires, size_t is, u64* __restrict__ out, size_t os) {
  UX_ROW();
  const int half = s.n_aux, nl = 2 * half, l0 = 6 * s.count + nl * u, c0 = (4 + nl) * u;
  const u64 m0 = W(6 * u), m1 = W(6 * u + 1), add = W(6 * u + 2), lo = W(6 * u + 3), hi = W(6 * u + 4), inv = W(6 * u + 5);
  const u64 computed = glc(gla(glm(m0, m1), add));
  const u64 ux_max = (1ull << s.bits) - 1;
  const u64 hi_not_max = gls(glm(inv, gls(ux_max, hi)), 1);
  O(c0) = glc(glm(hi_not_max, lo));
  O(c0 + 1) = glc(gls(gla(gl_shl_rt(hi, s.bits), lo), computed));
  Recomp<A> clo, chi;
#pragma unroll 4
  for (int j = nl - 1; j >= 0; j--) {
    const u64 limb = W(l0 + j);
    O(c0 + 2 + (nl - 1 - j)) = glc(range_product<A>(limb));
    if (j < half) clo.push(limb);
    else chi.push(limb);
  }
  O(c0 + 2 + nl) = glc(gls(clo.value(), lo));
  O(c0 + 3 + nl) = glc(gls(chi.value(), hi));
}

// ------------------------------------------------------------------------------------------
// generators: inputs [num_inputs][is] -> full gate rows [num_wires][os]
#define IN(k) in[(size_t)(k) * is + i]
#define ROW(k) row[(size_t)(k) * os + i]

__global__ void __launch_bounds__(UX_TPB)
k_ux_range_check_witness(UxShape s, size_t n, const u64* __restrict__ in, size_t is, u64* __restrict__ row, size_t os,
                         u32* __restrict__ err) {
  UX_ROW();
  u64 v = IN(u);
  ROW(u) = v;
  const u64 mask = (1ull << s.lb) - 1;
  for (int j = 0; j < s.n_aux; j++) {
    ROW(s.count + s.n_aux * u + j) = v & mask;
    v >>= s.lb;
  }
  if (v) atomicOr(&err[i], (u32)ERR_LIMB_RANGE);
}

struct InvTable { u64 v[16]; };      // inverses of 1..15 in Goldilocks (v[0] unused)

__global__ void __launch_bounds__(UX_TPB)
k_ux_comparison_witness(UxShape s, InvTable inv, size_t n, const u64* __restrict__ in, size_t is, u64* __restrict__ row,
                        size_t os, u32* __restrict__ err) {
  UX_ROW();
  (void)u;
  const int nc = s.n_aux, cb = s.lb;
  const u64 first = IN(0), second = IN(1);
  ROW(0) = first;
  ROW(1) = second;
  ROW(2) = first <= second;
  const u64 mask = (1ull << cb) - 1;
  u64 f = first, sv = second, msd = 0;
  for (int c = 0; c < nc; c++) {
    const u64 fc = f & mask, sc = sv & mask;
    f >>= cb;
    sv >>= cb;
    ROW(4 + c) = fc;
    ROW(4 + nc + c) = sc;
    if (fc == sc) {
      ROW(4 + 2 * nc + c) = 1;
      ROW(4 + 3 * nc + c) = 1;
      ROW(4 + 4 * nc + c) = msd;
    } else {
      ROW(4 + 2 * nc + c) = sc > fc ? inv.v[sc - fc] : GL_P - inv.v[fc - sc];
      ROW(4 + 3 * nc + c) = 0;
      ROW(4 + 4 * nc + c) = 0;
      msd = gl_sub(sc, fc);
    }
  }
  ROW(3) = msd;
  u64 t = gl_add(1ull << cb, msd);
  for (int b = 0; b <= cb; b++) {
    ROW(4 + 5 * nc + b) = t & 1;
    t >>= 1;
  }
  err[i] = (f | sv) ? (u32)ERR_LIMB_RANGE : 0u;
}

__global__ void __launch_bounds__(UX_TPB)
k_ux_add_many_witness(UxShape s, size_t n, const u64* __restrict__ in, size_t is, u64* __restrict__ row, size_t os,
                      u32* __restrict__ err) {
  UX_ROW();
  const int r0 = (s.na + 3) * u, nl = s.n_aux + s.n_carry, l0 = (s.na + 3) * s.count + nl * u;
  u64 total = 0;
  for (int j = 0; j <= s.na; j++) {          // addends, then the carry-in
    const u64 a = IN((s.na + 1) * u + j);
    ROW(r0 + j) = a;
    total = gl_add(total, a);
  }
  u64 res = total & ((1ull << s.bits) - 1), car = total >> s.bits;
  ROW(r0 + s.na + 1) = res;
  ROW(r0 + s.na + 2) = car;
  const u64 mask = (1ull << s.lb) - 1;
  for (int j = 0; j < s.n_aux; j++) { ROW(l0 + j) = res & mask; res >>= s.lb; }
  for (int j = 0; j < s.n_carry; j++) { ROW(l0 + s.n_aux + j) = car & mask; car >>= s.lb; }
  if (car) atomicOr(&err[i], (u32)ERR_LIMB_RANGE);
}

__global__ void __launch_bounds__(UX_TPB)
k_ux_subtraction_witness(UxShape s, size_t n, const u64* __restrict__ in, size_t is, u64* __restrict__ row, size_t os,
                         u32* __restrict__ err) {
  UX_ROW();
  const u64 x = IN(3 * u), y = IN(3 * u + 1), b = IN(3 * u + 2);
  ROW(5 * u) = x;
  ROW(5 * u + 1) = y;
  ROW(5 * u + 2) = b;
  const u64 init = gl_sub(gl_sub(x, y), b);
  const u64 bout = init > (1ull << s.bits);
  u64 res = gl_add(init, bout << s.bits);
  ROW(5 * u + 3) = res;
  ROW(5 * u + 4) = bout;
  const u64 mask = (1ull << s.lb) - 1;
  const int l0 = 5 * s.count + s.n_aux * u;
  for (int j = 0; j < s.n_aux; j++) { ROW(l0 + j) = res & mask; res >>= s.lb; }
  if (res) atomicOr(&err[i], (u32)ERR_LIMB_RANGE);
}

__global__ void __launch_bounds__(UX_TPB)
k_ux_arithmetic_witness(UxShape s, size_t n, const u64* __restrict__ in, size_t is, u64* __restrict__ row, size_t os,
                        u32* __restrict__ err) {
  UX_ROW();
  const u64 m0 = IN(3 * u), m1 = IN(3 * u + 1), add = IN(3 * u + 2);
  ROW(6 * u) = m0;
  ROW(6 * u + 1) = m1;
  ROW(6 * u + 2) = add;
  const u64 o = gl_add(gl_mul(m0, m1), add);
  const u64 ux_max = (1ull << s.bits) - 1;
  u64 lo = o & ux_max, hi = o >> s.bits;
  ROW(6 * u + 3) = lo;
  ROW(6 * u + 4) = hi;
  const u64 d = gl_sub(ux_max, hi);
  ROW(6 * u + 5) = d ? gl_inv(d) : 0;
  const u64 mask = (1ull << s.lb) - 1;
  const int l0 = 6 * s.count + 2 * s.n_aux * u;
  for (int j = 0; j < s.n_aux; j++) { ROW(l0 + j) = lo & mask; lo >>= s.lb; }
  for (int j = 0; j < s.n_aux; j++) { ROW(l0 + s.n_aux + j) = hi & mask; hi >>= s.lb; }
  if (hi) atomicOr(&err[i], (u32)ERR_LIMB_RANGE);
}

// ------------------------------------------------------------------------------------------
static u64 host_gl_mul(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) % GL_P); }
static u64 host_gl_inv(u64 x) {
  u64 r = 1, b = x, e = GL_P - 2;
  for (; e; e >>= 1) {
    if (e & 1) r = host_gl_mul(r, b);
    b = host_gl_mul(b, b);
  }
  return r;
}

// Range-check gate, fused with the selector filter and the challenge fold: one thread per LDE row walks the
// `count` input limbs; every constraint (recomposition, then one range product per aux limb, in the gate's order) is
// multiplied into un-reduced 96-bit accumulators per challenge instead of being canonicalised and stored -- 2 columns
// written instead of num_constraints, and no second pass over them.  pw: [NA][num_constraints] powers of the challenges.
template <int A, int NA>
__global__ void __launch_bounds__(UX_TPB)
k_eval_ux_range_check_fc(UxShape s, size_t n, const u64* __restrict__ wires, size_t is, const u64* __restrict__ pw,
                         FilterArgs F, u64* __restrict__ out, size_t os) {
  const size_t i = (size_t)blockIdx.x * UX_TPB + threadIdx.x;
  if (i >= n) return;
  Acc3 SA[NA], SB[NA], SC[NA];
#pragma unroll
  for (int a = 0; a < NA; a++) SA[a] = SB[a] = SC[a] = Acc3{0, 0, 0};
  const int C = s.num_constraints;
  auto fold = [&](u64 v, int c) {
    const u32 vl = (u32)v, vh = (u32)(v >> 32);
#pragma unroll
    for (int a = 0; a < NA; a++) {
      const u64 p = __ldg(pw + (size_t)a * C + c);
      const u32 pl = (u32)p, ph = (u32)(p >> 32);
      mac3(SA[a], vl, pl);
      mac3(SB[a], vl, ph);
      mac3(SB[a], vh, pl);
      mac3(SC[a], vh, ph);
    }
  };
  for (int u = 0; u < s.count; u++) {
    const int aux0 = s.count + s.n_aux * u, c0 = u * (1 + s.n_aux);
    const u64 input = W(u);
    Recomp<A> acc;
#pragma unroll 4
    for (int j = s.n_aux - 1; j >= 0; j--) {
      const u64 a = W(aux0 + j);
      fold(range_product<A>(a), c0 + 1 + j);
      acc.push(a);
    }
    fold(gls(acc.value(), input), c0);
  }
  const u64 f = compute_filter(F, i);
#pragma unroll
  for (int a = 0; a < NA; a++) {
    if (a < F.n_alphas) {
      u64 r = gl_mul(f, reduce_abc(SA[a], SB[a], SC[a], 0));
      if (F.accumulate) r = gl_add(r, out[a * os + i]);
      out[a * os + i] = r;
    }
  }
}

template <int A>
static void launch_rc_fc(const UxShape& s, size_t n, const u64* wires, size_t is, const u64* pw, const FilterArgs& F, u64* out,
                         size_t os, cudaStream_t st) {
  const unsigned g = (unsigned)((n + UX_TPB - 1) / UX_TPB);
  if (F.n_alphas == 1) k_eval_ux_range_check_fc<A, 1><<<g, UX_TPB, 0, st>>>(s, n, wires, is, pw, F, out, os);
  else if (F.n_alphas == 2) k_eval_ux_range_check_fc<A, 2><<<g, UX_TPB, 0, st>>>(s, n, wires, is, pw, F, out, os);
  else k_eval_ux_range_check_fc<A, 4><<<g, UX_TPB, 0, st>>>(s, n, wires, is, pw, F, out, os);
}

// fused filter + fold for the range-check gate; returns 1 when this gate has no fused form (caller takes the generic path)
int ux_eval_filtered_fused(int gate, int bits, int lb, int count, int na, size_t n, const u64* wires, size_t is, const u64* pw,
                           const FilterArgs& F, u64* out, size_t os, cudaStream_t st) {
  if (gate != P2E_UX_RANGE_CHECK) return 1;
  UxShape s;
  int rc = ux_shape(gate, bits, lb, count, na, &s);
  if (rc) return rc;
  switch (lb) {
    case 1: launch_rc_fc<1>(s, n, wires, is, pw, F, out, os, st); break;
    case 2: launch_rc_fc<2>(s, n, wires, is, pw, F, out, os, st); break;
    case 3: launch_rc_fc<3>(s, n, wires, is, pw, F, out, os, st); break;
    default: launch_rc_fc<4>(s, n, wires, is, pw, F, out, os, st); break;
  }
  P2E_CUDA(cudaGetLastError());
  return 0;
}

template <int A>
static void launch_eval(const UxShape& s, size_t n, const u64* wires, size_t is, u64* out, size_t os, cudaStream_t st) {
  dim3 g((unsigned)((n + UX_TPB - 1) / UX_TPB), (unsigned)s.units);
  switch (s.gate) {
    case P2E_UX_RANGE_CHECK: k_eval_ux_range_check<A><<<g, UX_TPB, 0, st>>>(s, n, wires, is, out, os); break;
    case P2E_UX_COMPARISON: k_eval_ux_comparison<A><<<g, UX_TPB, 0, st>>>(s, n, wires, is, out, os); break;
    case P2E_UX_ADD_MANY: k_eval_ux_add_many<A><<<g, UX_TPB, 0, st>>>(s, n, wires, is, out, os); break;
    case P2E_UX_SUBTRACTION: k_eval_ux_subtraction<A><<<g, UX_TPB, 0, st>>>(s, n, wires, is, out, os); break;
    default: k_eval_ux_arithmetic<A><<<g, UX_TPB, 0, st>>>(s, n, wires, is, out, os); break;
  }
}

}  // namespace p2e

using namespace p2e;

extern "C" {

int p2e_ux_gate_shape(int gate, int bits, int limb_bits, int count, int num_addends, int* num_wires,
                      int* num_constraints, int* num_inputs) {
  UxShape s;
  int rc = ux_shape(gate, bits, limb_bits, count, num_addends, &s);
  if (rc) return rc;
  if (num_wires) *num_wires = s.num_wires;
  if (num_constraints) *num_constraints = s.num_constraints;
  if (num_inputs) *num_inputs = s.num_inputs;
  return 0;
}

int p2e_eval_ux_dev(int gate, int bits, int limb_bits, int count, int num_addends, size_t n, const uint64_t* wires,
                    size_t is, uint64_t* out, size_t os, void* stream) {
  int rc = ensure_ready();
  if (rc) return rc;
  UxShape s;
  if ((rc = ux_shape(gate, bits, limb_bits, count, num_addends, &s))) return rc;
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  cudaStream_t st = pick_stream(stream);
  switch (limb_bits) {
    case 1: launch_eval<1>(s, n, wires, is, out, os, st); break;
    case 2: launch_eval<2>(s, n, wires, is, out, os, st); break;
    case 3: launch_eval<3>(s, n, wires, is, out, os, st); break;
    default: launch_eval<4>(s, n, wires, is, out, os, st); break;
  }
  P2E_CUDA(cudaGetLastError());
  return 0;
}

int p2e_ux_witness_dev(int gate, int bits, int limb_bits, int count, int num_addends, size_t n, const uint64_t* inputs,
                       size_t is, uint64_t* row, size_t os, uint32_t* err, void* stream) {
  int rc = ensure_ready();
  if (rc) return rc;
  UxShape s;
  if ((rc = ux_shape(gate, bits, limb_bits, count, num_addends, &s))) return rc;
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  cudaStream_t st = pick_stream(stream);
  dim3 g((unsigned)((n + UX_TPB - 1) / UX_TPB), (unsigned)s.units);
  P2E_CUDA(cudaMemsetAsync(err, 0, n * sizeof(uint32_t), st));
  switch (gate) {
    case P2E_UX_RANGE_CHECK: k_ux_range_check_witness<<<g, UX_TPB, 0, st>>>(s, n, inputs, is, row, os, err); break;
    case P2E_UX_COMPARISON: {
      static InvTable tab;
      static bool have = false;
      if (!have) {
        tab.v[0] = 0;
        for (u64 d = 1; d < 16; d++) tab.v[d] = host_gl_inv(d);
        have = true;
      }
      k_ux_comparison_witness<<<g, UX_TPB, 0, st>>>(s, tab, n, inputs, is, row, os, err);
      break;
    }
    case P2E_UX_ADD_MANY: k_ux_add_many_witness<<<g, UX_TPB, 0, st>>>(s, n, inputs, is, row, os, err); break;
    case P2E_UX_SUBTRACTION: k_ux_subtraction_witness<<<g, UX_TPB, 0, st>>>(s, n, inputs, is, row, os, err); break;
    default: k_ux_arithmetic_witness<<<g, UX_TPB, 0, st>>>(s, n, inputs, is, row, os, err); break;
  }
  P2E_CUDA(cudaGetLastError());
  return 0;
}

// host-pointer forms: H2D, kernel, D2H, synchronise (stride == n)
static int ux_host(size_t n, const uint64_t* in, size_t in_rows, uint64_t* out, size_t out_rows, uint32_t* err,
                   int (*launch)(void*, const uint64_t*, uint64_t*, uint32_t*, void*), void* closure) {
  int rc = ensure_ready();
  if (rc) return rc;
  if (n == 0) return 0;
  P2E_HOST_LOCK();
  cudaStream_t st = ctx().stream;
  void *din = nullptr, *dout = nullptr, *derr = nullptr;
  if ((rc = get_scratch(0, in_rows * n * 8, &din))) return rc;
  if ((rc = get_scratch(1, out_rows * n * 8, &dout))) return rc;
  if ((rc = get_scratch(2, n * 4, &derr))) return rc;
  P2E_CUDA(cudaMemcpyAsync(din, in, in_rows * n * 8, cudaMemcpyHostToDevice, st));
  if ((rc = launch(closure, (const uint64_t*)din, (uint64_t*)dout, (uint32_t*)derr, (void*)st))) return rc;
  P2E_CUDA(cudaMemcpyAsync(out, dout, out_rows * n * 8, cudaMemcpyDeviceToHost, st));
  if (err) P2E_CUDA(cudaMemcpyAsync(err, derr, n * 4, cudaMemcpyDeviceToHost, st));
  P2E_CUDA(cudaStreamSynchronize(st));
  return 0;
}

struct UxCall { int gate, bits, lb, count, na; size_t n; };

int p2e_eval_ux(int gate, int bits, int limb_bits, int count, int num_addends, size_t n, const uint64_t* wires,
                uint64_t* out) {
  UxShape s;
  int rc = ux_shape(gate, bits, limb_bits, count, num_addends, &s);
  if (rc) return rc;
  UxCall c = {gate, bits, limb_bits, count, num_addends, n};
  return ux_host(n, wires, (size_t)s.num_wires, out, (size_t)s.num_constraints, nullptr,
                 [](void* p, const uint64_t* d, uint64_t* o, uint32_t*, void* st) {
                   UxCall* c = (UxCall*)p;
                   return p2e_eval_ux_dev(c->gate, c->bits, c->lb, c->count, c->na, c->n, d, c->n, o, c->n, st);
                 }, &c);
}

int p2e_ux_witness(int gate, int bits, int limb_bits, int count, int num_addends, size_t n, const uint64_t* inputs,
                   uint64_t* row, uint32_t* err) {
  UxShape s;
  int rc = ux_shape(gate, bits, limb_bits, count, num_addends, &s);
  if (rc) return rc;
  UxCall c = {gate, bits, limb_bits, count, num_addends, n};
  return ux_host(n, inputs, (size_t)s.num_inputs, row, (size_t)s.num_wires, err,
                 [](void* p, const uint64_t* d, uint64_t* o, uint32_t* e, void* st) {
                   UxCall* c = (UxCall*)p;
                   return p2e_ux_witness_dev(c->gate, c->bits, c->lb, c->count, c->na, c->n, d, c->n, o, c->n, e, st);
                 }, &c);
}

}  // extern "C"
