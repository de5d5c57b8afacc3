// p2e_genstream.cu -- serialized generator streams run by a level-scheduled executor
// (SURVEY 8f-2): the GPU-resident counterpart of plonky2's generate_partial_witness for circuits whose
// witness is produced by the reference's in-tree generators.
//
// The stream is what a Rust host obtains by calling SimpleGenerator::serialize on each generator
// (gates/mul_nonnative.rs:326-330, 533-536; gadgets/nonnative.rs:647-652, 730-740, 812-817, 874-878;
// gadgets/biguint.rs:520-526; gadgets/glv.rs:144-150), each record prefixed with a u32 tag the way plonky2's
// WitnessGeneratorSerializer::write_generator prefixes its registry index (the registry is the host's; the
// tags are P2E_GEN_* of include/p2e.h and also carry the FF type parameter, which serialize() does not write).
// Primitive encodings are plonky2's util::serialization (out of tree, restated): usize / u64 field element =
// 8 bytes little-endian, u32 = 4 bytes LE, bool = 1 byte, Target = bool is_wire then (row, column) or (index),
// target_vec = usize length then the targets.
//
// Execution model.  Targets joined by copy constraints share one value (plonky2's PartitionWitness):
// P2E_GEN_COPY records feed a union-find and every partition becomes one SLOT, a row of the device table
// V[slot][instance].  A generator's level is 1 + the highest level among the producers of the slots it
// reads; generators of one level with the same kind / modulus / limb shape form a GROUP that runs as ONE
// launch over (instances of the group) x (n circuits): a fused kernel reads the operand rows through the group's
// slot lists, runs the same arithmetic core as the standalone batched kernel and writes the output rows back (a
// slot that already holds a value is compared instead, like plonky2's "partition set twice").  The staged form
// (gather kernel -> p2e_*_witness_dev -> scatter kernel) remains for BigUint div_rem and as a cross-check
// (P2E_GENSTREAM_FUSED=0).
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/p2e.h"
#include "p2e_common.h"
#include "p2e_field.cuh"
#include "p2e_f8.cuh"

namespace p2e {

constexpr u32 GS_NONE = 0xFFFFFFFFu;      // absent operand limb (reads 0) / discarded output row
constexpr u32 GS_CHECK = 0x80000000u;     // output slot already has a producer: compare, do not write
constexpr u64 GS_P = 0xFFFFFFFF00000001ull;
constexpr int GS_TPB = 128;
constexpr size_t GS_MAX_STAGED = (size_t)1 << 21;   // staged instances per launch (bounds the staging buffers)

enum GsKind { K_MUL, K_CHECKSUM, K_ADD, K_MULTIADD, K_SUB, K_INV, K_DIVREM, K_GLV, K_ARITH, K_CONST, K_BASESPLIT, K_RANDACC, K_EQUAL };

// ---- kernels --------------------------------------------------------------------------------------------
// staging[r][j * n + i] = V[idx[r * m + j0 + j]][i]   (GS_NONE -> 0)
__global__ void __launch_bounds__(GS_TPB)
k_gs_gather(size_t n, int m, int mc, int j0, int rows, const u32* __restrict__ idx, const u64* __restrict__ V, size_t vs,
            u64* __restrict__ stage) {
  const size_t i = (size_t)blockIdx.x * GS_TPB + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= n) return;
  const size_t M = (size_t)mc * n;
  for (int r = 0; r < rows; r++) {
    const u32 s = idx[(size_t)r * m + j0 + j];
    stage[(size_t)r * M + (size_t)j * n + i] = (s == GS_NONE) ? 0 : V[(size_t)s * vs + i];
  }
}

// V[oidx[r * m + j0 + j]][i] = stage[r][j * n + i]; GS_CHECK slots are compared; err[i] |= staged error bits
__global__ void __launch_bounds__(GS_TPB)
k_gs_scatter(size_t n, int m, int mc, int j0, int rows, const u32* __restrict__ oidx, u64* __restrict__ V, size_t vs,
             const u64* __restrict__ stage, const u32* __restrict__ stage_err, u32* __restrict__ err) {
  const size_t i = (size_t)blockIdx.x * GS_TPB + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= n) return;
  const size_t M = (size_t)mc * n;
  u32 e = stage_err ? stage_err[(size_t)j * n + i] : 0;
  for (int r = 0; r < rows; r++) {
    const u32 s = oidx[(size_t)r * m + j0 + j];
    if (s == GS_NONE) continue;
    const u64 v = stage[(size_t)r * M + (size_t)j * n + i];
    if (s & GS_CHECK) {
      if (V[(size_t)(s & ~GS_CHECK) * vs + i] != v) e |= P2E_ERR_CONNECT;
    } else {
      V[(size_t)s * vs + i] = v;
    }
  }
  if (e) atomicOr(&err[i], e);
}

__device__ __forceinline__ u64 gs_mul(u64 a, u64 b) {
  const u64 lo = a * b, hi = __umul64hi(a, b);
  // (hi:lo) mod P: 2^64 = 2^32 - 1, 2^96 = -1
  const u64 hh = hi >> 32, hl = hi & 0xFFFFFFFFull;
  u64 t = lo - hh;
  if (lo < hh) t -= 0xFFFFFFFFull;
  const u64 p = hl * 0xFFFFFFFFull;
  u64 r = t + p;
  if (r < p) r += 0xFFFFFFFFull;
  if (r >= GS_P) r -= GS_P;
  return r;
}
__device__ __forceinline__ u64 gs_add(u64 a, u64 b) {
  u64 r = a + b;
  if (r < a) r += 0xFFFFFFFFull;
  if (r >= GS_P) r -= GS_P;
  return r;
}

// plonky2 ArithmeticBaseGenerator (restated): out = c0 * m0 * m1 + c1 * addend over Goldilocks
__global__ void __launch_bounds__(GS_TPB)
k_gs_arith(size_t n, int m, const u32* __restrict__ idx, const u64* __restrict__ consts, u64* __restrict__ V, size_t vs,
           u32* __restrict__ err) {
  const size_t i = (size_t)blockIdx.x * GS_TPB + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= n) return;
  const u64 a = V[(size_t)idx[j] * vs + i], b = V[(size_t)idx[m + j] * vs + i], c = V[(size_t)idx[2 * m + j] * vs + i];
  const u64 v = gs_add(gs_mul(consts[j], gs_mul(a, b)), gs_mul(consts[m + j], c));
  const u32 s = idx[3 * m + j];
  if (s & GS_CHECK) {
    if (V[(size_t)(s & ~GS_CHECK) * vs + i] != v) atomicOr(&err[i], (u32)P2E_ERR_CONNECT);
  } else {
    V[(size_t)s * vs + i] = v;
  }
}

// plonky2 ConstantGenerator (restated): out = constant
__global__ void __launch_bounds__(GS_TPB)
k_gs_const(size_t n, int m, const u32* __restrict__ idx, const u64* __restrict__ consts, u64* __restrict__ V, size_t vs,
           u32* __restrict__ err) {
  const size_t i = (size_t)blockIdx.x * GS_TPB + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= n) return;
  const u32 s = idx[j];
  if (s & GS_CHECK) {
    if (V[(size_t)(s & ~GS_CHECK) * vs + i] != consts[j]) atomicOr(&err[i], (u32)P2E_ERR_CONNECT);
  } else {
    V[(size_t)s * vs + i] = consts[j];
  }
}

// V[slots[r]][i] = in[r][i]
__global__ void __launch_bounds__(GS_TPB)
k_gs_set_inputs(size_t n, const u32* __restrict__ slots, const u64* __restrict__ in, size_t is, u64* __restrict__ V, size_t vs) {
  const size_t i = (size_t)blockIdx.x * GS_TPB + threadIdx.x;
  if (i >= n) return;
  V[(size_t)slots[blockIdx.y] * vs + i] = in[(size_t)blockIdx.y * is + i];
}
// out[r][i] = V[slots[r]][i]
__global__ void __launch_bounds__(GS_TPB)
k_gs_get_outputs(size_t n, const u32* __restrict__ slots, const u64* __restrict__ V, size_t vs, u64* __restrict__ out, size_t os) {
  const size_t i = (size_t)blockIdx.x * GS_TPB + threadIdx.x;
  if (i >= n) return;
  out[(size_t)blockIdx.y * os + i] = V[(size_t)slots[blockIdx.y] * vs + i];
}

// ---- fused group kernels ----------------------------------------------------------------------------------
// One launch per group: thread (i, j) runs generator instance j of the group for circuit i, reading its operand
// rows straight from the value table through the group's slot lists and writing (or, for a slot that already has a
// producer, comparing) its outputs -- the same arithmetic cores as the standalone batched kernels (p2e_field.cuh),
// without the gather / scatter launches and their staging traffic.
struct GsIO {
  const u32* idx;    // [in_rows][m] operand slots
  const u32* oidx;   // [out_rows][m] output slots
  u64* V;
  size_t vs, n;
  u32* err;
  int m;
  __device__ __forceinline__ u64 in(int row, int j, size_t i) const {
    const u32 s = idx[(size_t)row * m + j];
    return s == GS_NONE ? 0 : V[(size_t)s * vs + i];
  }
  __device__ __forceinline__ u32 out(int row, int j, size_t i, u64 v) const {
    const u32 s = oidx[(size_t)row * m + j];
    if (s == GS_NONE) return 0;
    if (s & GS_CHECK) return V[(size_t)(s & ~GS_CHECK) * vs + i] != v ? (u32)P2E_ERR_CONNECT : 0u;
    V[(size_t)s * vs + i] = v;
    return 0;
  }
  __device__ __forceinline__ void fail(size_t i, u32 e) const {
    if (e) atomicOr(&err[i], e);
  }
};
template <class MOD>
__device__ __forceinline__ void gsb_mul(const GsIO& io, int j, size_t i) {
  u64 xw[NL], yw[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) {
    xw[l] = io.in(l, j, i);
    yw[l] = io.in(NL + l, j, i);
  }
  u32 xa[NL], ya[NL], r[NL], q[NL];
  u64 cs[2 * NL - 1], b[2 * NL - 2];
  u32 e = load_mul_wires(xw, xa) | load_mul_wires(yw, ya);
  if (e == 0) {
    // base field, factors below 2^256 (every canonical operand): the chain's fast core (p2e_f8.cuh, a quarter of the
    // instructions); it reports a quotient that does not fit 256 bits and the generic core takes over
    bool done = false;
    if (MOD::ID == 0 && (xa[NL - 1] >> 24) == 0 && (ya[NL - 1] >> 24) == 0) {
      i64 v[2 * NL - 1];
      u32 blo[2 * NL - 2], tops;
      if (mulw_p_fast(xa, ya, r, q, v, blo, tops) == 0) {
#pragma unroll
        for (int l = 0; l < 2 * NL - 1; l++) cs[l] = gl_from_i64(v[l]);
        done = true;
      }
    }
    if (!done) e = mul_witness_core<MOD>(xa, ya, r, q, cs, b);
  } else {
#pragma unroll
    for (int l = 0; l < NL; l++) r[l] = q[l] = 0;
#pragma unroll
    for (int l = 0; l < 2 * NL - 1; l++) cs[l] = 0;
  }
#pragma unroll
  for (int l = 0; l < NL; l++) {
    e |= io.out(P2E_MUL_OUT_R + l, j, i, r[l]);
    e |= io.out(P2E_MUL_OUT_Q + l, j, i, q[l]);
  }
#pragma unroll
  for (int l = 0; l < 2 * NL - 1; l++) e |= io.out(P2E_MUL_OUT_CHECK_SUM + l, j, i, cs[l]);
  io.fail(i, e);
}

__device__ __forceinline__ void gsb_check_sum(const GsIO& io, int j, size_t i) {
  u64 aw[2 * NL - 1], b[2 * NL - 2];
#pragma unroll
  for (int l = 0; l < 2 * NL - 1; l++) aw[l] = io.in(l, j, i);
  u32 e = check_sum_core(aw, b);
#pragma unroll
  for (int l = 0; l < 2 * NL - 2; l++) e |= io.out(l, j, i, b[l]);
  io.fail(i, e);
}

template <class MOD, bool IS_SUB>
__device__ __forceinline__ void gsb_addsub(const GsIO& io, int j, size_t i, int al, int bl) {
  u64 aw[NL], bw[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) {
    aw[l] = l < al ? io.in(l, j, i) : 0;
    bw[l] = l < bl ? io.in(al + l, j, i) : 0;
  }
  u32 ac[NL], bc[NL], r[NL];
  u32 e = load_canon<MOD>(aw, ac) | load_canon<MOD>(bw, bc);
  u32 ov = 0;
  if (e == 0) ov = IS_SUB ? sub_witness_core<MOD>(ac, bc, r) : add_witness_core<MOD>(ac, bc, r);
  else set_zero9(r);
#pragma unroll
  for (int l = 0; l < NL; l++) e |= io.out(l, j, i, r[l]);
  e |= io.out(NL, j, i, ov);
  io.fail(i, e);
}

template <class MOD>
__device__ __forceinline__ void gsb_add_many(const GsIO& io, int j, size_t i, int k) {
  u32 acc[NL];
  set_zero9(acc);
  u32 ov = 0, e = 0;
  for (int t = 0; t < k; t++) {
    u64 w[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) w[l] = io.in(t * NL + l, j, i);
    u32 v[NL], sum[NL];
    e |= load_canon<MOD>(w, v);
    add9(acc, v, sum);
    if (geq_mod<MOD>(sum)) { sub_mod9<MOD>(sum, acc); ov++; }
    else copy9(sum, acc);
  }
  if (e) { set_zero9(acc); ov = 0; }
#pragma unroll
  for (int l = 0; l < NL; l++) e |= io.out(l, j, i, acc[l]);
  e |= io.out(NL, j, i, ov);
  io.fail(i, e);
}

// one inversion per instance (the reference runs a Fermat ladder, nonnative.rs:863); a group is small, its instances
// run side by side
template <class MOD>
__device__ __forceinline__ void gsb_inv(const GsIO& io, int j, size_t i) {
  u64 w[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) w[l] = io.in(l, j, i);
  u32 xc[NL], inv[NL], div[NL];
  set_zero9(inv);
  set_zero9(div);
  u32 e = load_canon<MOD>(w, xc);
  if (e == 0 && is_zero9(xc)) e = ERR_INVERSE_ZERO;
  if (e == 0) {
    inv_mod<MOD>(xc, inv);     // division steps (p2e_modinv.cuh); the inverse is unique
    e |= inv_div_core<MOD>(xc, inv, div);
  }
#pragma unroll
  for (int l = 0; l < NL; l++) {
    e |= io.out(l, j, i, div[l]);
    e |= io.out(NL + l, j, i, inv[l]);
  }
  io.fail(i, e);
}

__device__ __forceinline__ void gsb_glv(const GsIO& io, int j, size_t i) {
  u64 w[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) w[l] = io.in(l, j, i);
  u32 kc[NL], k1[NL], k2[NL], n1 = 0, n2 = 0;
  u32 e = load_canon<ModN>(w, kc);
  if (e == 0) {
    glv_decompose_core(kc, k1, k2, n1, n2);
    u32 hi = 0;
#pragma unroll
    for (int l = 5; l < NL; l++) hi |= k1[l] | k2[l];
    if (hi) e |= ERR_LIMBS_OVERFLOW;
  } else {
    set_zero9(k1);
    set_zero9(k2);
  }
#pragma unroll
  for (int l = 0; l < 5; l++) {
    e |= io.out(l, j, i, k1[l]);
    e |= io.out(5 + l, j, i, k2[l]);
  }
  e |= io.out(10, j, i, n1);
  e |= io.out(11, j, i, n2);
  io.fail(i, e);
}

// ---- plonky2 glue generators (out of tree, restated) -----------------------------------
// BaseSplitGenerator<2> { row, num_limbs }: BaseSumGate<2> wire 0 = sum -> wires 1 .. num_limbs = its bits (LE).
// plonky2 asserts that the integer is fully consumed: a sum that does not fit sets P2E_ERR_LIMB_RANGE.
__device__ __forceinline__ void gsb_base_split(const GsIO& io, int j, size_t i, int nl) {
  const u64 v = io.in(0, j, i);
  u32 e = (nl < 64 && (v >> nl)) ? (u32)P2E_ERR_LIMB_RANGE : 0u;
  for (int r = 0; r < nl; r++) e |= io.out(r, j, i, (v >> r) & 1ull);
  io.fail(i, e);
}

// RandomAccessGenerator of one copy of a RandomAccessGate { bits }: claimed_element = list[access_index] and the
// little-endian bits of access_index.  An index outside the list is a panic upstream: P2E_ERR_LIMB_RANGE.
__device__ __forceinline__ void gsb_random_access(const GsIO& io, int j, size_t i, int bits) {
  const u64 idx = io.in(0, j, i);
  const u64 vs = 1ull << bits;
  u32 e = idx >= vs ? (u32)P2E_ERR_LIMB_RANGE : 0u;
  const int k = (int)(idx & (vs - 1));
  e |= io.out(0, j, i, io.in(1 + k, j, i));
  for (int b = 0; b < bits; b++) e |= io.out(1 + b, j, i, (idx >> b) & 1ull);
  io.fail(i, e);
}

__device__ u64 gs_inv(u64 x) {       // x^(P-2)
  u64 r = 1, b = x;
  const u64 e = GS_P - 2;
  for (int k = 0; k < 64; k++) {
    if ((e >> k) & 1) r = gs_mul(r, b);
    b = gs_mul(b, b);
  }
  return r;
}
// EqualityGenerator { x, y, equal, inv }: equal = [x == y], inv = (x - y)^-1 or 0
__device__ __forceinline__ void gsb_equality(const GsIO& io, int j, size_t i) {
  const u64 x = io.in(0, j, i), y = io.in(1, j, i);
  u64 d = x - y;
  if (x < y) d += GS_P;
  u32 e = io.out(0, j, i, x == y ? 1ull : 0ull);
  e |= io.out(1, j, i, x == y ? 0ull : gs_inv(d));
  io.fail(i, e);
}

// the per-group launches of the bodies above: thread (i, j0 + blockIdx.y)
#define GS_THREAD()                                              \
  const size_t i = (size_t)blockIdx.x * GS_TPB + threadIdx.x;    \
  const int j = j0 + blockIdx.y;                                 \
  if (i >= io.n) return
template <class MOD>
__global__ void __launch_bounds__(GS_TPB) k_gs_mul(GsIO io, int j0) { GS_THREAD(); gsb_mul<MOD>(io, j, i); }
__global__ void __launch_bounds__(GS_TPB) k_gs_check_sum(GsIO io, int j0) { GS_THREAD(); gsb_check_sum(io, j, i); }
template <class MOD, bool IS_SUB>
__global__ void __launch_bounds__(GS_TPB) k_gs_addsub(GsIO io, int j0, int al, int bl) { GS_THREAD(); gsb_addsub<MOD, IS_SUB>(io, j, i, al, bl); }
template <class MOD>
__global__ void __launch_bounds__(GS_TPB) k_gs_add_many(GsIO io, int j0, int k) { GS_THREAD(); gsb_add_many<MOD>(io, j, i, k); }
template <class MOD>
__global__ void __launch_bounds__(GS_TPB) k_gs_inv(GsIO io, int j0) { GS_THREAD(); gsb_inv<MOD>(io, j, i); }
__global__ void __launch_bounds__(GS_TPB) k_gs_glv(GsIO io, int j0) { GS_THREAD(); gsb_glv(io, j, i); }
__global__ void __launch_bounds__(GS_TPB) k_gs_base_split(GsIO io, int j0, int nl) { GS_THREAD(); gsb_base_split(io, j, i, nl); }
__global__ void __launch_bounds__(GS_TPB) k_gs_random_access(GsIO io, int j0, int bits) { GS_THREAD(); gsb_random_access(io, j, i, bits); }
__global__ void __launch_bounds__(GS_TPB) k_gs_equality(GsIO io, int j0) { GS_THREAD(); gsb_equality(io, j, i); }

// ---- the level walker -------------------------------------------------------------------------------------
// Circuits are independent, so a deep schedule needs no grid-wide ordering at all: a CTA owns 32 circuits (the lanes
// of its warps) and walks ALL levels itself -- warp w runs generator w, w + W, ... of the level for those 32 circuits,
// __syncthreads() closes the level.  One launch for the whole stream; a level costs the latency of its slowest
// generator core instead of a kernel boundary per group (the ECDSA circuit: 1 862 levels, 4 397 groups).
struct GsStep {
  const u32* idx;
  const u32* oidx;
  const u64* consts;
  int m, kind, modulus, shape_a, shape_b, item0;   // item0: generators of the level before this group (warp rotation)
};

__device__ __forceinline__ void gsb_arith(const GsStep& s, u64* V, size_t vs, u32* err, int j, size_t i) {
  const int m = s.m;
  const u64 a = V[(size_t)s.idx[j] * vs + i], b = V[(size_t)s.idx[m + j] * vs + i], c = V[(size_t)s.idx[2 * m + j] * vs + i];
  const u64 v = gs_add(gs_mul(s.consts[j], gs_mul(a, b)), gs_mul(s.consts[m + j], c));
  const u32 o = s.idx[3 * m + j];
  if (o & GS_CHECK) {
    if (V[(size_t)(o & ~GS_CHECK) * vs + i] != v) atomicOr(&err[i], (u32)P2E_ERR_CONNECT);
  } else {
    V[(size_t)o * vs + i] = v;
  }
}
__device__ __forceinline__ void gsb_const(const GsStep& s, u64* V, size_t vs, u32* err, int j, size_t i) {
  const u32 o = s.oidx[j];
  const u64 v = s.consts[j];
  if (o & GS_CHECK) {
    if (V[(size_t)(o & ~GS_CHECK) * vs + i] != v) atomicOr(&err[i], (u32)P2E_ERR_CONNECT);
  } else {
    V[(size_t)o * vs + i] = v;
  }
}

// out-of-line copies for the walker: one body resident per warp at a time, the interpreter itself stays small
template <class MOD> __device__ __noinline__ void gsn_mul(GsIO io, int j, size_t i) { gsb_mul<MOD>(io, j, i); }
template <class MOD> __device__ __noinline__ void gsn_inv(GsIO io, int j, size_t i) { gsb_inv<MOD>(io, j, i); }
template <class MOD> __device__ __noinline__ void gsn_add(GsIO io, int j, size_t i, int al, int bl) { gsb_addsub<MOD, false>(io, j, i, al, bl); }
template <class MOD> __device__ __noinline__ void gsn_sub(GsIO io, int j, size_t i, int al, int bl) { gsb_addsub<MOD, true>(io, j, i, al, bl); }
template <class MOD> __device__ __noinline__ void gsn_add_many(GsIO io, int j, size_t i, int k) { gsb_add_many<MOD>(io, j, i, k); }
__device__ __noinline__ void gsn_check_sum(GsIO io, int j, size_t i) { gsb_check_sum(io, j, i); }
__device__ __noinline__ void gsn_glv(GsIO io, int j, size_t i) { gsb_glv(io, j, i); }
__device__ __noinline__ void gsn_equality(GsIO io, int j, size_t i) { gsb_equality(io, j, i); }

__device__ __forceinline__ void gs_exec(const GsStep& s, u64* V, size_t vs, size_t n, u32* err, int j, size_t i) {
  const GsIO io = {s.idx, s.oidx, V, vs, n, err, s.m};
  const bool base = s.modulus == P2E_SECP256K1_BASE;
  switch (s.kind) {
    case K_MUL: if (base) gsn_mul<ModP>(io, j, i); else gsn_mul<ModN>(io, j, i); break;
    case K_CHECKSUM: gsn_check_sum(io, j, i); break;
    case K_ADD: if (base) gsn_add<ModP>(io, j, i, s.shape_a, s.shape_b); else gsn_add<ModN>(io, j, i, s.shape_a, s.shape_b); break;
    case K_SUB: if (base) gsn_sub<ModP>(io, j, i, s.shape_a, s.shape_b); else gsn_sub<ModN>(io, j, i, s.shape_a, s.shape_b); break;
    case K_MULTIADD: if (base) gsn_add_many<ModP>(io, j, i, s.shape_a); else gsn_add_many<ModN>(io, j, i, s.shape_a); break;
    case K_INV: if (base) gsn_inv<ModP>(io, j, i); else gsn_inv<ModN>(io, j, i); break;
    case K_GLV: gsn_glv(io, j, i); break;
    case K_ARITH: gsb_arith(s, V, vs, err, j, i); break;
    case K_CONST: gsb_const(s, V, vs, err, j, i); break;
    case K_BASESPLIT: gsb_base_split(io, j, i, s.shape_a); break;
    case K_RANDACC: gsb_random_access(io, j, i, s.shape_a); break;
    case K_EQUAL: gsn_equality(io, j, i); break;
    default: break;                       // K_DIVREM: the host does not choose the walker for such a stream
  }
}

// W warps per CTA, C circuits per CTA (a power of two <= 32): a warp runs 32 / C generators of one group side by side
// (lane = sub * C + circuit), so a small batch still fills the machine with CTAs -- two to four of them per SM overlap
// each other's long levels (an inverse is ~20 us of one warp) -- without idle lanes in the wide levels.
template <int W, int C>
__global__ void __launch_bounds__(32 * W) k_gs_walk(const GsStep* __restrict__ steps, const int* __restrict__ level_first, int n_levels,
                                                    u64* V, size_t vs, size_t n, u32* err) {
  constexpr int SUBS = 32 / C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / C;
  const size_t i = (size_t)blockIdx.x * C + (lane % C);
  const bool active = i < n;
  int s0 = level_first[0];
  for (int L = 0; L < n_levels; L++) {
    const int s1 = level_first[L + 1];
    for (int k = s0; k < s1; k++) {
      const GsStep s = steps[k];
      int w = warp - s.item0 % W;
      if (w < 0) w += W;
      if (active)
        for (int j = w * SUBS + sub; j < s.m; j += W * SUBS) gs_exec(s, V, vs, n, err, j, i);
    }
    s0 = s1;
    __syncthreads();
  }
}

template <int W, int C>
static void launch_walk(const GsStep* steps, const int* level_first, int n_levels, u64* V, size_t vs, size_t n, u32* err, cudaStream_t st) {
  k_gs_walk<W, C><<<(unsigned)((n + C - 1) / C), 32 * W, 0, st>>>(steps, level_first, n_levels, V, vs, n, err);
}

// ---- host: stream parsing ---------------------------------------------------------------------------------
struct TargetKey {
  u64 is_wire, a, b;
  bool operator<(const TargetKey& o) const {
    if (is_wire != o.is_wire) return is_wire < o.is_wire;
    if (a != o.a) return a < o.a;
    return b < o.b;
  }
};

struct Reader {
  const uint8_t* p;
  size_t len, pos = 0;
  bool ok = true;
  u64 u64le() {
    if (pos + 8 > len) { ok = false; return 0; }
    u64 v = 0;
    for (int k = 7; k >= 0; k--) v = (v << 8) | p[pos + k];
    pos += 8;
    return v;
  }
  u32 u32le() {
    if (pos + 4 > len) { ok = false; return 0; }
    u32 v = 0;
    for (int k = 3; k >= 0; k--) v = (v << 8) | p[pos + k];
    pos += 4;
    return v;
  }
  int u8() {
    if (pos + 1 > len) { ok = false; return 0; }
    return p[pos++];
  }
};

struct GsGen {
  int kind, modulus;
  int shape_a = 0, shape_b = 0;          // limb counts (ADD/SUB/DIVREM), summand count (MULTIADD)
  std::vector<u32> ins, outs;            // target ids (before partitioning), GS_NONE = absent
  u64 c0 = 0, c1 = 0;
  int level = -1;
};

struct GsGroup {
  int level, kind, modulus, shape_a, shape_b, m;
  int in_rows = 0, out_rows = 0;
  size_t idx_off = 0, oidx_off = 0, const_off = 0;   // into the device index / constant pools
};

struct GenStream {
  std::vector<TargetKey> targets;        // id -> key
  std::map<TargetKey, u32> target_id;
  std::vector<u32> parent;               // union-find over target ids
  std::vector<u32> slot_of_target;
  std::vector<GsGen> gens;
  std::vector<GsGroup> groups;
  std::vector<u32> input_slots;          // slots read but never produced, and the declared inputs
  std::vector<u32> ext_targets;          // P2E_GEN_INPUT: targets the host sets before generation (PartialWitness)
  u32 n_slots = 0;
  int n_levels = 0;
  // device side (created lazily on the bound device)
  int device = -1;
  u32* d_idx = nullptr;
  u64* d_consts = nullptr;
  GsStep* d_steps = nullptr;             // the walker's program: one step per group, level by level
  int* d_level_first = nullptr;
  bool walkable = false;                 // no group kind the walker lacks (BigUint div_rem)
  std::vector<u32> h_idx;
  std::vector<u64> h_consts;
  // grow-only staging
  u64* d_stage = nullptr;
  size_t stage_bytes = 0;
  u32* d_stage_err = nullptr;
  size_t stage_err_bytes = 0;
  u32* d_io = nullptr;
  size_t io_bytes = 0;
  // CUDA graph of the last launch sequence (deep circuits are launch-bound: 3 launches per group); valid for
  // the same (n, table, stride, err) and staging buffers
  cudaGraphExec_t graph = nullptr;
  size_t g_n = 0, g_vs = 0;
  const void *g_V = nullptr, *g_err = nullptr, *g_stage = nullptr, *g_stage_err = nullptr;
  bool g_fused = false;
  std::string error;

  u32 tid(const TargetKey& k) {
    auto it = target_id.find(k);
    if (it != target_id.end()) return it->second;
    u32 id = (u32)targets.size();
    targets.push_back(k);
    parent.push_back(id);
    target_id.emplace(k, id);
    return id;
  }
  u32 find(u32 x) {
    while (parent[x] != x) {
      parent[x] = parent[parent[x]];
      x = parent[x];
    }
    return x;
  }
  void drop_graph() {
    if (graph) cudaGraphExecDestroy(graph);
    graph = nullptr;
  }
  void release_device() {
    drop_graph();
    if (d_idx) cudaFree(d_idx);
    if (d_consts) cudaFree(d_consts);
    if (d_steps) cudaFree(d_steps);
    if (d_level_first) cudaFree(d_level_first);
    d_steps = nullptr; d_level_first = nullptr;
    if (d_stage) cudaFree(d_stage);
    if (d_stage_err) cudaFree(d_stage_err);
    if (d_io) cudaFree(d_io);
    d_idx = nullptr; d_consts = nullptr; d_stage = nullptr; d_stage_err = nullptr; d_io = nullptr;
    stage_bytes = stage_err_bytes = io_bytes = 0;
    device = -1;
  }
};

static std::mutex g_gs_mu;
static std::vector<GenStream*> g_streams;

static u32 read_target(GenStream& S, Reader& R) {
  const int w = R.u8();
  if (w) {
    const u64 row = R.u64le(), col = R.u64le();
    return S.tid({1, row, col});
  }
  return S.tid({0, R.u64le(), 0});
}
static bool read_target_vec(GenStream& S, Reader& R, std::vector<u32>& out, size_t max_len) {
  const u64 n = R.u64le();
  if (!R.ok || n > max_len) return false;
  for (u64 k = 0; k < n; k++) out.push_back(read_target(S, R));
  return R.ok;
}
static void pad_to(std::vector<u32>& v, size_t n) {
  while (v.size() < n) v.push_back(GS_NONE);
}
static u32 wire(GenStream& S, u64 row, u64 col) { return S.tid({1, row, col}); }

static int rows_of(const GsGen& g, int* in_rows, int* out_rows) {
  switch (g.kind) {
    case K_MUL: *in_rows = 18; *out_rows = P2E_MUL_OUT_ROWS; return 0;
    case K_CHECKSUM: *in_rows = 17; *out_rows = 16; return 0;
    case K_ADD: case K_SUB: *in_rows = g.shape_a + g.shape_b; *out_rows = 10; return 0;
    case K_MULTIADD: *in_rows = 9 * g.shape_a; *out_rows = 10; return 0;
    case K_INV: *in_rows = 9; *out_rows = 18; return 0;
    case K_DIVREM: {
      const int dl = (g.shape_b > g.shape_a + 1) ? 0 : g.shape_a - g.shape_b + 1;
      *in_rows = g.shape_a + g.shape_b; *out_rows = dl + g.shape_b; return 0;
    }
    case K_GLV: *in_rows = 9; *out_rows = 12; return 0;
    case K_ARITH: *in_rows = 3; *out_rows = 1; return 0;
    case K_CONST: *in_rows = 0; *out_rows = 1; return 0;
    case K_BASESPLIT: *in_rows = 1; *out_rows = g.shape_a; return 0;
    case K_RANDACC: *in_rows = 1 + (1 << g.shape_a); *out_rows = 1 + g.shape_a; return 0;
    case K_EQUAL: *in_rows = 2; *out_rows = 2; return 0;
  }
  return -1;
}

// parses one record; returns false (with S.error set) on a malformed stream
static bool parse_record(GenStream& S, Reader& R) {
  const u32 tag = R.u32le();
  if (!R.ok) { S.error = "genstream: truncated record tag"; return false; }
  GsGen g;
  g.modulus = P2E_SECP256K1_BASE;
  auto bad = [&](const char* m) { S.error = std::string("genstream: ") + m; return false; };
  switch (tag) {
    case P2E_GEN_MUL_NONNATIVE_SCALAR: g.modulus = P2E_SECP256K1_SCALAR; [[fallthrough]];
    case P2E_GEN_MUL_NONNATIVE_BASE: {
      // MulNonnativeGenerator { row, num_limbs }: wires of the gate row, gates/mul_nonnative.rs:41-59
      const u64 row = R.u64le(), nl = R.u64le();
      if (!R.ok) return bad("truncated MulNonnativeGenerator");
      if (nl != 9) return bad("MulNonnativeGenerator: num_limbs must be 9");
      g.kind = K_MUL;
      for (int c = 0; c < 18; c++) g.ins.push_back(wire(S, row, c));
      for (int c = 18; c < 53; c++) g.outs.push_back(wire(S, row, c));       // r, q, check_sum
      pad_to(g.outs, P2E_MUL_OUT_ROWS);                                        // the carries belong to CheckSumGenerator
      break;
    }
    case P2E_GEN_CHECK_SUM: {
      // CheckSumGenerator { row, num_limbs }: a = wires 0..16, b = wires 17..32, gates/mul_nonnative.rs:384-390
      const u64 row = R.u64le(), nl = R.u64le();
      if (!R.ok) return bad("truncated CheckSumGenerator");
      if (nl != 9) return bad("CheckSumGenerator: num_limbs must be 9");
      g.kind = K_CHECKSUM;
      for (int c = 0; c < 17; c++) g.ins.push_back(wire(S, row, c));
      for (int c = 17; c < 33; c++) g.outs.push_back(wire(S, row, c));
      break;
    }
    case P2E_GEN_NONNATIVE_ADDITION_SCALAR: case P2E_GEN_NONNATIVE_SUBTRACTION_SCALAR: g.modulus = P2E_SECP256K1_SCALAR; [[fallthrough]];
    case P2E_GEN_NONNATIVE_ADDITION_BASE: case P2E_GEN_NONNATIVE_SUBTRACTION_BASE: {
      // { a, b, sum | diff, overflow }: gadgets/nonnative.rs:647-652, 812-817
      g.kind = (tag == P2E_GEN_NONNATIVE_ADDITION_BASE || tag == P2E_GEN_NONNATIVE_ADDITION_SCALAR) ? K_ADD : K_SUB;
      std::vector<u32> a, b, s;
      if (!read_target_vec(S, R, a, 9) || !read_target_vec(S, R, b, 9) || !read_target_vec(S, R, s, 64)) return bad("bad NonNativeAddition/SubtractionGenerator");
      const u32 ov = read_target(S, R);
      if (!R.ok) return bad("truncated NonNativeAddition/SubtractionGenerator");
      if (s.size() != 9) return bad("sum / diff must have 9 limbs");
      g.shape_a = (int)a.size(); g.shape_b = (int)b.size();
      g.ins = a; g.ins.insert(g.ins.end(), b.begin(), b.end());
      g.outs = s; g.outs.push_back(ov);
      break;
    }
    case P2E_GEN_NONNATIVE_MULTIPLE_ADDS_SCALAR: g.modulus = P2E_SECP256K1_SCALAR; [[fallthrough]];
    case P2E_GEN_NONNATIVE_MULTIPLE_ADDS_BASE: {
      // { summands_len, summands.., sum, overflow (UXTarget) }: gadgets/nonnative.rs:730-740
      const u64 k = R.u64le();
      if (!R.ok || k < 2 || k > 64) return bad("NonNativeMultipleAddsGenerator: summand count must be 2..64");
      g.kind = K_MULTIADD;
      g.shape_a = (int)k;
      for (u64 t = 0; t < k; t++) {
        std::vector<u32> sm;
        if (!read_target_vec(S, R, sm, 9)) return bad("bad summand");
        pad_to(sm, 9);
        g.ins.insert(g.ins.end(), sm.begin(), sm.end());
      }
      std::vector<u32> s;
      if (!read_target_vec(S, R, s, 64) || s.size() != 9) return bad("sum must have 9 limbs");
      const u32 ov = read_target(S, R);
      if (!R.ok) return bad("truncated NonNativeMultipleAddsGenerator");
      g.outs = s; g.outs.push_back(ov);
      break;
    }
    case P2E_GEN_NONNATIVE_INVERSE_SCALAR: g.modulus = P2E_SECP256K1_SCALAR; [[fallthrough]];
    case P2E_GEN_NONNATIVE_INVERSE_BASE: {
      // { x, inv, div }: gadgets/nonnative.rs:874-878; the kernel's rows are div[9], inv[9] (set_* order :870-871)
      std::vector<u32> x, inv, div;
      if (!read_target_vec(S, R, x, 9) || !read_target_vec(S, R, inv, 9) || !read_target_vec(S, R, div, 9)) return bad("bad NonNativeInverseGenerator");
      if (inv.size() != 9 || div.size() != 9) return bad("inv / div must have 9 limbs");
      g.kind = K_INV;
      pad_to(x, 9);
      g.ins = x;
      g.outs = div; g.outs.insert(g.outs.end(), inv.begin(), inv.end());
      break;
    }
    case P2E_GEN_BIGUINT_DIV_REM: {
      // { a, b, div, rem }: gadgets/biguint.rs:520-526; output sizes :391-399
      std::vector<u32> a, b, d, r;
      if (!read_target_vec(S, R, a, 18) || !read_target_vec(S, R, b, 18) || !read_target_vec(S, R, d, 64) || !read_target_vec(S, R, r, 64)) return bad("bad BigUintDivRemGenerator (operands have at most 18 limbs)");
      if (b.empty()) return bad("BigUintDivRemGenerator: b has no limbs");
      g.kind = K_DIVREM;
      g.shape_a = (int)a.size(); g.shape_b = (int)b.size();
      const size_t dl = (b.size() > a.size() + 1) ? 0 : a.size() - b.size() + 1;
      if (d.size() != dl || r.size() != b.size()) return bad("BigUintDivRemGenerator: div / rem limb counts do not follow biguint.rs:391-399");
      g.ins = a; g.ins.insert(g.ins.end(), b.begin(), b.end());
      g.outs = d; g.outs.insert(g.outs.end(), r.begin(), r.end());
      break;
    }
    case P2E_GEN_GLV_DECOMPOSITION: {
      // { k, k1, k2, k1_neg, k2_neg }: gadgets/glv.rs:144-150
      std::vector<u32> k, k1, k2;
      if (!read_target_vec(S, R, k, 9) || !read_target_vec(S, R, k1, 64) || !read_target_vec(S, R, k2, 64)) return bad("bad GLVDecompositionGenerator");
      const u32 n1 = read_target(S, R), n2 = read_target(S, R);
      if (!R.ok) return bad("truncated GLVDecompositionGenerator");
      if (k1.size() != 5 || k2.size() != 5) return bad("k1 / k2 must have 5 limbs (glv.rs:62-63)");
      g.kind = K_GLV;
      g.modulus = P2E_SECP256K1_SCALAR;
      pad_to(k, 9);
      g.ins = k;
      g.outs = k1; g.outs.insert(g.outs.end(), k2.begin(), k2.end());
      g.outs.push_back(n1); g.outs.push_back(n2);
      break;
    }
    case P2E_GEN_COPY: {
      const u32 a = read_target(S, R), b = read_target(S, R);
      if (!R.ok) return bad("truncated copy record");
      const u32 ra = S.find(a), rb = S.find(b);
      if (ra != rb) S.parent[std::max(ra, rb)] = std::min(ra, rb);
      return true;
    }
    case P2E_GEN_INPUT: {
      // a target of the host's PartialWitness: set before generation; a generator that also writes its partition (a
      // connect_* against a public input, gadgets/ecdsa.rs:48-52) is checked against it instead of producing it
      const u32 t = read_target(S, R);
      if (!R.ok) return bad("truncated input record");
      S.ext_targets.push_back(t);
      return true;
    }
    case P2E_GEN_CONSTANT: {
      // plonky2 ConstantGenerator { row, constant_index, wire_index, constant }
      const u64 row = R.u64le();
      (void)R.u64le();
      const u64 wi = R.u64le(), c = R.u64le();
      if (!R.ok) return bad("truncated ConstantGenerator");
      if (c >= GS_P) return bad("ConstantGenerator: non-canonical field element");
      g.kind = K_CONST;
      g.outs.push_back(wire(S, row, wi));
      g.c0 = c;
      break;
    }
    case P2E_GEN_ARITHMETIC: {
      // plonky2 ArithmeticBaseGenerator { row, const_0, const_1, i }: wires 4i, 4i+1, 4i+2 -> 4i+3
      const u64 row = R.u64le(), c0 = R.u64le(), c1 = R.u64le(), i = R.u64le();
      if (!R.ok) return bad("truncated ArithmeticBaseGenerator");
      if (c0 >= GS_P || c1 >= GS_P) return bad("ArithmeticBaseGenerator: non-canonical field element");
      g.kind = K_ARITH;
      for (int c = 0; c < 3; c++) g.ins.push_back(wire(S, row, 4 * i + c));
      g.outs.push_back(wire(S, row, 4 * i + 3));
      g.c0 = c0; g.c1 = c1;
      break;
    }
    case P2E_GEN_BASE_SPLIT: {
      // plonky2 BaseSplitGenerator<2> { row, num_limbs } (restated): BaseSumGate::WIRE_SUM = 0, limbs 1 .. num_limbs
      const u64 row = R.u64le(), nl = R.u64le();
      if (!R.ok) return bad("truncated BaseSplitGenerator");
      if (nl < 1 || nl > 63) return bad("BaseSplitGenerator: num_limbs must be 1..63");
      g.kind = K_BASESPLIT;
      g.shape_a = (int)nl;
      g.ins.push_back(wire(S, row, 0));
      for (u64 c = 0; c < nl; c++) g.outs.push_back(wire(S, row, 1 + c));
      break;
    }
    case P2E_GEN_RANDOM_ACCESS: {
      // plonky2 RandomAccessGenerator { row, gate { bits, num_copies, num_extra_constants }, copy } (restated); wire map
      // of RandomAccessGate: access_index (2 + 2^bits) copy, claimed_element + 1, list items + 2 .., bits after the
      // routed wires: (2 + 2^bits) num_copies + num_extra_constants + bits copy + i
      const u64 row = R.u64le(), bits = R.u64le(), copies = R.u64le(), extra = R.u64le(), copy = R.u64le();
      if (!R.ok) return bad("truncated RandomAccessGenerator");
      if (bits < 1 || bits > 6 || copy >= copies || copies > 64 || extra > 64) return bad("RandomAccessGenerator: bad gate parameters");
      const u64 vs = 1ull << bits, base = (2 + vs) * copy, bit0 = (2 + vs) * copies + extra + bits * copy;
      g.kind = K_RANDACC;
      g.shape_a = (int)bits;
      g.ins.push_back(wire(S, row, base));
      for (u64 k = 0; k < vs; k++) g.ins.push_back(wire(S, row, base + 2 + k));
      g.outs.push_back(wire(S, row, base + 1));
      for (u64 b = 0; b < bits; b++) g.outs.push_back(wire(S, row, bit0 + b));
      break;
    }
    case P2E_GEN_EQUALITY: {
      // plonky2 EqualityGenerator { x, y, equal (BoolTarget), inv } (restated)
      const u32 x = read_target(S, R), y = read_target(S, R), eq = read_target(S, R), inv = read_target(S, R);
      if (!R.ok) return bad("truncated EqualityGenerator");
      g.kind = K_EQUAL;
      g.ins = {x, y};
      g.outs = {eq, inv};
      break;
    }
    default:
      return bad("unknown generator tag");
  }
  S.gens.push_back(std::move(g));
  return true;
}

// partitions -> slots, producers, levels, groups
static bool finalize(GenStream& S) {
  const u32 nt = (u32)S.targets.size();
  S.slot_of_target.assign(nt, GS_NONE);
  std::vector<u32> slot_of_root(nt, GS_NONE);
  S.n_slots = 0;
  for (u32 t = 0; t < nt; t++) {
    const u32 r = S.find(t);
    if (slot_of_root[r] == GS_NONE) slot_of_root[r] = S.n_slots++;
    S.slot_of_target[t] = slot_of_root[r];
  }
  auto slot = [&](u32 t) { return t == GS_NONE ? GS_NONE : S.slot_of_target[t]; };
  const int ng = (int)S.gens.size();
  std::vector<int> producer(S.n_slots, -1);
  std::vector<char> is_ext(S.n_slots, 0);
  for (u32 t : S.ext_targets) is_ext[slot(t)] = 1;
  for (int gi = 0; gi < ng; gi++)
    for (u32 t : S.gens[gi].outs)
      if (t != GS_NONE && !is_ext[slot(t)] && producer[slot(t)] < 0) producer[slot(t)] = gi;
  // levels: iterative DFS over the producer relation (a second writer of a slot also waits for the first)
  std::vector<int> state(ng, 0);
  for (int root = 0; root < ng; root++) {
    if (state[root] == 2) continue;
    std::vector<std::pair<int, size_t>> stack;
    stack.push_back({root, 0});
    state[root] = 1;
    while (!stack.empty()) {
      int gi = stack.back().first;
      GsGen& g = S.gens[gi];
      const size_t n_dep = g.ins.size() + g.outs.size();
      bool descended = false;
      while (stack.back().second < n_dep) {
        const size_t k = stack.back().second++;
        const u32 t = k < g.ins.size() ? g.ins[k] : g.outs[k - g.ins.size()];
        if (t == GS_NONE) continue;
        const int p = producer[slot(t)];
        if (p < 0 || p == gi) continue;
        if (state[p] == 1) {
          S.error = "genstream: the generators' dependencies form a cycle (generator #" + std::to_string(gi) + " of kind " +
                    std::to_string(g.kind) + " waits for generator #" + std::to_string(p) + " of kind " + std::to_string(S.gens[p].kind) + "; chain:";
          for (size_t q = stack.size() > 10 ? stack.size() - 10 : 0; q < stack.size(); q++)
            S.error += " #" + std::to_string(stack[q].first) + "/k" + std::to_string(S.gens[stack[q].first].kind);
          S.error += ")";
          return false;
        }
        if (state[p] == 0) {
          state[p] = 1;
          stack.push_back({p, 0});
          descended = true;
          break;
        }
      }
      if (descended) continue;
      int lv = 0;
      for (size_t k = 0; k < n_dep; k++) {
        const u32 t = k < g.ins.size() ? g.ins[k] : g.outs[k - g.ins.size()];
        if (t == GS_NONE) continue;
        const int p = producer[slot(t)];
        if (p >= 0 && p != gi) lv = std::max(lv, S.gens[p].level + 1);
      }
      g.level = lv;
      state[gi] = 2;
      stack.pop_back();
    }
  }
  // inputs: slots read by a generator and produced by none
  std::vector<char> is_input(S.n_slots, 0);
  for (const GsGen& g : S.gens)
    for (u32 t : g.ins)
      if (t != GS_NONE && producer[slot(t)] < 0) is_input[slot(t)] = 1;
  for (u32 s = 0; s < S.n_slots; s++)
    if (is_input[s] || is_ext[s]) S.input_slots.push_back(s);
  // groups: (level, kind, modulus, shape) in stream order inside a group
  std::vector<int> order(ng);
  for (int i = 0; i < ng; i++) order[i] = i;
  auto key = [&](int i) {
    const GsGen& g = S.gens[i];
    return std::make_tuple(g.level, g.kind, g.modulus, g.shape_a, g.shape_b);
  };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
  S.n_levels = 0;
  for (int a = 0; a < ng;) {
    int b = a;
    while (b < ng && key(order[b]) == key(order[a])) b++;
    const GsGen& g0 = S.gens[order[a]];
    GsGroup G;
    G.level = g0.level; G.kind = g0.kind; G.modulus = g0.modulus; G.shape_a = g0.shape_a; G.shape_b = g0.shape_b;
    G.m = b - a;
    rows_of(g0, &G.in_rows, &G.out_rows);
    G.idx_off = S.h_idx.size();
    for (int r = 0; r < G.in_rows; r++)
      for (int j = a; j < b; j++) S.h_idx.push_back(slot(S.gens[order[j]].ins[r]));
    G.oidx_off = S.h_idx.size();
    for (int r = 0; r < G.out_rows; r++)
      for (int j = a; j < b; j++) {
        const int gi = order[j];
        const u32 t = S.gens[gi].outs[r];
        u32 s = slot(t);
        if (s != GS_NONE && producer[s] != gi) s |= GS_CHECK;
        else if (s != GS_NONE) {
          // the same generator may list one partition twice (outputs joined by a copy): write once, check the rest
          for (int r2 = 0; r2 < r; r2++)
            if (S.gens[gi].outs[r2] != GS_NONE && slot(S.gens[gi].outs[r2]) == s) { s |= GS_CHECK; break; }
        }
        S.h_idx.push_back(s);
      }
    G.const_off = S.h_consts.size();
    if (G.kind == K_ARITH || G.kind == K_CONST) {
      for (int j = a; j < b; j++) S.h_consts.push_back(S.gens[order[j]].c0);
      if (G.kind == K_ARITH)
        for (int j = a; j < b; j++) S.h_consts.push_back(S.gens[order[j]].c1);
    }
    S.groups.push_back(G);
    S.n_levels = std::max(S.n_levels, G.level + 1);
    a = b;
  }
  return true;
}

// p2e_shutdown(): device buffers and graphs of every loaded stream go with the context (the parsed streams stay)
void genstream_release_device(Context& c) {
  std::lock_guard<std::mutex> lock(g_gs_mu);
  for (GenStream* S : g_streams)
    if (S && S->device == c.device) S->release_device();
}

static GenStream* get_stream(int h) {
  if (h < 0 || h >= (int)g_streams.size() || !g_streams[h]) {
    set_error_msg("genstream: bad handle");
    return nullptr;
  }
  return g_streams[h];
}

static int grow(void** p, size_t* have, size_t need) {
  if (*have >= need) return 0;
  if (*p) P2E_CUDA(cudaFree(*p));
  *p = nullptr;
  *have = 0;
  P2E_CUDA(cudaMalloc(p, need));
  *have = need;
  return 0;
}

static int ensure_device(GenStream& S) {
  Context& c = ctx();
  if (S.device == c.device && S.d_idx) return 0;
  S.release_device();
  P2E_CUDA(cudaMalloc((void**)&S.d_idx, std::max<size_t>(S.h_idx.size(), 1) * 4));
  P2E_CUDA(cudaMemcpy(S.d_idx, S.h_idx.data(), S.h_idx.size() * 4, cudaMemcpyHostToDevice));
  P2E_CUDA(cudaMalloc((void**)&S.d_consts, std::max<size_t>(S.h_consts.size(), 1) * 8));
  P2E_CUDA(cudaMemcpy(S.d_consts, S.h_consts.data(), S.h_consts.size() * 8, cudaMemcpyHostToDevice));
  // the walker's program (groups are sorted by level)
  std::vector<GsStep> steps;
  std::vector<int> first(S.n_levels + 1, 0);
  S.walkable = true;
  int level = -1, item0 = 0;
  for (const GsGroup& G : S.groups) {
    if (G.kind == K_DIVREM) S.walkable = false;
    while (level < G.level) {
      first[++level] = (int)steps.size();
      item0 = 0;
    }
    GsStep st;
    st.idx = S.d_idx + G.idx_off;
    st.oidx = S.d_idx + G.oidx_off;
    st.consts = S.d_consts + G.const_off;
    st.m = G.m; st.kind = G.kind; st.modulus = G.modulus; st.shape_a = G.shape_a; st.shape_b = G.shape_b; st.item0 = item0;
    item0 += G.m;
    steps.push_back(st);
  }
  while (level < S.n_levels) first[++level] = (int)steps.size();
  P2E_CUDA(cudaMalloc((void**)&S.d_steps, std::max<size_t>(steps.size(), 1) * sizeof(GsStep)));
  P2E_CUDA(cudaMemcpy(S.d_steps, steps.data(), steps.size() * sizeof(GsStep), cudaMemcpyHostToDevice));
  P2E_CUDA(cudaMalloc((void**)&S.d_level_first, first.size() * sizeof(int)));
  P2E_CUDA(cudaMemcpy(S.d_level_first, first.data(), first.size() * sizeof(int), cudaMemcpyHostToDevice));
  S.device = c.device;
  return 0;
}

// 1 (default): one fused kernel per group; 0: gather -> standalone batched kernel -> scatter (BigUint div_rem always)
static bool use_fused() { return option(OPT_GENSTREAM_FUSED) != 0; }   // p2e_set_option("genstream_fused", 0 | 1): the tests run both paths

static int per_launch(size_t n) {
  int per = (int)std::max<size_t>(1, GS_MAX_STAGED / std::max<size_t>(n, 1));
  return std::min(per, 65535);
}

// staging buffers for the widest group at this n (grown before any launch: nothing allocates inside a graph capture)
static int size_staging(GenStream& S, size_t n) {
  size_t need = 16, need_err = 16;
  const int per = per_launch(n);
  for (const GsGroup& G : S.groups) {
    if (G.kind == K_ARITH || G.kind == K_CONST || G.kind == K_BASESPLIT || G.kind == K_RANDACC || G.kind == K_EQUAL) continue;
    const size_t M = (size_t)std::min(G.m, per) * n;
    need = std::max(need, (size_t)(G.in_rows + G.out_rows) * M * 8);
    need_err = std::max(need_err, M * 4);
  }
  int rc;
  if ((rc = grow((void**)&S.d_stage, &S.stage_bytes, need))) return rc;
  if ((rc = grow((void**)&S.d_stage_err, &S.stage_err_bytes, need_err))) return rc;
  return 0;
}

static int run_groups(GenStream& S, size_t n, u64* V, size_t vs, u32* err, cudaStream_t st) {
  int rc;
  const unsigned gx = (unsigned)((n + GS_TPB - 1) / GS_TPB);
  const int per = per_launch(n);
  for (const GsGroup& G : S.groups) {
    const u32* idx = S.d_idx + G.idx_off;
    const u32* oidx = S.d_idx + G.oidx_off;
    if (G.kind == K_ARITH || G.kind == K_CONST) {
      for (int j0 = 0; j0 < G.m; j0 += 65535) {
        const int mc = std::min(G.m - j0, 65535);
        if (G.kind == K_CONST) {
          k_gs_const<<<dim3(gx, mc), GS_TPB, 0, st>>>(n, G.m, oidx + j0, S.d_consts + G.const_off + j0, V, vs, err);
        } else {
          // the index rows m0, m1, addend ([3][m]) and the output row follow each other in the pool: one [4][m] block
          k_gs_arith<<<dim3(gx, mc), GS_TPB, 0, st>>>(n, G.m, idx + j0, S.d_consts + G.const_off + j0, V, vs, err);
        }
      }
      P2E_CUDA(cudaGetLastError());
      continue;
    }
    if (G.kind == K_BASESPLIT || G.kind == K_RANDACC || G.kind == K_EQUAL) {
      GsIO io = {idx, oidx, V, vs, n, err, G.m};
      for (int j0 = 0; j0 < G.m; j0 += 65535) {
        const dim3 grid(gx, (unsigned)std::min(G.m - j0, 65535));
        if (G.kind == K_BASESPLIT) k_gs_base_split<<<grid, GS_TPB, 0, st>>>(io, j0, G.shape_a);
        else if (G.kind == K_RANDACC) k_gs_random_access<<<grid, GS_TPB, 0, st>>>(io, j0, G.shape_a);
        else k_gs_equality<<<grid, GS_TPB, 0, st>>>(io, j0);
      }
      P2E_CUDA(cudaGetLastError());
      continue;
    }
    if (use_fused() && G.kind != K_DIVREM) {
      GsIO io = {idx, oidx, V, vs, n, err, G.m};
      const bool base = G.modulus == P2E_SECP256K1_BASE;
      for (int j0 = 0; j0 < G.m; j0 += 65535) {
        const dim3 grid(gx, (unsigned)std::min(G.m - j0, 65535));
        switch (G.kind) {
          case K_MUL:
            if (base) k_gs_mul<ModP><<<grid, GS_TPB, 0, st>>>(io, j0); else k_gs_mul<ModN><<<grid, GS_TPB, 0, st>>>(io, j0);
            break;
          case K_CHECKSUM: k_gs_check_sum<<<grid, GS_TPB, 0, st>>>(io, j0); break;
          case K_ADD:
            if (base) k_gs_addsub<ModP, false><<<grid, GS_TPB, 0, st>>>(io, j0, G.shape_a, G.shape_b);
            else k_gs_addsub<ModN, false><<<grid, GS_TPB, 0, st>>>(io, j0, G.shape_a, G.shape_b);
            break;
          case K_SUB:
            if (base) k_gs_addsub<ModP, true><<<grid, GS_TPB, 0, st>>>(io, j0, G.shape_a, G.shape_b);
            else k_gs_addsub<ModN, true><<<grid, GS_TPB, 0, st>>>(io, j0, G.shape_a, G.shape_b);
            break;
          case K_MULTIADD:
            if (base) k_gs_add_many<ModP><<<grid, GS_TPB, 0, st>>>(io, j0, G.shape_a); else k_gs_add_many<ModN><<<grid, GS_TPB, 0, st>>>(io, j0, G.shape_a);
            break;
          case K_INV:
            if (base) k_gs_inv<ModP><<<grid, GS_TPB, 0, st>>>(io, j0); else k_gs_inv<ModN><<<grid, GS_TPB, 0, st>>>(io, j0);
            break;
          default: k_gs_glv<<<grid, GS_TPB, 0, st>>>(io, j0); break;
        }
      }
      P2E_CUDA(cudaGetLastError());
      continue;
    }
    for (int j0 = 0; j0 < G.m; j0 += per) {
      const int mc = std::min(G.m - j0, per);
      const size_t M = (size_t)mc * n;
      u64* sin = S.d_stage;
      u64* sout = S.d_stage + (size_t)G.in_rows * M;
      k_gs_gather<<<dim3(gx, mc), GS_TPB, 0, st>>>(n, G.m, mc, j0, G.in_rows, idx, V, vs, sin);
      P2E_CUDA(cudaGetLastError());
      void* s = (void*)st;
      switch (G.kind) {
        case K_MUL: rc = p2e_mul_nonnative_witness_dev(G.modulus, M, sin, sin + 9 * M, M, sout, M, S.d_stage_err, s); break;
        case K_CHECKSUM: rc = p2e_check_sum_witness_dev(M, sin, M, sout, M, S.d_stage_err, s); break;
        case K_ADD: rc = p2e_add_nonnative_witness_dev(G.modulus, M, sin, G.shape_a, sin + (size_t)G.shape_a * M, G.shape_b, M, sout, M, S.d_stage_err, s); break;
        case K_SUB: rc = p2e_sub_nonnative_witness_dev(G.modulus, M, sin, G.shape_a, sin + (size_t)G.shape_a * M, G.shape_b, M, sout, M, S.d_stage_err, s); break;
        case K_MULTIADD: rc = p2e_add_many_nonnative_witness_dev(G.modulus, M, G.shape_a, sin, M, sout, M, S.d_stage_err, s); break;
        case K_INV: rc = p2e_inv_nonnative_witness_dev(G.modulus, M, sin, M, sout, M, S.d_stage_err, s); break;
        case K_DIVREM: rc = p2e_biguint_div_rem_witness_dev(M, sin, G.shape_a, sin + (size_t)G.shape_a * M, G.shape_b, M, sout, M, S.d_stage_err, s); break;
        case K_GLV: rc = p2e_glv_decompose_witness_dev(M, sin, M, sout, M, S.d_stage_err, s); break;
        default: rc = set_error_msg("genstream: internal: unknown group kind");
      }
      if (rc) return rc;
      k_gs_scatter<<<dim3(gx, mc), GS_TPB, 0, st>>>(n, G.m, mc, j0, G.out_rows, oidx, V, vs, sout, S.d_stage_err, err);
      P2E_CUDA(cudaGetLastError());
    }
  }
  return 0;
}

static bool use_graph() { return option(OPT_GENSTREAM_GRAPH) != 0; }

// the launch sequence as a CUDA graph, re-used while (n, table, stride, err, staging) stay the same
static int run_groups_graphed(GenStream& S, size_t n, u64* V, size_t vs, u32* err, cudaStream_t st) {
  int rc;
  // deep schedules: one launch of the level walker (option "genstream_walk": 1 = when the stream has 32 levels or more,
  // 2 = always, 0 = never)
  const long walk = option(OPT_GENSTREAM_WALK);
  if (S.walkable && S.n_levels > 0 && (walk == 2 || (walk == 1 && S.n_levels >= 32))) {
    // shape: more CTAs than SMs for a small batch ("genstream_walk_shape" = 100 * warps + circuits per CTA overrides)
    long shape = option(OPT_GENSTREAM_WALK_SHAPE);
    if (shape == 0) {
      // measured, ECDSA circuit x 4096: 832 20.1 ms, 816 19.2, 416 20.5, 408 26.7, 808 38.5 (idle lanes cost issue slots)
      shape = n >= 48 * (size_t)ctx().sm_count ? 832 : 816;
    }
    switch (shape) {
      case 832: launch_walk<8, 32>(S.d_steps, S.d_level_first, S.n_levels, V, vs, n, err, st); break;
      case 816: launch_walk<8, 16>(S.d_steps, S.d_level_first, S.n_levels, V, vs, n, err, st); break;
      case 808: launch_walk<8, 8>(S.d_steps, S.d_level_first, S.n_levels, V, vs, n, err, st); break;
      case 416: launch_walk<4, 16>(S.d_steps, S.d_level_first, S.n_levels, V, vs, n, err, st); break;
      case 408: launch_walk<4, 8>(S.d_steps, S.d_level_first, S.n_levels, V, vs, n, err, st); break;
      case 404: launch_walk<4, 4>(S.d_steps, S.d_level_first, S.n_levels, V, vs, n, err, st); break;
      case 208: launch_walk<2, 8>(S.d_steps, S.d_level_first, S.n_levels, V, vs, n, err, st); break;
      default: return set_error_msg("genstream: option genstream_walk_shape: one of 832, 816, 808, 416, 408, 404, 208 (100 * warps + circuits per CTA), 0 = by batch size");
    }
    P2E_CUDA(cudaGetLastError());
    return 0;
  }
  if ((rc = size_staging(S, n))) return rc;
  if (!use_graph() || S.groups.size() < 8) return run_groups(S, n, V, vs, err, st);
  const bool fused = use_fused();
  const bool hit = S.graph && S.g_fused == fused && S.g_n == n && S.g_V == V && S.g_vs == vs && S.g_err == err && S.g_stage == S.d_stage &&
                   S.g_stage_err == S.d_stage_err;
  if (!hit) {
    S.drop_graph();
    cudaGraph_t g = nullptr;
    P2E_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    rc = run_groups(S, n, V, vs, err, st);
    cudaError_t ce = cudaStreamEndCapture(st, &g);
    if (rc || ce != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      if (rc) return rc;
      return run_groups(S, n, V, vs, err, st);        // capture not possible on this stream: plain launches
    }
    ce = cudaGraphInstantiate(&S.graph, g, 0);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) {
      S.graph = nullptr;
      cudaGetLastError();
      return run_groups(S, n, V, vs, err, st);
    }
    S.g_fused = fused; S.g_n = n; S.g_V = V; S.g_vs = vs; S.g_err = err; S.g_stage = S.d_stage; S.g_stage_err = S.d_stage_err;
  }
  P2E_CUDA(cudaGraphLaunch(S.graph, st));
  return 0;
}

}  // namespace p2e

using namespace p2e;

extern "C" {

int p2e_genstream_load(const uint8_t* bytes, size_t len) {
  if (!bytes && len) return set_error_msg("genstream: bytes is NULL");
  GenStream* S = new GenStream();
  Reader R{bytes, len};
  while (R.pos < R.len) {
    if (!parse_record(*S, R)) {
      set_error_msg((S->error + " (record " + std::to_string(S->gens.size()) + ", byte " + std::to_string(R.pos) + ")").c_str());
      delete S;
      return -1;
    }
  }
  if (!finalize(*S)) {
    set_error_msg(S->error.c_str());
    delete S;
    return -1;
  }
  std::lock_guard<std::mutex> lock(g_gs_mu);
  for (size_t h = 0; h < g_streams.size(); h++)
    if (!g_streams[h]) {
      g_streams[h] = S;
      return (int)h;
    }
  g_streams.push_back(S);
  return (int)g_streams.size() - 1;
}

int p2e_genstream_free(int h) {
  std::lock_guard<std::mutex> lock(g_gs_mu);
  GenStream* S = get_stream(h);
  if (!S) return -1;
  if (ctx().ready) S->release_device();
  delete S;
  g_streams[h] = nullptr;
  return 0;
}

/* what: 0 slots, 1 generators, 2 levels, 3 launch groups, 4 input slots, 5 targets */
long p2e_genstream_info(int h, int what) {
  std::lock_guard<std::mutex> lock(g_gs_mu);
  GenStream* S = get_stream(h);
  if (!S) return -1;
  switch (what) {
    case 0: return (long)S->n_slots;
    case 1: return (long)S->gens.size();
    case 2: return (long)S->n_levels;
    case 3: return (long)S->groups.size();
    case 4: return (long)S->input_slots.size();
    case 5: return (long)S->targets.size();
  }
  set_error_msg("genstream: unknown info selector");
  return -1;
}

long p2e_genstream_slot(int h, int is_wire, uint64_t row_or_index, uint64_t column) {
  std::lock_guard<std::mutex> lock(g_gs_mu);
  GenStream* S = get_stream(h);
  if (!S) return -1;
  auto it = S->target_id.find({is_wire ? 1ull : 0ull, row_or_index, is_wire ? column : 0});
  if (it == S->target_id.end()) {
    set_error_msg("genstream: no generator of the stream mentions this target");
    return -1;
  }
  return (long)S->slot_of_target[it->second];
}

long p2e_genstream_input_slots(int h, uint32_t* out, size_t cap) {
  std::lock_guard<std::mutex> lock(g_gs_mu);
  GenStream* S = get_stream(h);
  if (!S) return -1;
  for (size_t k = 0; k < S->input_slots.size() && k < cap; k++) out[k] = S->input_slots[k];
  return (long)S->input_slots.size();
}

long p2e_genstream_generator_levels(int h, int32_t* out, size_t cap) {
  std::lock_guard<std::mutex> lock(g_gs_mu);
  GenStream* S = get_stream(h);
  if (!S) return -1;
  for (size_t k = 0; k < S->gens.size() && k < cap; k++) out[k] = S->gens[k].level;
  return (long)S->gens.size();
}

int p2e_genstream_run_dev(int h, size_t n, const uint32_t* in_slots, int n_in, const uint64_t* in_values, size_t in_stride,
                          uint64_t* table, size_t stride, uint32_t* err, void* stream) {
  int rc = ensure_ready();
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(g_gs_mu);
  GenStream* S = get_stream(h);
  if (!S) return -1;
  if (n_in < 0 || (n_in && (!in_slots || !in_values))) return set_error_msg("genstream: bad inputs");
  if (n == 0) return 0;
  if (stride < n || (n_in && in_stride < n)) return set_error_msg("stride < n");
  // every slot no generator produces must be supplied (plonky2 would stall on it: "generators left unfinished")
  std::vector<char> given(S->n_slots, 0);
  for (int k = 0; k < n_in; k++) {
    if (in_slots[k] >= S->n_slots) return set_error_msg("genstream: input slot out of range");
    given[in_slots[k]] = 1;
  }
  for (u32 s : S->input_slots)
    if (!given[s]) return set_error_msg("genstream: an input slot (read by a generator, produced by none) was not supplied");
  if ((rc = ensure_device(*S))) return rc;
  cudaStream_t st = pick_stream(stream);
  const unsigned gx = (unsigned)((n + GS_TPB - 1) / GS_TPB);
  P2E_CUDA(cudaMemsetAsync(err, 0, n * 4, st));
  if (n_in) {
    if ((rc = grow((void**)&S->d_io, &S->io_bytes, (size_t)n_in * 4))) return rc;
    P2E_CUDA(cudaMemcpyAsync(S->d_io, in_slots, (size_t)n_in * 4, cudaMemcpyHostToDevice, st));
    for (int k0 = 0; k0 < n_in; k0 += 65535) {
      const int kc = std::min(n_in - k0, 65535);
      k_gs_set_inputs<<<dim3(gx, kc), GS_TPB, 0, st>>>(n, S->d_io + k0, in_values + (size_t)k0 * in_stride, in_stride, table, stride);
    }
    P2E_CUDA(cudaGetLastError());
  }
  return run_groups_graphed(*S, n, table, stride, err, st);
}

int p2e_genstream_run(int h, size_t n, const uint32_t* in_slots, int n_in, const uint64_t* in_values,
                      const uint32_t* out_slots, int n_out, uint64_t* out_values, uint32_t* err) {
  int rc = ensure_ready();
  if (rc) return rc;
  long n_slots = p2e_genstream_info(h, 0);
  if (n_slots < 0) return -1;
  if (n_out < 0 || (n_out && (!out_slots || !out_values))) return set_error_msg("genstream: bad outputs");
  for (int k = 0; k < n_out; k++)
    if (out_slots[k] >= (u32)n_slots) return set_error_msg("genstream: output slot out of range");
  if (n == 0) return 0;
  P2E_HOST_LOCK();
  cudaStream_t st = ctx().stream;
  void *dV = nullptr, *din = nullptr, *derr = nullptr, *dout = nullptr;
  if ((rc = get_scratch(1, (size_t)n_slots * n * 8, &dV))) return rc;
  if ((rc = get_scratch(0, ((size_t)std::max(n_in, 1)) * n * 8, &din))) return rc;
  if ((rc = get_scratch(2, n * 4, &derr))) return rc;
  if ((rc = get_scratch(3, ((size_t)std::max(n_out, 1)) * n * 8 + (size_t)std::max(n_out, 1) * 4, &dout))) return rc;
  P2E_CUDA(cudaMemsetAsync(dV, 0, (size_t)n_slots * n * 8, st));
  if (n_in) P2E_CUDA(cudaMemcpyAsync(din, in_values, (size_t)n_in * n * 8, cudaMemcpyHostToDevice, st));
  if ((rc = p2e_genstream_run_dev(h, n, in_slots, n_in, (const uint64_t*)din, n, (uint64_t*)dV, n, (uint32_t*)derr, (void*)st))) return rc;
  if (n_out) {
    u32* dslots = (u32*)((u64*)dout + (size_t)n_out * n);
    P2E_CUDA(cudaMemcpyAsync(dslots, out_slots, (size_t)n_out * 4, cudaMemcpyHostToDevice, st));
    const unsigned gx = (unsigned)((n + GS_TPB - 1) / GS_TPB);
    for (int k0 = 0; k0 < n_out; k0 += 65535) {
      const int kc = std::min(n_out - k0, 65535);
      k_gs_get_outputs<<<dim3(gx, kc), GS_TPB, 0, st>>>(n, dslots + k0, (const u64*)dV, n, (u64*)dout + (size_t)k0 * n, n);
    }
    P2E_CUDA(cudaGetLastError());
    P2E_CUDA(cudaMemcpyAsync(out_values, dout, (size_t)n_out * n * 8, cudaMemcpyDeviceToHost, st));
  }
  if (err) P2E_CUDA(cudaMemcpyAsync(err, derr, n * 4, cudaMemcpyDeviceToHost, st));
  P2E_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
