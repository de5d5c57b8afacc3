// p2e_kernels.cu -- batched witness-generator kernels (one instance per thread, limbs in
// registers, wire-major coalesced loads/stores), the gate constraint evaluators over
// Goldilocks LDE rows, and their extern "C" entry points (include/p2e.h).
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "../../include/p2e.h"
#include "p2e_common.h"
#include "p2e_field.cuh"
#include "p2e_fold.cuh"
#include "p2e_f8inv.cuh"

namespace p2e {

// ------------------------------------------------------------------------------------------
// contexts: one per device, selected per thread
static Context g_ctx[MAX_DEVICES];
static Context g_none;                         // never ready: what ctx() answers before any p2e_init
static std::mutex g_ctx_mu;                    // guards creation / destruction of contexts
static int g_default_device = -1;              // first initialised device: current device of threads that never chose
static thread_local int tls_device = -1;
static thread_local std::string tls_error;
static long g_options[N_OPTIONS] = {65536, 4, 1, 1, 12, 0, 0, 63, 1, 0, 0, 1};

Context& ctx() {
  const int d = tls_device >= 0 ? tls_device : g_default_device;
  if (d < 0 || d >= MAX_DEVICES) return g_none;
  return g_ctx[d];
}
long option(int id) { return (id >= 0 && id < N_OPTIONS) ? g_options[id] : 0; }
int set_error(const char* where, cudaError_t e) {
  tls_error = std::string(where) + ": " + cudaGetErrorString(e);
  return -(int)e - 1000;
}
int set_error_msg(const char* msg) {
  tls_error = msg;
  return -1;
}
int ensure_ready() {
  Context& c = ctx();
  if (!c.ready) return set_error_msg("p2e_init() has not been called (no CUDA device bound; there is no CPU fallback)");
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != c.device) P2E_CUDA(cudaSetDevice(c.device));
  return 0;
}
int get_scratch(int slot, size_t bytes, void** out) {
  Context& c = ctx();
  if (bytes == 0) bytes = 16;
  if (c.scratch_bytes[slot] < bytes) {
    if (c.scratch[slot]) P2E_CUDA(cudaFree(c.scratch[slot]));
    c.scratch[slot] = nullptr;
    c.scratch_bytes[slot] = 0;
    P2E_CUDA(cudaMalloc(&c.scratch[slot], bytes));
    c.scratch_bytes[slot] = bytes;
  }
  *out = c.scratch[slot];
  return 0;
}

constexpr int TPB = 128;
static inline unsigned grid_for(size_t n) { return (unsigned)((n + TPB - 1) / TPB); }

// ------------------------------------------------------------------------------------------
// A8 + A9
template <class MOD>
__global__ void __launch_bounds__(TPB)
k_mul_witness(size_t n, const u64* __restrict__ x, const u64* __restrict__ y, size_t is,
              u64* __restrict__ out, size_t os, u32* __restrict__ err) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  u64 xw[NL], yw[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) {
    xw[l] = x[l * is + i];
    yw[l] = y[l * is + i];
  }
  u32 xa[NL], ya[NL];
  u32 e = load_mul_wires(xw, xa) | load_mul_wires(yw, ya);
  u32 r[NL], q[NL];
  u64 cs[2 * NL - 1], b[2 * NL - 2];
  if (e == 0) {
    e = mul_witness_core<MOD>(xa, ya, r, q, cs, b);
  } else {
    // the reference panics before writing anything: emit zeros
#pragma unroll
    for (int l = 0; l < NL; l++) r[l] = q[l] = 0;
#pragma unroll
    for (int l = 0; l < 2 * NL - 1; l++) cs[l] = 0;
#pragma unroll
    for (int l = 0; l < 2 * NL - 2; l++) b[l] = 0;
  }
#pragma unroll
  for (int l = 0; l < NL; l++) {
    out[(P2E_MUL_OUT_R + l) * os + i] = r[l];
    out[(P2E_MUL_OUT_Q + l) * os + i] = q[l];
  }
#pragma unroll
  for (int l = 0; l < 2 * NL - 1; l++) out[(P2E_MUL_OUT_CHECK_SUM + l) * os + i] = cs[l];
#pragma unroll
  for (int l = 0; l < 2 * NL - 2; l++) out[(P2E_MUL_OUT_CARRY + l) * os + i] = b[l];
  err[i] = e;
}

__global__ void __launch_bounds__(TPB)
k_check_sum_witness(size_t n, const u64* __restrict__ a, size_t is, u64* __restrict__ out, size_t os,
                    u32* __restrict__ err) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  u64 aw[2 * NL - 1], b[2 * NL - 2];
#pragma unroll
  for (int l = 0; l < 2 * NL - 1; l++) aw[l] = a[l * is + i];
  u32 e = check_sum_core(aw, b);
#pragma unroll
  for (int l = 0; l < 2 * NL - 2; l++) out[l * os + i] = b[l];
  err[i] = e;
}

// loads `nl` (0..9) wires of a nonnative operand, absent limbs = 0
__device__ __forceinline__ void load_wires(const u64* __restrict__ p, int nl, size_t is, size_t i, u64 (&w)[NL]) {
#pragma unroll
  for (int l = 0; l < NL; l++) w[l] = (l < nl) ? p[l * is + i] : 0;
}

// A4 / A6
template <class MOD, bool IS_SUB>
__global__ void __launch_bounds__(TPB)
k_addsub_witness(size_t n, const u64* __restrict__ a, int al, const u64* __restrict__ b, int bl, size_t is,
                 u64* __restrict__ out, size_t os, u32* __restrict__ err) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  u64 aw[NL], bw[NL];
  load_wires(a, al, is, i, aw);
  load_wires(b, bl, is, i, bw);
  u32 ac[NL], bc[NL], r[NL];
  u32 e = load_canon<MOD>(aw, ac) | load_canon<MOD>(bw, bc);
  u32 ov = 0;
  if (e == 0) {
    ov = IS_SUB ? sub_witness_core<MOD>(ac, bc, r) : add_witness_core<MOD>(ac, bc, r);
  } else {
    set_zero9(r);
  }
#pragma unroll
  for (int l = 0; l < NL; l++) out[l * os + i] = r[l];
  out[NL * os + i] = ov;
  err[i] = e;
}

// A5: k summands, [k][9][stride]
template <class MOD>
__global__ void __launch_bounds__(TPB)
k_add_many_witness(size_t n, int k, const u64* __restrict__ s, size_t is, u64* __restrict__ out, size_t os,
                   u32* __restrict__ err) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  u32 acc[NL];
  set_zero9(acc);
  u32 ov = 0, e = 0;
  for (int t = 0; t < k; t++) {
    u64 w[NL];
    load_wires(s + (size_t)t * NL * is, NL, is, i, w);
    u32 v[NL], sum[NL];
    e |= load_canon<MOD>(w, v);
    add9(acc, v, sum);               // acc < M, v < M  ->  < 2M < 2^257
    if (geq_mod<MOD>(sum)) {         // running div_rem: sum of k canonical values, quotient counts wraps
      sub_mod9<MOD>(sum, acc);
      ov++;
    } else {
      copy9(sum, acc);
    }
  }
  if (e) { set_zero9(acc); ov = 0; }
#pragma unroll
  for (int l = 0; l < NL; l++) out[l * os + i] = acc[l];
  out[NL * os + i] = ov;
  err[i] = e;
}

// A7: one Fermat ladder per thread is the literal reference algorithm; the batched entry uses
// Montgomery's trick over CHUNK consecutive instances per thread (inverse is unique => bit-exact).
constexpr int INV_CHUNK = 8;
template <class MOD>
__global__ void __launch_bounds__(TPB)
k_inv_witness(size_t n, const u64* __restrict__ x, size_t is, u64* __restrict__ out, size_t os,
              u32* __restrict__ err) {
  // thread t handles instances t, t + T, t + 2T, ... (T = total threads) so loads stay coalesced
  size_t T = (size_t)gridDim.x * TPB;
  size_t t0 = (size_t)blockIdx.x * TPB + threadIdx.x;
  u32 xc[INV_CHUNK][NL];
  u32 pre[INV_CHUNK][NL];      // prefix products
  u32 e[INV_CHUNK];
  u32 run[NL];
  set_zero9(run);
  run[0] = 1;
  int cnt = 0;
#pragma unroll
  for (int c = 0; c < INV_CHUNK; c++) {
    size_t i = t0 + (size_t)c * T;
    e[c] = 0;
    set_zero9(xc[c]);
    if (i < n) {
      u64 w[NL];
      load_wires(x, NL, is, i, w);
      e[c] = load_canon<MOD>(w, xc[c]);
      if (e[c] == 0 && is_zero9(xc[c])) e[c] = ERR_INVERSE_ZERO;
      cnt = c + 1;
    } else {
      e[c] = ERR_INTERNAL;       // marks "not present"
    }
    copy9(run, pre[c]);
    if (e[c] == 0) {
      u32 t[NL];
      mulmod<MOD>(run, xc[c], t);
      copy9(t, run);
    }
  }
  if (cnt == 0) return;
  u32 inv_run[NL];
  inv_mod<MOD>(run, inv_run);     // division steps (p2e_modinv.cuh), ~20 k instructions for the whole chunk
#pragma unroll
  for (int c = INV_CHUNK - 1; c >= 0; c--) {
    size_t i = t0 + (size_t)c * T;
    if (i >= n) continue;
    u32 inv[NL], div[NL];
    set_zero9(inv);
    set_zero9(div);
    u32 ee = e[c];
    if (ee == 0) {
      u32 t[NL];
      mulmod<MOD>(inv_run, pre[c], inv);
      mulmod<MOD>(inv_run, xc[c], t);
      copy9(t, inv_run);
      ee |= inv_div_core<MOD>(xc[c], inv, div);
    }
#pragma unroll
    for (int l = 0; l < NL; l++) {
      out[l * os + i] = div[l];
      out[(NL + l) * os + i] = inv[l];
    }
    err[i] = ee;
  }
}

// A10
__global__ void __launch_bounds__(TPB)
k_glv_witness(size_t n, const u64* __restrict__ k, size_t is, u64* __restrict__ out, size_t os,
              u32* __restrict__ err) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  u64 w[NL];
  load_wires(k, NL, is, i, w);
  u32 kc[NL], k1[NL], k2[NL], n1 = 0, n2 = 0;
  u32 e = load_canon<ModN>(w, kc);
  if (e == 0) {
    glv_decompose_core(kc, k1, k2, n1, n2);
    // set_biguint_target into 5-limb targets (glv.rs:62-63): must fit 145 bits
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
    out[l * os + i] = k1[l];
    out[(5 + l) * os + i] = k2[l];
  }
  out[10 * os + i] = n1;
  out[11 * os + i] = n2;
  err[i] = e;
}

// A3: Knuth algorithm D in base 2^29 on local arrays (generic limb counts; not on the ECDSA path)
constexpr int DR_MAX = 18;      // max wires per operand
constexpr int DR_W = DR_MAX + 3; // normalised limbs an operand of DR_MAX raw u64 wires can need
__device__ int dr_load(const u64* __restrict__ p, int nl, size_t is, size_t i, u32* v) {
  // Horner over raw wires -> normalised limbs; returns significant limb count
  u64 c = 0;
  int top = 0;
  for (int l = 0; l < DR_W; l++) {
    u64 w = (l < nl) ? p[l * is + i] : 0;
    u64 lo = w + c;
    u64 ov = lo < w ? 1 : 0;
    v[l] = (u32)lo & LMASK;
    c = (lo >> LB) | (ov << (64 - LB));
    if (v[l]) top = l + 1;
  }
  return top;
}

// Length of convert_base(value.to_u32_digits(), 32, 29) (biguint.rs:27-51): the loop emits
// floor(32 * n32 / 29) limbs -- possibly including zero top limbs -- plus one iff bits remain.
__device__ int dr_conv_len(const u32* v) {
  int top = -1;
  for (int l = 0; l < DR_W; l++) if (v[l]) top = l;
  if (top < 0) return 0;
  int bits = LB * top + (32 - __clz(v[top]));
  int n32 = (bits + 31) / 32;
  int k = (32 * n32) / LB;
  return k + (bits > LB * k ? 1 : 0);
}

__global__ void __launch_bounds__(TPB)
k_divrem_witness(size_t n, const u64* __restrict__ a, int al, const u64* __restrict__ b, int bl, size_t is,
                 u64* __restrict__ out, size_t os, u32* __restrict__ err) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const int div_limbs = (bl > al + 1) ? 0 : al - bl + 1;
  u32 u[DR_W + 1], v[DR_W], q[DR_W];
  int m = dr_load(a, al, is, i, u);
  int nn = dr_load(b, bl, is, i, v);
  u[DR_W] = 0;
  for (int l = 0; l < DR_W; l++) q[l] = 0;
  u32 e = 0;
  if (nn == 0) {
    e = ERR_DIV_ZERO;
  } else if (m >= nn) {
    if (nn == 1) {
      u64 rem = 0;
      for (int j = m - 1; j >= 0; j--) {
        u64 cur = (rem << LB) | u[j];
        q[j] = (u32)(cur / v[0]);
        rem = cur % v[0];
      }
      for (int l = 0; l < DR_W; l++) u[l] = 0;
      u[0] = (u32)rem;
    } else {
      int s = __clz(v[nn - 1]) - 3;            // shift so that v[nn-1] >= 2^28
      if (s > 0) {
        for (int l = nn - 1; l > 0; l--) v[l] = ((v[l] << s) | (v[l - 1] >> (LB - s))) & LMASK;
        v[0] = (v[0] << s) & LMASK;
        u[m] = u[m - 1] >> (LB - s);
        for (int l = m - 1; l > 0; l--) u[l] = ((u[l] << s) | (u[l - 1] >> (LB - s))) & LMASK;
        u[0] = (u[0] << s) & LMASK;
      } else {
        u[m] = 0;
      }
      for (int j = m - nn; j >= 0; j--) {
        u64 num = ((u64)u[j + nn] << LB) | u[j + nn - 1];
        u64 qh = num / v[nn - 1];
        u64 rh = num % v[nn - 1];
        while (qh >> LB || qh * v[nn - 2] > ((rh << LB) | u[j + nn - 2])) {
          qh--;
          rh += v[nn - 1];
          if (rh >> LB) break;
        }
        i64 br = 0;
        u64 cy = 0;
        for (int l = 0; l < nn; l++) {
          u64 p = qh * v[l] + cy;
          cy = p >> LB;
          i64 t = (i64)u[j + l] - (i64)(p & LMASK) + br;
          u[j + l] = (u32)t & LMASK;
          br = t >> LB;
        }
        i64 t = (i64)u[j + nn] - (i64)cy + br;
        u[j + nn] = (u32)t & LMASK;
        if (t < 0) {
          qh--;
          u32 c2 = 0;
          for (int l = 0; l < nn; l++) {
            u32 t2 = u[j + l] + v[l] + c2;
            u[j + l] = t2 & LMASK;
            c2 = t2 >> LB;
          }
          u[j + nn] = (u[j + nn] + c2) & LMASK;
        }
        q[j] = (u32)qh;
      }
      if (s > 0) {
        for (int l = 0; l < nn; l++) u[l] = ((u[l] >> s) | (u[l + 1] << (LB - s))) & LMASK;
      }
      for (int l = nn; l <= DR_W; l++) u[l] = 0;
    }
  }
  // set_biguint_target asserts target.num_limbs() >= convert_base(32 -> 29).len() (biguint.rs:472-473)
  if (e == 0) {
    if (dr_conv_len(q) > div_limbs) e |= ERR_LIMBS_OVERFLOW;
    if (dr_conv_len(u) > bl) e |= ERR_LIMBS_OVERFLOW;
  }
  for (int l = 0; l < div_limbs; l++) out[l * os + i] = e ? 0 : q[l];
  for (int l = 0; l < bl; l++) out[(div_limbs + l) * os + i] = e ? 0 : u[l];
  err[i] = e;
}

// ------------------------------------------------------------------------------------------
// Constraint evaluation over Goldilocks rows (A11, A12).  One LDE row per thread; wires are
// arbitrary canonical field elements.  Column sums are accumulated un-reduced in 96-bit
// accumulators per 32x32 partial-product class and reduced once per constraint.
// (QA0 + QA1 + (QB0 + QB1)*2^32 + add) mod P for carry-free 64-bit partial sums
__device__ __forceinline__ u64 reduce_q(u64 qa0, u64 qa1, u64 qb0, u64 qb1, u64 add) {
  u64 t = (u64)(u32)qa0 + (u32)qa1 + (u32)add;
  u32 W0 = (u32)t;
  t = (t >> 32) + (qa0 >> 32) + (qa1 >> 32) + (add >> 32) + (u32)qb0 + (u32)qb1;
  u32 W1 = (u32)t;
  t = (t >> 32) + (qb0 >> 32) + (qb1 >> 32);
  return gl_reduce5(W0, W1, (u32)t, (u32)(t >> 32), 0);
}

// One row of MulNonnativeGate::eval_unfiltered (gates/mul_nonnative.rs:101-130); constraint c goes to sink.put(c, v)
template <class MOD, class SINK>
__device__ __forceinline__ SINK eval_mul_nonnative_row(const u64* __restrict__ wires, size_t is, size_t i, SINK sink) {
  u32 xl[NL], xh[NL], yl[NL], yh[NL], ql[NL], qh[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) {
    u64 x = wires[l * is + i], y = wires[(NL + l) * is + i], q = wires[(3 * NL + l) * is + i];
    xl[l] = (u32)x; xh[l] = (u32)(x >> 32);
    yl[l] = (u32)y; yh[l] = (u32)(y >> 32);
    ql[l] = (u32)q; qh[l] = (u32)(q >> 32);
  }
#pragma unroll
  for (int c = 0; c < 2 * NL - 1; c++) {
    const u64 csw = wires[(4 * NL + c) * is + i];
    const u64 rw = c < NL ? wires[(2 * NL + c) * is + i] : 0;
    Acc3 A = {0, 0, 0}, B = {0, 0, 0}, C = {0, 0, 0};
    u64 qa0 = 0, qa1 = 0, qb0 = 0, qb1 = 0;     // m_j < 2^29: five terms stay below 2^64
#pragma unroll
    for (int j = 0; j < NL; j++) {
      const int k = c - j;
      if (k >= 0 && k < NL) {
        mac3(A, xl[j], yl[k]);
        mac3(B, xl[j], yh[k]);
        mac3(B, xh[j], yl[k]);
        mac3(C, xh[j], yh[k]);
        if (j < 5) {
          qa0 += (u64)MOD::m(j) * ql[k];
          qb0 += (u64)MOD::m(j) * qh[k];
        } else {
          qa1 += (u64)MOD::m(j) * ql[k];
          qb1 += (u64)MOD::m(j) * qh[k];
        }
      }
    }
    u64 neg = reduce_abc(A, B, C, csw);
    u64 pos = reduce_q(qa0, qa1, qb0, qb1, rw);
    sink.put(c, gl_sub(pos, neg));
  }
  return sink;
}

// One row of CheckSumGate::eval_unfiltered (gates/mul_nonnative.rs:411-427)
template <class SINK>
__device__ __forceinline__ SINK eval_check_sum_row(const u64* __restrict__ wires, size_t is, size_t i, SINK sink) {
  u64 last = 0;
#pragma unroll
  for (int c = 0; c < 2 * NL - 1; c++) {
    u64 a = wires[c * is + i];
    u64 s = gl_add(a, last);
    if (c < 2 * NL - 2) {
      u64 o = gl_sub(wires[(2 * NL - 1 + c) * is + i], 1ull << 33);
      u64 prod = gl_reduce128(o << LB, o >> (64 - LB));   // 2^29 * o
      sink.put(c, gl_sub(s, prod));
      last = o;
    } else {
      sink.put(c, s);
    }
  }
  return sink;
}

// Gate::eval_unfiltered_base_batch: every constraint to its own column
struct StoreSink {
  u64* __restrict__ out;
  size_t os, i;
  __device__ __forceinline__ void put(int c, u64 v) { out[c * os + i] = v; }
};

template <class MOD>
__global__ void __launch_bounds__(TPB, 4)
k_eval_mul_nonnative(size_t n, const u64* __restrict__ wires, size_t is, u64* __restrict__ out, size_t os) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  StoreSink sink = {out, os, i};
  eval_mul_nonnative_row<MOD>(wires, is, i, sink);
}

__global__ void __launch_bounds__(TPB)
k_eval_check_sum(size_t n, const u64* __restrict__ wires, size_t is, u64* __restrict__ out, size_t os) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  StoreSink sink = {out, os, i};
  eval_check_sum_row(wires, is, i, sink);
}

// ------------------------------------------------------------------------------------------
// Selector filtering + alpha-combination around the gate evaluators.  plonky2 (out of tree,
// restated; see include/p2e.h): Gate::eval_filtered_base_batch multiplies every constraint of a row by
//   filter(s) = prod_{j in group_range, j != gate_index} (j - s) * (many_selectors ? (u32::MAX - s) : 1),
// s = the row's selector constant; evaluate_gate_constraints_base_batch adds the gates' constraint vectors
// position by position, and eval_vanishing_poly_base_batch folds the terms with reduce_with_powers_multi:
// term t contributes term * alpha^t for every challenge alpha.  Fused here: one pass over the wires produces
//   out[a][row] (+)= filter(row) * sum_c constraint_c(row) * alpha_a^(first_power + c)
// so a gate costs (wires + 1 selector + n_alphas) columns of HBM traffic instead of (wires + constraints).
struct AlphaPow17 {
  u64 p[MAX_ALPHAS][2 * NL - 1];   // alpha_a^c
  u64 scale[MAX_ALPHAS];           // alpha_a^first_power
  u64 ma[MAX_ALPHAS];              // MulNonnative: M(alpha_a) = sum_j m_j alpha_a^j for the gate's modulus
                                   // CheckSum: alpha_a - 2^29
  u64 kb[MAX_ALPHAS];              // CheckSum: 2^33 * sum_{i<16} alpha_a^i
};

// MulNonnativeGate, fused.  The alpha-combination of a convolution is a product of evaluations:
//   sum_c alpha^c constraint_c = M(alpha) Q(alpha) - X(alpha) Y(alpha) + R(alpha) - CS(alpha),
// X(alpha) = sum_j x_j alpha^j etc. (constraint_c = sum_j (m_j q_{c-j} - x_j y_{c-j}) + r_c - check_sum_c,
// gates/mul_nonnative.rs:101-130) -- the same field element as evaluating the 17 constraints and folding them,
// for 53 lazily accumulated products and 2 multiplications per challenge instead of 162 + 17.
// res[a] = sum_{j < LEN} v[j] * alpha_a^(j) over already-loaded wires, accumulated un-reduced (three 96-bit
// accumulators per challenge: 12 instructions per product), one reduction per challenge
template <int NA, int LEN>
__device__ __forceinline__ void poly_at_alphas(const u64 (&v)[LEN], const AlphaPow17& P, u64 (&res)[NA]) {
  Acc3 A[NA], B[NA], C[NA];
#pragma unroll
  for (int a = 0; a < NA; a++) A[a] = B[a] = C[a] = Acc3{0, 0, 0};
#pragma unroll
  for (int j = 0; j < LEN; j++) {
    const u32 vl = (u32)v[j], vh = (u32)(v[j] >> 32);
#pragma unroll
    for (int a = 0; a < NA; a++) {
      const u64 pw = P.p[a][j];
      const u32 pl = (u32)pw, ph = (u32)(pw >> 32);
      mac3(A[a], vl, pl);
      mac3(B[a], vl, ph);
      mac3(B[a], vh, pl);
      mac3(C[a], vh, ph);
    }
  }
#pragma unroll
  for (int a = 0; a < NA; a++) res[a] = reduce_abc(A[a], B[a], C[a], 0);
}
// A run of wires loaded as one wave: volatile asm keeps the loads together at the top (ptxas otherwise sinks each
// load next to its first use to save registers and the kernel becomes latency-bound: long_scoreboard 15 per issue)
__device__ __forceinline__ u64 ldg_wave(const u64* p) {
  u64 v;
  asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
template <int LEN>
__device__ __forceinline__ void load_run(const u64* __restrict__ wires, size_t is, size_t i, int first_wire, u64 (&v)[LEN]) {
#pragma unroll
  for (int j = 0; j < LEN; j++) v[j] = ldg_wave(wires + (size_t)(first_wire + j) * is + i);
}

template <int NA>
__device__ __forceinline__ void store_scaled(const FilterArgs& F, const AlphaPow17& P, const u64 (&t)[NA], u64 f,
                                             u64* __restrict__ out, size_t os, size_t i) {
#pragma unroll
  for (int a = 0; a < NA; a++) {
    if (a < F.n_alphas) {
      u64 r = gl_mul(gl_mul(f, P.scale[a]), t[a]);
      if (F.accumulate) r = gl_add(r, out[a * os + i]);
      out[a * os + i] = r;
    }
  }
}

// Loads are issued in two waves (x, y, r: 27 wires; q, check_sum: 26 wires) so that a warp keeps ~7 KB in flight:
// the kernel is bound by load latency otherwise (ncu: long_scoreboard 10.6 per issue with 9-wire waves).
template <int NA>
__global__ void __launch_bounds__(TPB, 4)
k_eval_mul_nonnative_fc(size_t n, const u64* __restrict__ wires, size_t is, FilterArgs F, const __grid_constant__ AlphaPow17 P,
                        u64* __restrict__ out, size_t os) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  u64 X[NA], Y[NA], R[NA], Q[NA], CS[NA];
  {
    u64 x[NL], y[NL], r[NL];
    load_run<NL>(wires, is, i, 0, x);
    load_run<NL>(wires, is, i, NL, y);
    load_run<NL>(wires, is, i, 2 * NL, r);
    poly_at_alphas<NA, NL>(x, P, X);
    poly_at_alphas<NA, NL>(y, P, Y);
    poly_at_alphas<NA, NL>(r, P, R);
  }
  {
    u64 q[NL], cs[2 * NL - 1];
    load_run<NL>(wires, is, i, 3 * NL, q);
    load_run<2 * NL - 1>(wires, is, i, 4 * NL, cs);
    poly_at_alphas<NA, NL>(q, P, Q);
    poly_at_alphas<NA, 2 * NL - 1>(cs, P, CS);
  }
  const u64 f = compute_filter(F, i);
  u64 t[NA];
#pragma unroll
  for (int a = 0; a < NA; a++) {
    t[a] = gl_sub(gl_mul(Q[a], P.ma[a]), gl_mul(X[a], Y[a]));
    t[a] = gl_sub(gl_add(t[a], R[a]), CS[a]);
  }
  store_scaled<NA>(F, P, t, f, out, os, i);
}

// CheckSumGate, fused: with o_i = b_i - 2^33 the constraints are c_i = a_i + o_{i-1} - 2^29 o_i (i < 16), c_16 = a_16 + o_15
// (gates/mul_nonnative.rs:411-427), so  sum_i alpha^i c_i = A(alpha) + (alpha - 2^29) (B(alpha) - 2^33 sum_{i<16} alpha^i).
template <int NA>
__global__ void __launch_bounds__(TPB)
k_eval_check_sum_fc(size_t n, const u64* __restrict__ wires, size_t is, FilterArgs F, const __grid_constant__ AlphaPow17 P,
                    u64* __restrict__ out, size_t os) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  u64 a[2 * NL - 1], b[2 * NL - 2];
  load_run<2 * NL - 1>(wires, is, i, 0, a);
  load_run<2 * NL - 2>(wires, is, i, 2 * NL - 1, b);
  u64 Aa[NA], Bb[NA], t[NA];
  poly_at_alphas<NA, 2 * NL - 1>(a, P, Aa);
  poly_at_alphas<NA, 2 * NL - 2>(b, P, Bb);
  const u64 f = compute_filter(F, i);
#pragma unroll
  for (int k = 0; k < NA; k++) t[k] = gl_add(Aa[k], gl_mul(P.ma[k], gl_sub(Bb[k], P.kb[k])));
  store_scaled<NA>(F, P, t, f, out, os, i);
}

// generic form (any gate): constraints [C][n] already evaluated -> filter, combine with the power table [n_alphas][C].
// Un-reduced accumulation (C <= 2^16 terms stay below 2^96 per accumulator), one reduction per challenge; the
// constraint column is read in waves of 16 loads per thread.
template <int NA>
__global__ void __launch_bounds__(TPB)
k_filter_combine(size_t n, int C, const u64* __restrict__ cons, size_t cs, const u64* __restrict__ pw, FilterArgs F,
                 u64* __restrict__ out, size_t os) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  Acc3 A[NA], B[NA], Cc[NA];
#pragma unroll
  for (int a = 0; a < NA; a++) A[a] = B[a] = Cc[a] = Acc3{0, 0, 0};
  constexpr int WAVE = 16;
  for (int c0 = 0; c0 < C; c0 += WAVE) {
    u64 v[WAVE];
#pragma unroll
    for (int k = 0; k < WAVE; k++) v[k] = (c0 + k < C) ? ldg_wave(cons + (size_t)(c0 + k) * cs + i) : 0;
#pragma unroll
    for (int k = 0; k < WAVE; k++) {
      const u32 vl = (u32)v[k], vh = (u32)(v[k] >> 32);
      const int c = (c0 + k < C) ? c0 + k : C - 1;        // padded terms multiply v = 0
#pragma unroll
      for (int a = 0; a < NA; a++) {
        const u64 pwv = __ldg(pw + (size_t)a * C + c);
        const u32 pl = (u32)pwv, ph = (u32)(pwv >> 32);
        mac3(A[a], vl, pl);
        mac3(B[a], vl, ph);
        mac3(B[a], vh, pl);
        mac3(Cc[a], vh, ph);
      }
    }
  }
  const u64 f = compute_filter(F, i);
#pragma unroll
  for (int a = 0; a < NA; a++) {
    if (a < F.n_alphas) {
      u64 r = gl_mul(f, reduce_abc(A[a], B[a], Cc[a], 0));
      if (F.accumulate) r = gl_add(r, out[a * os + i]);
      out[a * os + i] = r;
    }
  }
}

// pw[a][c] = alpha_a^(first_power + c)
struct Alphas {
  u64 v[MAX_ALPHAS];
};
__global__ void k_alpha_powers(Alphas al, int n_alphas, u64 first_power, int C, u64* __restrict__ pw) {
  const int a = threadIdx.x;
  if (a >= n_alphas) return;
  u64 p = 1, b = al.v[a];
  for (u64 e = first_power; e; e >>= 1) {
    if (e & 1) p = gl_mul(p, b);
    b = gl_mul(b, b);
  }
  for (int c = 0; c < C; c++) {
    pw[(size_t)a * C + c] = p;
    p = gl_mul(p, al.v[a]);
  }
}

// ------------------------------------------------------------------------------------------
// INT32 multiply-add peak microbenchmark (roofline denominator for the witness kernels)
template <bool WIDE>
__global__ void __launch_bounds__(256) k_imad_bench(int iters, u32 seed, u64* sink) {
  constexpr int CH = 8;
  u32 m = seed | 1u;
  if (WIDE) {
    u64 acc[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) acc[k] = threadIdx.x + k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < CH; k++) acc[k] = (u64)(u32)acc[k] * (u64)m + acc[k];   // IMAD.WIDE.U32
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) s ^= acc[k];
    if (s == 0x1234567ull) sink[0] = s;
  } else {
    u32 acc[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) acc[k] = threadIdx.x + k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < CH; k++) acc[k] = acc[k] * m + (u32)k;                    // IMAD
    }
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) s ^= acc[k];
    if (s == 0x1234567u) sink[0] = s;
  }
}

}  // namespace p2e

// ==========================================================================================
// extern "C"
using namespace p2e;

#define P2E_CHECK_READY() \
  do {                    \
    int rc__ = ensure_ready(); \
    if (rc__) return rc__;     \
  } while (0)
#define P2E_LAUNCH_CHECK() P2E_CUDA(cudaGetLastError())

extern "C" {

int p2e_init(int device) {
  // the chain entry points pipeline independent batches on many streams: ask for the maximum number of hardware work
  // queues unless the host already chose (only effective when the CUDA context does not exist yet)
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess) return set_error("cudaGetDeviceCount", e);
  if (cnt == 0) return set_error_msg("no CUDA device visible (there is no CPU fallback)");
  if (device < 0 || device >= cnt || device >= MAX_DEVICES) return set_error_msg("p2e_init: device index out of range");
  std::lock_guard<std::mutex> lock(g_ctx_mu);
  P2E_CUDA(cudaSetDevice(device));
  Context& c = g_ctx[device];
  if (!c.ready) {
    cudaDeviceProp prop;
    P2E_CUDA(cudaGetDeviceProperties(&prop, device));
    c.sm_count = prop.multiProcessorCount;
    P2E_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.device = device;
    c.ready = true;
  }
  if (g_default_device < 0) g_default_device = device;
  tls_device = device;
  return 0;
}

int p2e_init_devices(int n_devices) {
  int cnt = p2e_device_count();
  if (n_devices <= 0 || n_devices > cnt) return set_error_msg("p2e_init_devices: n_devices out of range");
  const int keep = tls_device;
  for (int d = 0; d < n_devices; d++) {
    int rc = p2e_init(d);
    if (rc) return rc;
  }
  tls_device = keep >= 0 ? keep : 0;
  return ensure_ready();
}

int p2e_set_device(int device) {
  if (device < 0 || device >= MAX_DEVICES || !g_ctx[device].ready) return set_error_msg("p2e_set_device: device not initialised (p2e_init)");
  tls_device = device;
  return ensure_ready();
}

int p2e_current_device(void) { return ctx().ready ? ctx().device : -1; }

static void shutdown_one(Context& c) {
  if (!c.ready) return;
  cudaSetDevice(c.device);
  cudaDeviceSynchronize();
  chain_release_device(c);
  ntt_release_device(c);
  genstream_release_device(c);
  for (int i = 0; i < 4; i++) {
    if (c.scratch[i]) cudaFree(c.scratch[i]);
    c.scratch[i] = nullptr;
    c.scratch_bytes[i] = 0;
  }
  cudaStreamDestroy(c.stream);
  c.stream = nullptr;
  c.ready = false;
  c.device = -1;
}

void p2e_shutdown(void) {
  std::lock_guard<std::mutex> lock(g_ctx_mu);
  for (auto& c : g_ctx) shutdown_one(c);
  g_default_device = -1;
  tls_device = -1;
}

int p2e_shutdown_device(int device) {
  if (device < 0 || device >= MAX_DEVICES) return set_error_msg("p2e_shutdown_device: device index out of range");
  std::lock_guard<std::mutex> lock(g_ctx_mu);
  shutdown_one(g_ctx[device]);
  if (g_default_device == device) {
    g_default_device = -1;
    for (int d = 0; d < MAX_DEVICES; d++) if (g_ctx[d].ready) { g_default_device = d; break; }
  }
  if (tls_device == device) tls_device = -1;
  return 0;
}

const char* p2e_last_error(void) { return tls_error.c_str(); }

static const char* const OPTION_NAMES[N_OPTIONS] = {"chain_chunk", "host_chunks", "genstream_fused", "genstream_graph", "ntt_log_tile", "chain_trace", "chain_split", "chain_phases", "genstream_walk", "genstream_walk_shape", "chain_replay_lanes", "chain_graph"};
int p2e_set_option(const char* name, long value) {
  for (int i = 0; i < N_OPTIONS; i++)
    if (name && !strcmp(name, OPTION_NAMES[i])) {
      g_options[i] = value;
      return 0;
    }
  return set_error_msg("p2e_set_option: unknown option");
}
long p2e_get_option(const char* name) {
  for (int i = 0; i < N_OPTIONS; i++)
    if (name && !strcmp(name, OPTION_NAMES[i])) return g_options[i];
  return -1;
}

int p2e_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) return 0;
  return cnt;
}

int p2e_synchronize(void) {
  P2E_CHECK_READY();
  P2E_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

int p2e_malloc(void** dptr, size_t bytes) {
  P2E_CHECK_READY();
  P2E_CUDA(cudaMalloc(dptr, bytes ? bytes : 16));
  return 0;
}
int p2e_free(void* dptr) {
  P2E_CUDA(cudaFree(dptr));
  return 0;
}
int p2e_host_alloc(void** hptr, size_t bytes) {
  P2E_CHECK_READY();
  P2E_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 16, cudaHostAllocDefault));
  return 0;
}
int p2e_host_free(void* hptr) {
  P2E_CUDA(cudaFreeHost(hptr));
  return 0;
}
int p2e_memcpy_h2d(void* dst, const void* src, size_t bytes) {
  P2E_CHECK_READY();
  P2E_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx().stream));
  P2E_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}
int p2e_memcpy_d2h(void* dst, const void* src, size_t bytes) {
  P2E_CHECK_READY();
  P2E_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx().stream));
  P2E_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

static int bad_modulus(int m) {
  if (m != P2E_SECP256K1_BASE && m != P2E_SECP256K1_SCALAR) return set_error_msg("unknown modulus id");
  return 0;
}
static int bad_limbs(int l) {
  if (l < 0 || l > NL) return set_error_msg("limb count must be in 0..9");
  return 0;
}

// ---- device-pointer entry points
int p2e_mul_nonnative_witness_dev(int modulus, size_t n, const uint64_t* x, const uint64_t* y, size_t is,
                                  uint64_t* out, size_t os, uint32_t* err, void* stream) {
  P2E_CHECK_READY();
  if (bad_modulus(modulus)) return -1;
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  cudaStream_t st = pick_stream(stream);
  if (modulus == P2E_SECP256K1_BASE)
    k_mul_witness<ModP><<<grid_for(n), TPB, 0, st>>>(n, x, y, is, out, os, err);
  else
    k_mul_witness<ModN><<<grid_for(n), TPB, 0, st>>>(n, x, y, is, out, os, err);
  P2E_LAUNCH_CHECK();
  return 0;
}

int p2e_check_sum_witness_dev(size_t n, const uint64_t* a, size_t is, uint64_t* b, size_t os, uint32_t* err,
                              void* stream) {
  P2E_CHECK_READY();
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  k_check_sum_witness<<<grid_for(n), TPB, 0, pick_stream(stream)>>>(n, a, is, b, os, err);
  P2E_LAUNCH_CHECK();
  return 0;
}

int p2e_add_nonnative_witness_dev(int modulus, size_t n, const uint64_t* a, int al, const uint64_t* b, int bl,
                                  size_t is, uint64_t* out, size_t os, uint32_t* err, void* stream) {
  P2E_CHECK_READY();
  if (bad_modulus(modulus) || bad_limbs(al) || bad_limbs(bl)) return -1;
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  cudaStream_t st = pick_stream(stream);
  if (modulus == P2E_SECP256K1_BASE)
    k_addsub_witness<ModP, false><<<grid_for(n), TPB, 0, st>>>(n, a, al, b, bl, is, out, os, err);
  else
    k_addsub_witness<ModN, false><<<grid_for(n), TPB, 0, st>>>(n, a, al, b, bl, is, out, os, err);
  P2E_LAUNCH_CHECK();
  return 0;
}

int p2e_sub_nonnative_witness_dev(int modulus, size_t n, const uint64_t* a, int al, const uint64_t* b, int bl,
                                  size_t is, uint64_t* out, size_t os, uint32_t* err, void* stream) {
  P2E_CHECK_READY();
  if (bad_modulus(modulus) || bad_limbs(al) || bad_limbs(bl)) return -1;
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  cudaStream_t st = pick_stream(stream);
  if (modulus == P2E_SECP256K1_BASE)
    k_addsub_witness<ModP, true><<<grid_for(n), TPB, 0, st>>>(n, a, al, b, bl, is, out, os, err);
  else
    k_addsub_witness<ModN, true><<<grid_for(n), TPB, 0, st>>>(n, a, al, b, bl, is, out, os, err);
  P2E_LAUNCH_CHECK();
  return 0;
}

int p2e_add_many_nonnative_witness_dev(int modulus, size_t n, int k, const uint64_t* s, size_t is, uint64_t* out,
                                       size_t os, uint32_t* err, void* stream) {
  P2E_CHECK_READY();
  if (bad_modulus(modulus)) return -1;
  if (k < 2) return set_error_msg("add_many needs k >= 2 summands (k == 1 returns its input, nonnative.rs:315-317)");
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  cudaStream_t st = pick_stream(stream);
  if (modulus == P2E_SECP256K1_BASE)
    k_add_many_witness<ModP><<<grid_for(n), TPB, 0, st>>>(n, k, s, is, out, os, err);
  else
    k_add_many_witness<ModN><<<grid_for(n), TPB, 0, st>>>(n, k, s, is, out, os, err);
  P2E_LAUNCH_CHECK();
  return 0;
}

int p2e_inv_nonnative_witness_dev(int modulus, size_t n, const uint64_t* x, size_t is, uint64_t* out, size_t os,
                                  uint32_t* err, void* stream) {
  P2E_CHECK_READY();
  if (bad_modulus(modulus)) return -1;
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  cudaStream_t st = pick_stream(stream);
  size_t threads = (n + INV_CHUNK - 1) / INV_CHUNK;
  if (modulus == P2E_SECP256K1_BASE)
    k_inv_witness<ModP><<<grid_for(threads), TPB, 0, st>>>(n, x, is, out, os, err);
  else
    k_inv_witness<ModN><<<grid_for(threads), TPB, 0, st>>>(n, x, is, out, os, err);
  P2E_LAUNCH_CHECK();
  return 0;
}

int p2e_glv_decompose_witness_dev(size_t n, const uint64_t* k, size_t is, uint64_t* out, size_t os, uint32_t* err,
                                  void* stream) {
  P2E_CHECK_READY();
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  k_glv_witness<<<grid_for(n), TPB, 0, pick_stream(stream)>>>(n, k, is, out, os, err);
  P2E_LAUNCH_CHECK();
  return 0;
}

int p2e_biguint_div_rem_witness_dev(size_t n, const uint64_t* a, int al, const uint64_t* b, int bl, size_t is,
                                    uint64_t* out, size_t os, uint32_t* err, void* stream) {
  P2E_CHECK_READY();
  if (al < 0 || al > DR_MAX || bl < 1 || bl > DR_MAX) return set_error_msg("div_rem: limb counts out of range");
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  k_divrem_witness<<<grid_for(n), TPB, 0, pick_stream(stream)>>>(n, a, al, b, bl, is, out, os, err);
  P2E_LAUNCH_CHECK();
  return 0;
}

int p2e_eval_mul_nonnative_dev(int modulus, size_t n, const uint64_t* wires, size_t is, uint64_t* out, size_t os,
                               void* stream) {
  P2E_CHECK_READY();
  if (bad_modulus(modulus)) return -1;
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  cudaStream_t st = pick_stream(stream);
  if (modulus == P2E_SECP256K1_BASE)
    k_eval_mul_nonnative<ModP><<<grid_for(n), TPB, 0, st>>>(n, wires, is, out, os);
  else
    k_eval_mul_nonnative<ModN><<<grid_for(n), TPB, 0, st>>>(n, wires, is, out, os);
  P2E_LAUNCH_CHECK();
  return 0;
}

int p2e_eval_check_sum_dev(size_t n, const uint64_t* wires, size_t is, uint64_t* out, size_t os, void* stream) {
  P2E_CHECK_READY();
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  k_eval_check_sum<<<grid_for(n), TPB, 0, pick_stream(stream)>>>(n, wires, is, out, os);
  P2E_LAUNCH_CHECK();
  return 0;
}


// ---- filtered + alpha-combined evaluation
static int gate_shape(int gate, const int* prm, int* wires, int* cons) {
  if (!prm) return set_error_msg("eval_filtered: params is NULL");
  switch (gate) {
    case P2E_GATE_MUL_NONNATIVE:
      if (bad_modulus(prm[0])) return -1;
      *wires = 6 * NL - 1; *cons = 2 * NL - 1;
      return 0;
    case P2E_GATE_CHECK_SUM:
      *wires = 4 * NL - 3; *cons = 2 * NL - 1;
      return 0;
    case P2E_GATE_UX: {
      int ni;
      return p2e_ux_gate_shape(prm[0], prm[1], prm[2], prm[3], prm[4], wires, cons, &ni);
    }
  }
  return set_error_msg("eval_filtered: unknown gate");
}

int p2e_gate_shape(int gate, const int* params, int* num_wires, int* num_constraints) {
  int w = 0, c = 0;
  int rc = gate_shape(gate, params, &w, &c);
  if (rc) return rc;
  if (num_wires) *num_wires = w;
  if (num_constraints) *num_constraints = c;
  return 0;
}

size_t p2e_eval_filtered_scratch_bytes(int gate, const int* params, size_t n) {
  int w = 0, c = 0;
  if (gate != P2E_GATE_UX || gate_shape(gate, params, &w, &c)) return 16;
  if (params[0] == P2E_UX_RANGE_CHECK) return (size_t)MAX_ALPHAS * c * 8;       // fused: only the table of powers
  return ((size_t)c * n + (size_t)MAX_ALPHAS * c) * 8;
}

int p2e_eval_filtered_combined_dev(int gate, const int* params, int gate_index, int group_start, int group_end,
                                   int many_selectors, size_t n, const uint64_t* wires, size_t is,
                                   const uint64_t* selector, const uint64_t* alphas, int n_alphas, uint64_t first_power,
                                   int accumulate, uint64_t* out, size_t os, uint64_t* scratch, void* stream) {
  P2E_CHECK_READY();
  int nw = 0, nc = 0, rc;
  if ((rc = gate_shape(gate, params, &nw, &nc))) return rc;
  if (n_alphas < 1 || n_alphas > MAX_ALPHAS || !alphas) return set_error_msg("eval_filtered: n_alphas must be 1..4");
  if (selector && (group_end <= group_start || gate_index < group_start || gate_index >= group_end))
    return set_error_msg("eval_filtered: gate_index outside its selector group");
  if (n == 0) return 0;
  if (is < n || os < n) return set_error_msg("stride < n");
  for (int a = 0; a < n_alphas; a++)
    if (alphas[a] >= GL_P) return set_error_msg("eval_filtered: alpha is not a canonical field element");
  cudaStream_t st = pick_stream(stream);
  FilterArgs F = {selector, gate_index, group_start, group_end, many_selectors ? 1 : 0, n_alphas, accumulate ? 1 : 0};
  if (gate == P2E_GATE_UX) {
    if (!scratch) return set_error_msg("eval_filtered: this gate needs p2e_eval_filtered_scratch_bytes() of scratch");
    u64* pw = scratch;                                  // [n_alphas][nc]
    u64* cons = scratch + (size_t)MAX_ALPHAS * nc;      // [nc][n]
    Alphas al = {{0, 0, 0, 0}};
    for (int a = 0; a < n_alphas; a++) al.v[a] = alphas[a];
    k_alpha_powers<<<1, 32, 0, st>>>(al, n_alphas, first_power, nc, pw);
    // the range-check gate has a fused form (one pass, nothing but the folded columns written)
    rc = ux_eval_filtered_fused(params[0], params[1], params[2], params[3], params[4], n, wires, is, pw, F, out, os, st);
    if (rc <= 0) return rc;
    if ((rc = p2e_eval_ux_dev(params[0], params[1], params[2], params[3], params[4], n, wires, is, cons, n, (void*)st))) return rc;
    if (n_alphas == 1) k_filter_combine<1><<<grid_for(n), TPB, 0, st>>>(n, nc, cons, n, pw, F, out, os);
    else if (n_alphas == 2) k_filter_combine<2><<<grid_for(n), TPB, 0, st>>>(n, nc, cons, n, pw, F, out, os);
    else k_filter_combine<4><<<grid_for(n), TPB, 0, st>>>(n, nc, cons, n, pw, F, out, os);
    P2E_LAUNCH_CHECK();
    return 0;
  }
  AlphaPow17 P;
  for (int a = 0; a < MAX_ALPHAS; a++) {
    const u64 al = a < n_alphas ? alphas[a] : 0;
    u64 sc = a < n_alphas ? 1 : 0, b = al;
    for (u64 e = first_power; e; e >>= 1) {
      if (e & 1) sc = gl_mul(sc, b);
      b = gl_mul(b, b);
    }
    P.scale[a] = sc;
    u64 p = a < n_alphas ? 1 : 0, ma = 0;
    for (int c = 0; c < 2 * NL - 1; c++) {
      P.p[a][c] = p;
      if (gate == P2E_GATE_MUL_NONNATIVE && c < NL)
        ma = gl_add(ma, gl_mul(p, params[0] == P2E_SECP256K1_BASE ? ModP::m(c) : ModN::m(c)));
      p = gl_mul(p, al);
    }
    P.ma[a] = ma;
    P.kb[a] = 0;
    if (gate == P2E_GATE_CHECK_SUM && a < n_alphas) {
      u64 g16 = 0;
      for (int c = 0; c < 2 * NL - 2; c++) g16 = gl_add(g16, P.p[a][c]);
      P.ma[a] = gl_sub(al, 1ull << LB);
      P.kb[a] = gl_mul(g16, 1ull << 33);
    }
  }
  // instantiated for 1, 2 and 4 challenges (3 runs as 4 with a zero power row; plonky2's standard config has 2)
#define P2E_FC_LAUNCH(NA)                                                                                  \
  do {                                                                                                     \
    if (gate == P2E_GATE_CHECK_SUM)                                                                        \
      k_eval_check_sum_fc<NA><<<grid_for(n), TPB, 0, st>>>(n, wires, is, F, P, out, os);                   \
    else                                                                                                   \
      k_eval_mul_nonnative_fc<NA><<<grid_for(n), TPB, 0, st>>>(n, wires, is, F, P, out, os);               \
  } while (0)
  if (n_alphas == 1) P2E_FC_LAUNCH(1);
  else if (n_alphas == 2) P2E_FC_LAUNCH(2);
  else P2E_FC_LAUNCH(4);
#undef P2E_FC_LAUNCH
  P2E_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// ---- host-pointer entry points: H2D, kernel, D2H, synchronise (stride == n)
struct HostIn {
  const void* ptr;
  size_t rows;
};

template <class Launch>
static int run_host(size_t n, const HostIn* ins, int n_in, void* out, size_t out_rows, uint32_t* err, Launch launch) {
  P2E_CHECK_READY();
  if (n == 0) return 0;
  P2E_HOST_LOCK();
  cudaStream_t st = ctx().stream;
  size_t in_rows = 0;
  for (int k = 0; k < n_in; k++) in_rows += ins[k].rows;
  void *din = nullptr, *dout = nullptr, *derr = nullptr;
  int rc;
  if ((rc = get_scratch(0, in_rows * n * 8, &din))) return rc;
  if ((rc = get_scratch(1, out_rows * n * 8, &dout))) return rc;
  if ((rc = get_scratch(2, n * 4, &derr))) return rc;
  const uint64_t* dptr[8];
  size_t off = 0;
  for (int k = 0; k < n_in; k++) {
    dptr[k] = (const uint64_t*)din + off;
    if (ins[k].rows)
      P2E_CUDA(cudaMemcpyAsync((uint64_t*)din + off, ins[k].ptr, ins[k].rows * n * 8, cudaMemcpyHostToDevice, st));
    off += ins[k].rows * n;
  }
  if ((rc = launch(dptr, (uint64_t*)dout, (uint32_t*)derr, (void*)st))) return rc;
  P2E_CUDA(cudaMemcpyAsync(out, dout, out_rows * n * 8, cudaMemcpyDeviceToHost, st));
  if (err) P2E_CUDA(cudaMemcpyAsync(err, derr, n * 4, cudaMemcpyDeviceToHost, st));
  P2E_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" {

int p2e_mul_nonnative_witness(int modulus, size_t n, const uint64_t* x, const uint64_t* y, uint64_t* out,
                              uint32_t* err) {
  HostIn ins[2] = {{x, NL}, {y, NL}};
  return run_host(n, ins, 2, out, P2E_MUL_OUT_ROWS, err, [&](const uint64_t** d, uint64_t* o, uint32_t* e, void* s) {
    return p2e_mul_nonnative_witness_dev(modulus, n, d[0], d[1], n, o, n, e, s);
  });
}

int p2e_check_sum_witness(size_t n, const uint64_t* a, uint64_t* b, uint32_t* err) {
  HostIn ins[1] = {{a, 2 * NL - 1}};
  return run_host(n, ins, 1, b, 2 * NL - 2, err, [&](const uint64_t** d, uint64_t* o, uint32_t* e, void* s) {
    return p2e_check_sum_witness_dev(n, d[0], n, o, n, e, s);
  });
}

int p2e_add_nonnative_witness(int modulus, size_t n, const uint64_t* a, int al, const uint64_t* b, int bl,
                              uint64_t* out, uint32_t* err) {
  if (bad_limbs(al) || bad_limbs(bl)) return -1;
  HostIn ins[2] = {{a, (size_t)al}, {b, (size_t)bl}};
  return run_host(n, ins, 2, out, NL + 1, err, [&](const uint64_t** d, uint64_t* o, uint32_t* e, void* s) {
    return p2e_add_nonnative_witness_dev(modulus, n, d[0], al, d[1], bl, n, o, n, e, s);
  });
}

int p2e_sub_nonnative_witness(int modulus, size_t n, const uint64_t* a, int al, const uint64_t* b, int bl,
                              uint64_t* out, uint32_t* err) {
  if (bad_limbs(al) || bad_limbs(bl)) return -1;
  HostIn ins[2] = {{a, (size_t)al}, {b, (size_t)bl}};
  return run_host(n, ins, 2, out, NL + 1, err, [&](const uint64_t** d, uint64_t* o, uint32_t* e, void* s) {
    return p2e_sub_nonnative_witness_dev(modulus, n, d[0], al, d[1], bl, n, o, n, e, s);
  });
}

int p2e_add_many_nonnative_witness(int modulus, size_t n, int k, const uint64_t* summands, uint64_t* out,
                                   uint32_t* err) {
  if (k < 2) return set_error_msg("add_many needs k >= 2 summands");
  HostIn ins[1] = {{summands, (size_t)k * NL}};
  return run_host(n, ins, 1, out, NL + 1, err, [&](const uint64_t** d, uint64_t* o, uint32_t* e, void* s) {
    return p2e_add_many_nonnative_witness_dev(modulus, n, k, d[0], n, o, n, e, s);
  });
}

int p2e_inv_nonnative_witness(int modulus, size_t n, const uint64_t* x, uint64_t* out, uint32_t* err) {
  HostIn ins[1] = {{x, NL}};
  return run_host(n, ins, 1, out, 2 * NL, err, [&](const uint64_t** d, uint64_t* o, uint32_t* e, void* s) {
    return p2e_inv_nonnative_witness_dev(modulus, n, d[0], n, o, n, e, s);
  });
}

int p2e_glv_decompose_witness(size_t n, const uint64_t* k, uint64_t* out, uint32_t* err) {
  HostIn ins[1] = {{k, NL}};
  return run_host(n, ins, 1, out, 12, err, [&](const uint64_t** d, uint64_t* o, uint32_t* e, void* s) {
    return p2e_glv_decompose_witness_dev(n, d[0], n, o, n, e, s);
  });
}

int p2e_biguint_div_rem_witness(size_t n, const uint64_t* a, int al, const uint64_t* b, int bl, uint64_t* out,
                                uint32_t* err) {
  if (al < 0 || al > DR_MAX || bl < 1 || bl > DR_MAX) return set_error_msg("div_rem: limb counts out of range");
  const int div_limbs = (bl > al + 1) ? 0 : al - bl + 1;
  HostIn ins[2] = {{a, (size_t)al}, {b, (size_t)bl}};
  return run_host(n, ins, 2, out, (size_t)(div_limbs + bl), err,
                  [&](const uint64_t** d, uint64_t* o, uint32_t* e, void* s) {
                    return p2e_biguint_div_rem_witness_dev(n, d[0], al, d[1], bl, n, o, n, e, s);
                  });
}

int p2e_eval_mul_nonnative(int modulus, size_t n, const uint64_t* wires, uint64_t* out) {
  HostIn ins[1] = {{wires, 6 * NL - 1}};
  return run_host(n, ins, 1, out, 2 * NL - 1, nullptr, [&](const uint64_t** d, uint64_t* o, uint32_t*, void* s) {
    return p2e_eval_mul_nonnative_dev(modulus, n, d[0], n, o, n, s);
  });
}

int p2e_eval_check_sum(size_t n, const uint64_t* wires, uint64_t* out) {
  HostIn ins[1] = {{wires, 4 * NL - 3}};
  return run_host(n, ins, 1, out, 2 * NL - 1, nullptr, [&](const uint64_t** d, uint64_t* o, uint32_t*, void* s) {
    return p2e_eval_check_sum_dev(n, d[0], n, o, n, s);
  });
}

int p2e_eval_filtered_combined(int gate, const int* params, int gate_index, int group_start, int group_end,
                               int many_selectors, size_t n, const uint64_t* wires, const uint64_t* selector,
                               const uint64_t* alphas, int n_alphas, uint64_t first_power, int accumulate, uint64_t* out) {
  P2E_CHECK_READY();
  int nw = 0, nc = 0, rc;
  if ((rc = gate_shape(gate, params, &nw, &nc))) return rc;
  if (n_alphas < 1 || n_alphas > MAX_ALPHAS) return set_error_msg("eval_filtered: n_alphas must be 1..4");
  if (n == 0) return 0;
  P2E_HOST_LOCK();
  cudaStream_t st = ctx().stream;
  void *din = nullptr, *dout = nullptr, *dscr = nullptr;
  if ((rc = get_scratch(0, ((size_t)nw + 1) * n * 8, &din))) return rc;
  if ((rc = get_scratch(1, (size_t)n_alphas * n * 8, &dout))) return rc;
  if ((rc = get_scratch(3, p2e_eval_filtered_scratch_bytes(gate, params, n), &dscr))) return rc;
  u64* dsel = (u64*)din + (size_t)nw * n;
  P2E_CUDA(cudaMemcpyAsync(din, wires, (size_t)nw * n * 8, cudaMemcpyHostToDevice, st));
  if (selector) P2E_CUDA(cudaMemcpyAsync(dsel, selector, n * 8, cudaMemcpyHostToDevice, st));
  if (accumulate) P2E_CUDA(cudaMemcpyAsync(dout, out, (size_t)n_alphas * n * 8, cudaMemcpyHostToDevice, st));
  if ((rc = p2e_eval_filtered_combined_dev(gate, params, gate_index, group_start, group_end, many_selectors, n,
                                           (const uint64_t*)din, n, selector ? dsel : nullptr, alphas, n_alphas,
                                           first_power, accumulate, (uint64_t*)dout, n, (uint64_t*)dscr, (void*)st)))
    return rc;
  P2E_CUDA(cudaMemcpyAsync(out, dout, (size_t)n_alphas * n * 8, cudaMemcpyDeviceToHost, st));
  P2E_CUDA(cudaStreamSynchronize(st));
  return 0;
}

double p2e_bench_imad(int wide, int iters) {
  if (ensure_ready()) return -1.0;
  P2E_HOST_LOCK();
  Context& c = ctx();
  void* sink;
  if (get_scratch(3, 64, &sink)) return -1.0;
  const int blocks = c.sm_count * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0, c.stream);
    if (wide)
      k_imad_bench<true><<<blocks, threads, 0, c.stream>>>(iters, 12345u + rep, (u64*)sink);
    else
      k_imad_bench<false><<<blocks, threads, 0, c.stream>>>(iters, 12345u + rep, (u64*)sink);
    cudaEventRecord(e1, c.stream);
    if (cudaEventSynchronize(e1) != cudaSuccess) return -2.0;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  double ops = (double)blocks * threads * 8.0 * (double)iters;
  return ops / (best * 1e-3) / 1e9;
}

}  // extern "C"
