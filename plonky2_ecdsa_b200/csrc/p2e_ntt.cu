// p2e_ntt.cu -- Goldilocks NTT / coset low-degree extension: the step that FEEDS the gate evaluators
// (SURVEY 8f-3).  plonky2's quotient stage evaluates every gate over the coset LDE of the wire
// polynomials (PolynomialValues::lde_onto_coset: ifft -> zero-pad by 2^rate_bits -> coset_fft with
// shift g = F::coset_shift()); the output here is wire-major [poly][lde_row], i.e. exactly the matrix
// p2e_eval_* reads.
//
// Conventions (plonky2 field crate, out of tree, restated; the constants are checked against their
// defining identities in the tests): omega_n = POWER_OF_TWO_GENERATOR^(2^(32 - log n)) with
// POWER_OF_TWO_GENERATOR = 7^((p-1)/2^32) = 1753635133440165772; coset shift g = 7
// (MULTIPLICATIVE_GROUP_GENERATOR); natural order in and out: X[k] = sum_i x[i] omega_n^(i k).
//
// Design: four-step NTT, n = n1 * n2, two HBM passes.  Pass A: each CTA stages T adjacent COLUMNS of
// the n1 x n2 matrix (T * 8-byte contiguous segments per row) in shared memory, runs T radix-2 DIF
// NTTs of length n1 there (three stages per shared-memory round trip, 8 elements per thread in registers), multiplies by omega_n^(i2 k1) and stores; pass B does the same for T
// adjacent ROWS (length n2) and writes X[k1 + n1 k2] (T-element contiguous segments).  Lines are
// padded to an odd stride so both the line-fastest and the element-fastest access patterns are
// bank-conflict free; the bit reversal of the DIF output is folded into the shared-memory read of the
// store loop.  Zero-padded inputs (LDE) are not read.  8192-element tiles (64 KB) keep 3 CTAs per SM so
// one CTA's butterflies overlap another's loads / stores (default tile: 4096 elements, 32 KB, 6 CTAs per SM).
#include <map>
#include <mutex>
#include <vector>

#include "../../include/p2e.h"
#include "p2e_common.h"
#include "p2e_field.cuh"

namespace p2e {

constexpr int NTT_THREADS = 256;
constexpr int NTT_TILE = 8192;            // elements staged per CTA
constexpr u64 GL_TWO_ADIC_GEN = 1753635133440165772ull;   // order 2^32
constexpr u64 GL_COSET_SHIFT = 7;

struct NttPass {
  const u64* in;
  u64* out;
  size_t in_stride, out_stride;   // between polynomials
  size_t n_in;                    // elements of a polynomial present in `in` (the rest are zero)
  int log_n1, log_n2;             // n = n1 * n2, input index i = n2 * i1 + i2, output index k = k1 + n1 * k2
  int log_t;                      // lines per CTA
  const u64* tw;                  // DIF twiddles of this pass's line length L: w_L^e, e < L / 2
  const u64 *cross_lo, *cross_hi; // pass A: w_n^e = cross_lo[e & (n2 - 1)] * cross_hi[e >> log_n2]
  const u64 *scale_lo, *scale_hi; // pass B: optional out[k] *= scale_lo[k & (n2 - 1)] * scale_hi[k >> log_n2]
  const u64 *cross_full;          // pass A: w_n^e, e < n, as ONE table (n <= 2^22: one multiplication instead of two)
  const u64 *scale_full;          // pass B: premul * base^k, k < n, likewise
};

__device__ __forceinline__ unsigned brev(unsigned v, int bits) { return __brev(v) >> (32 - bits); }

// G consecutive radix-2 DIF stages (s .. s+G-1) of T lines, 2^G elements per thread held in registers:
// one shared-memory round trip, one barrier and one set of index arithmetic per G stages.
// Element q of a thread sits at j = hi * (L >> s) + q * (L >> (s + G)) + lo; stage s+u pairs the elements
// whose q differ in bit G-1-u, with twiddle w_L^(pos << (s+u)), pos = (q mod 2^(G-1-u)) * (L >> (s+G)) + lo.
template <int G>
__device__ __forceinline__ void dif_round(u64* sm, int log_l, int log_t, int s, const u64* __restrict__ tw) {
  constexpr int R = 1 << G;
  const int L = 1 << log_l, ls = L + 1;
  const int log_step = log_l - s - G;                 // distance between a thread's elements
  const int step = 1 << log_step;
  const int per_line = L >> G;                         // threads' work items per line
  const int total = per_line << log_t;
  for (int b = threadIdx.x; b < total; b += NTT_THREADS) {
    const int t = b >> (log_l - G), w = b & (per_line - 1);
    const int lo = w & (step - 1), hi = w >> log_step;
    u64* base = sm + t * ls + (hi << (log_l - s)) + lo;
    u64 v[R];
#pragma unroll
    for (int q = 0; q < R; q++) v[q] = base[q << log_step];
#pragma unroll
    for (int u = 0; u < G; u++) {
      const int bit = G - 1 - u;                       // q bit paired by this stage
      const bool last = (s + u == log_l - 1);          // final stage of the NTT: twiddle 1
#pragma unroll
      for (int q = 0; q < R; q++) {
        if (q & (1 << bit)) continue;
        const int q1 = q | (1 << bit);
        const u64 a = v[q], c = v[q1];
        v[q] = glc(gla(a, c));
        const u64 d = gls(a, c);
        if (last) {
          v[q1] = glc(d);
        } else {
          const int pos = ((q & ((1 << bit) - 1)) << log_step) + lo;
          v[q1] = glc(glm(d, __ldg(tw + ((size_t)pos << (s + u)))));
        }
      }
    }
#pragma unroll
    for (int q = 0; q < R; q++) base[q << log_step] = v[q];
  }
  __syncthreads();
}

// any * 2^A -> any
template <int A>
__device__ __forceinline__ u64 gl_shl(u64 v) { return gl_red_lazy(v << A, v >> (64 - A)); }

// Three DIF stages (s .. s+2) as ONE radix-8 step: an 8-point DFT whose twiddles are powers of w_8, followed by one
// general twiddle per output.  In this field w_8 = w_L^(L/8) = -2^24 for plonky2's roots (2^96 = -1, so 2^24 has
// order 8; inverse transform: w_8^-1 = 2^72), i.e. the 5 non-trivial twiddles inside the DFT are shifts (the sign
// is absorbed by swapping the operands of the subtraction) and only 7 of the 12 multiplications of three radix-2
// stages remain -- none at all in the last step of a line, whose outer twiddles are 1.
// Output r of the 8-point DFT lands where three radix-2 stages would have put it (position bitrev3(r)); its outer
// twiddle is w_m^(lo r) = w_L^((lo r) << s), m = L >> s.  tw: the FULL table w_L^e, e < L.
template <bool INV>
__device__ __forceinline__ void dif_round8(u64* sm, int log_l, int log_t, int s, const u64* __restrict__ tw) {
  const int L = 1 << log_l, ls = L + 1;
  const int log_step = log_l - s - 3;
  const int step = 1 << log_step;
  const int per_line = L >> 3;
  const int total = per_line << log_t;
  for (int b = threadIdx.x; b < total; b += NTT_THREADS) {
    const int t = b >> (log_l - 3), w = b & (per_line - 1);
    const int lo = w & (step - 1), hi = w >> log_step;
    u64* base = sm + t * ls + (hi << (log_l - s)) + lo;
    u64 v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = base[q << log_step];
    // Canonicalisation is needed only where a value becomes the SECOND operand of an addition / subtraction
    // (gla / gls take any u64 as their first operand): positions with bit 1 set after stage A (plus the odd ones for
    // the inverse transform, whose stage-B subtraction is swapped), odd positions after stage B.
    // stage A: (q, q + 4), twiddle w_8^q
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const u64 a = v[q], c = v[q + 4];
      const bool canon = (q & 2) || (INV && (q & 1));
      const u64 sum = gla(a, c);
      v[q] = canon ? glc(sum) : sum;
      u64 d;
      if (q == 0) d = gls(a, c);
      else if (q == 1) d = INV ? gl_shl<48>(gl_shl<24>(gls(a, c))) : gl_shl<24>(gls(c, a));      // 2^72 | -2^24
      else if (q == 2) d = INV ? gl_shl<48>(gls(c, a)) : gl_shl<48>(gls(a, c));                   // -2^48 | 2^48
      else d = INV ? gl_shl<24>(gls(a, c)) : gl_shl<48>(gl_shl<24>(gls(c, a)));                   // 2^24 | -2^72
      v[q + 4] = canon ? glc(d) : d;
    }
    // stage B: (q, q + 2) inside each half, twiddle w_4^q = w_8^(2q)
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const u64 a = v[h + q], c = v[h + q + 2];
        const u64 sum = gla(a, c);
        v[h + q] = (q & 1) ? glc(sum) : sum;
        u64 d;
        if (q == 0) d = gls(a, c);
        else d = INV ? gl_shl<48>(gls(c, a)) : gl_shl<48>(gls(a, c));
        v[h + q + 2] = (q & 1) ? glc(d) : d;
      }
    }
    // stage C: (q, q + 1)
#pragma unroll
    for (int h = 0; h < 8; h += 2) {
      const u64 a = v[h], c = v[h + 1];
      const u64 sum = gla(a, c), d = gls(a, c);
      v[h] = (h == 0 || log_step == 0) ? glc(sum) : sum;        // the others are canonicalised after their twiddle
      v[h + 1] = log_step == 0 ? glc(d) : d;
    }
    if (log_step > 0) {
#pragma unroll
      for (int qp = 1; qp < 8; qp++) {
        const int r = ((qp & 1) << 2) | (qp & 2) | (qp >> 2);
        v[qp] = glc(glm(v[qp], __ldg(tw + ((size_t)(lo * r) << s))));
      }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) base[q << log_step] = v[q];
  }
  __syncthreads();
}

// T DIF NTTs of length L = 1 << log_l on sm[t * (L + 1) + j]; output in bit-reversed order.  The (log_l mod 3)
// leading stages run as radix-2 / radix-4 steps with general twiddles, the rest as radix-8 steps.
template <bool INV>
__device__ __forceinline__ void dif_lines(u64* sm, int log_l, int log_t, const u64* __restrict__ tw) {
  const int rem = log_l % 3;
  if (rem == 1) dif_round<1>(sm, log_l, log_t, 0, tw);
  else if (rem == 2) dif_round<2>(sm, log_l, log_t, 0, tw);
  for (int s = rem; s < log_l; s += 3) dif_round8<INV>(sm, log_l, log_t, s, tw);
}

// pass A: columns
template <bool INV>
__global__ void __launch_bounds__(NTT_THREADS) k_ntt_cols(NttPass P) {
  extern __shared__ u64 sm[];
  const int log_l = P.log_n1, L = 1 << log_l, ls = L + 1, T = 1 << P.log_t;
  const size_t n2 = (size_t)1 << P.log_n2;
  const size_t c0 = (size_t)blockIdx.x << P.log_t;
  const u64* in = P.in + (size_t)blockIdx.y * P.in_stride;
  u64* out = P.out + (size_t)blockIdx.y * P.out_stride;
  for (int idx = threadIdx.x; idx < (L << P.log_t); idx += NTT_THREADS) {
    const int t = idx & (T - 1), j = idx >> P.log_t;
    const size_t gi = n2 * j + c0 + t;
    sm[t * ls + j] = gi < P.n_in ? in[gi] : 0;
  }
  __syncthreads();
  dif_lines<INV>(sm, log_l, P.log_t, P.tw);
  for (int idx = threadIdx.x; idx < (L << P.log_t); idx += NTT_THREADS) {
    const int t = idx & (T - 1), k1 = idx >> P.log_t;
    u64 v = sm[t * ls + brev((unsigned)k1, log_l)];
    const size_t e = (c0 + t) * (size_t)k1;
    if (P.cross_full) v = glm(v, __ldg(P.cross_full + e));
    else v = glm(v, glm(__ldg(P.cross_lo + (e & (n2 - 1))), __ldg(P.cross_hi + (e >> P.log_n2))));
    out[n2 * k1 + c0 + t] = glc(v);
  }
}

// pass B: rows
template <bool INV>
__global__ void __launch_bounds__(NTT_THREADS) k_ntt_rows(NttPass P) {
  extern __shared__ u64 sm[];
  const int log_l = P.log_n2, L = 1 << log_l, ls = L + 1, T = 1 << P.log_t;
  const size_t n1 = (size_t)1 << P.log_n1, n2 = (size_t)L;
  const size_t r0 = (size_t)blockIdx.x << P.log_t;
  const u64* in = P.in + (size_t)blockIdx.y * P.in_stride;
  u64* out = P.out + (size_t)blockIdx.y * P.out_stride;
  for (int idx = threadIdx.x; idx < (L << P.log_t); idx += NTT_THREADS) {
    const int j = idx & (L - 1), t = idx >> log_l;
    sm[t * ls + j] = in[n2 * (r0 + t) + j];
  }
  __syncthreads();
  dif_lines<INV>(sm, log_l, P.log_t, P.tw);
  for (int idx = threadIdx.x; idx < (L << P.log_t); idx += NTT_THREADS) {
    const int t = idx & (T - 1), k2 = idx >> P.log_t;
    u64 v = sm[t * ls + brev((unsigned)k2, log_l)];
    const size_t k = r0 + t + n1 * (size_t)k2;
    if (P.scale_full) v = glc(glm(v, __ldg(P.scale_full + k)));
    else if (P.scale_lo) v = glc(glm(v, glm(__ldg(P.scale_lo + (k & (n2 - 1))), __ldg(P.scale_hi + (k >> P.log_n2)))));
    out[k] = v;
  }
}

// ------------------------------------------------------------------------------------------
// host: twiddle tables (built once per size / direction, kept on the device)
static u64 h_mul(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) % GL_P); }
static u64 h_pow(u64 b, u64 e) {
  u64 r = 1;
  for (; e; e >>= 1) {
    if (e & 1) r = h_mul(r, b);
    b = h_mul(b, b);
  }
  return r;
}
static u64 h_inv(u64 x) { return h_pow(x, GL_P - 2); }
static u64 h_root(int log_n) { return h_pow(GL_TWO_ADIC_GEN, 1ull << (32 - log_n)); }

struct DevTable {
  u64* ptr = nullptr;
};
// per-device state: twiddle / scaling tables and the "function attributes set" flag (function attributes are per device)
struct NttState {
  std::mutex mu;
  std::map<std::vector<u64>, DevTable> tables;   // key: {base, count, premul}
  bool attr_set = false;
};
static NttState& ntt_state() {
  Context& c = ctx();
  static std::mutex create_mu;
  std::lock_guard<std::mutex> lock(create_mu);
  if (!c.ntt) c.ntt = new NttState();
  return *(NttState*)c.ntt;
}

// device table of premul * base^e, e < count
static int power_table(u64 base, size_t count, u64 premul, const u64** out) {
  NttState& S = ntt_state();
  std::lock_guard<std::mutex> lock(S.mu);
  std::vector<u64> key = {base, (u64)count, premul};
  auto it = S.tables.find(key);
  if (it == S.tables.end()) {
    std::vector<u64> h(count ? count : 1);
    u64 v = premul;
    for (size_t e = 0; e < count; e++) {
      h[e] = v;
      v = h_mul(v, base);
    }
    DevTable t;
    P2E_CUDA(cudaMalloc(&t.ptr, h.size() * 8));
    P2E_CUDA(cudaMemcpy(t.ptr, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    it = S.tables.emplace(key, t).first;
  }
  *out = it->second.ptr;
  return 0;
}

void ntt_release_device(Context& c) {
  if (!c.ntt) return;
  NttState* S = (NttState*)c.ntt;
  for (auto& kv : S->tables)
    if (kv.second.ptr) cudaFree(kv.second.ptr);
  delete S;
  c.ntt = nullptr;
}

static int log_tile() {
  int v = (int)option(OPT_NTT_LOG_TILE);   // measured on B200 (LDE 2^17 -> 2^20 x 135): 2^13 4.33 ms, 2^12 4.16 ms, 2^11 4.50 ms
  if (v < 10 || v > 13) v = 12;
  return v;
}

static int ilog_t(int log_l, int log_other) {
  int lt = log_tile() - log_l;    // elements staged per CTA (<= NTT_TILE = 2^13)
  if (lt > 4) lt = 4;             // 16 lines = 128-byte segments
  if (lt > log_other) lt = log_other;
  if (lt < 0) lt = 0;
  return lt;
}

// One NTT of size 2^log_n over n_polys polynomials: in -> tmp (pass A) -> out (pass B).
// root: the primitive n-th root to use (omega or omega^-1); scale_base / scale_premul: out[k] *= premul * base^k
// (0 = no scaling); n_in: input elements present (zero-padded beyond).
static int ntt_run(int log_n, bool inverse, u64 scale_base, u64 scale_premul, int n_polys, const u64* in, size_t in_stride,
                   size_t n_in, u64* tmp, size_t tmp_stride, u64* out, size_t out_stride, cudaStream_t st) {
  const int log_n1 = (log_n + 1) / 2, log_n2 = log_n / 2;
  const size_t n = (size_t)1 << log_n, n1 = (size_t)1 << log_n1, n2 = (size_t)1 << log_n2;
  const u64 root = inverse ? h_inv(h_root(log_n)) : h_root(log_n);
  // the radix-8 steps hard-code w_8 = -2^24 (inverse: 2^72): true for plonky2's roots, checked once
  if (log_n >= 3) {
    const u64 w8 = h_pow(root, n / 8);
    if (w8 != (inverse ? h_pow(2, 72) : GL_P - h_pow(2, 24))) return set_error_msg("ntt: unexpected primitive 8th root of unity");
  }
  NttPass A;
  A.log_n1 = log_n1;
  A.log_n2 = log_n2;
  int rc;
  const u64 *tw1, *tw2, *clo = nullptr, *chi = nullptr, *slo = nullptr, *shi = nullptr, *cfull = nullptr, *sfull = nullptr;
  const bool full = log_n <= 22;          // one-table twiddles up to 32 MB
  // full tables of w_n1 = root^n2 and w_n2 = root^n1 (the radix-8 steps index up to 7/8 of a turn)
  if ((rc = power_table(h_pow(root, n2), n1, 1, &tw1))) return rc;
  if ((rc = power_table(h_pow(root, n1), n2, 1, &tw2))) return rc;
  if (full) {
    if ((rc = power_table(root, n, 1, &cfull))) return rc;
  } else {
    if ((rc = power_table(root, n2, 1, &clo))) return rc;
    chi = tw1;
  }
  if (scale_base) {
    if (full) {
      if ((rc = power_table(scale_base, n, scale_premul, &sfull))) return rc;
    } else {
      if ((rc = power_table(scale_base, n2, 1, &slo))) return rc;
      if ((rc = power_table(h_pow(scale_base, n2), n1, scale_premul, &shi))) return rc;
    }
  }
  NttState& NS = ntt_state();
  const size_t smem_max = (size_t)(NTT_TILE + 16) * 8;
  if (!NS.attr_set) {
    P2E_CUDA(cudaFuncSetAttribute(k_ntt_cols<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    P2E_CUDA(cudaFuncSetAttribute(k_ntt_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    P2E_CUDA(cudaFuncSetAttribute(k_ntt_cols<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    P2E_CUDA(cudaFuncSetAttribute(k_ntt_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    NS.attr_set = true;
  }
  // pass A
  A.in = in; A.out = tmp; A.in_stride = in_stride; A.out_stride = tmp_stride; A.n_in = n_in;
  A.log_t = ilog_t(log_n1, log_n2);
  A.tw = tw1; A.cross_lo = clo; A.cross_hi = chi; A.scale_lo = nullptr; A.scale_hi = nullptr;
  A.cross_full = cfull; A.scale_full = nullptr;
  {
    const size_t smem = (((size_t)n1 + 1) << A.log_t) * 8;
    const dim3 grid((unsigned)(n2 >> A.log_t), (unsigned)n_polys);
    if (inverse) k_ntt_cols<true><<<grid, NTT_THREADS, smem, st>>>(A);
    else k_ntt_cols<false><<<grid, NTT_THREADS, smem, st>>>(A);
  }
  // pass B
  NttPass B = A;
  B.in = tmp; B.out = out; B.in_stride = tmp_stride; B.out_stride = out_stride; B.n_in = n;
  B.log_t = ilog_t(log_n2, log_n1);
  B.tw = tw2; B.scale_lo = slo; B.scale_hi = shi; B.scale_full = sfull;
  {
    const size_t smem = (((size_t)n2 + 1) << B.log_t) * 8;
    const dim3 grid((unsigned)(n1 >> B.log_t), (unsigned)n_polys);
    if (inverse) k_ntt_rows<true><<<grid, NTT_THREADS, smem, st>>>(B);
    else k_ntt_rows<false><<<grid, NTT_THREADS, smem, st>>>(B);
  }
  P2E_CUDA(cudaGetLastError());
  return 0;
}

static int bad_ntt_args(int log_n, int n_polys) {
  if (log_n < 2 || log_n > 24) return set_error_msg("ntt: log_n must be 2..24");
  if (n_polys < 0) return set_error_msg("ntt: n_polys < 0");
  return 0;
}

}  // namespace p2e

using namespace p2e;

extern "C" {

size_t p2e_ntt_scratch_bytes(int log_n, int rate_bits, int n_polys) {
  // tmp for the size-n pass + coefficients + tmp for the size-(n << rate_bits) pass
  const size_t n = (size_t)1 << log_n, m = n << rate_bits;
  return (size_t)n_polys * (rate_bits ? 2 * n + m : n) * 8;
}

int p2e_ntt_dev(int log_n, int inverse, int n_polys, const uint64_t* in, size_t in_stride, uint64_t* out,
                size_t out_stride, uint64_t* scratch, void* stream) {
  int rc = ensure_ready();
  if (rc) return rc;
  if (bad_ntt_args(log_n, n_polys)) return -1;
  if (n_polys == 0) return 0;
  const size_t n = (size_t)1 << log_n;
  if (in_stride < n || out_stride < n) return set_error_msg("ntt: stride < n");
  if (inverse) return ntt_run(log_n, true, 1, h_inv((u64)n % GL_P), n_polys, in, in_stride, n, scratch, n, out, out_stride, pick_stream(stream));
  return ntt_run(log_n, false, 0, 0, n_polys, in, in_stride, n, scratch, n, out, out_stride, pick_stream(stream));
}

int p2e_coset_lde_dev(int log_n, int rate_bits, int n_polys, const uint64_t* values, size_t in_stride, uint64_t* out,
                      size_t out_stride, uint64_t* scratch, void* stream) {
  int rc = ensure_ready();
  if (rc) return rc;
  if (bad_ntt_args(log_n, n_polys) || bad_ntt_args(log_n + rate_bits, n_polys)) return -1;
  if (rate_bits < 1 || rate_bits > 8) return set_error_msg("coset_lde: rate_bits must be 1..8");
  if (n_polys == 0) return 0;
  const size_t n = (size_t)1 << log_n, m = n << rate_bits;
  if (in_stride < n || out_stride < m) return set_error_msg("coset_lde: stride too small");
  cudaStream_t st = pick_stream(stream);
  u64* tmp_n = scratch;                               // [n_polys][n]
  u64* coeffs = scratch + (size_t)n_polys * n;        // [n_polys][n]: c_k * g^k
  u64* tmp_m = coeffs + (size_t)n_polys * n;          // [n_polys][m]
  // ifft, with the coset shift folded into the output scaling: c_k * n^-1 * g^k
  if ((rc = ntt_run(log_n, true, GL_COSET_SHIFT, h_inv((u64)n % GL_P), n_polys, values, in_stride, n, tmp_n, n,
                    coeffs, n, st)))
    return rc;
  // fft of the zero-padded coefficients over the larger subgroup
  return ntt_run(log_n + rate_bits, false, 0, 0, n_polys, coeffs, n, n, tmp_m, m, out, out_stride, st);
}

// host-pointer forms
static int ntt_host(int log_in, int log_out, int rate_bits, int n_polys, const uint64_t* in, uint64_t* out,
                    int (*launch)(void*, const uint64_t*, uint64_t*, uint64_t*, void*), void* closure) {
  int rc = ensure_ready();
  if (rc) return rc;
  if (n_polys <= 0) return n_polys < 0 ? set_error_msg("ntt: n_polys < 0") : 0;
  P2E_HOST_LOCK();
  const size_t n_in = (size_t)1 << log_in, n_out = (size_t)1 << log_out;
  cudaStream_t st = ctx().stream;
  void *din = nullptr, *dout = nullptr, *dscr = nullptr;
  if ((rc = get_scratch(0, (size_t)n_polys * n_in * 8, &din))) return rc;
  if ((rc = get_scratch(1, (size_t)n_polys * n_out * 8, &dout))) return rc;
  if ((rc = get_scratch(3, p2e_ntt_scratch_bytes(log_in, rate_bits, n_polys), &dscr))) return rc;
  P2E_CUDA(cudaMemcpyAsync(din, in, (size_t)n_polys * n_in * 8, cudaMemcpyHostToDevice, st));
  if ((rc = launch(closure, (const uint64_t*)din, (uint64_t*)dout, (uint64_t*)dscr, (void*)st))) return rc;
  P2E_CUDA(cudaMemcpyAsync(out, dout, (size_t)n_polys * n_out * 8, cudaMemcpyDeviceToHost, st));
  P2E_CUDA(cudaStreamSynchronize(st));
  return 0;
}

struct NttCall { int log_n, flag, n_polys; };

int p2e_ntt(int log_n, int inverse, int n_polys, const uint64_t* in, uint64_t* out) {
  if (bad_ntt_args(log_n, n_polys)) return -1;
  NttCall c = {log_n, inverse, n_polys};
  return ntt_host(log_n, log_n, 0, n_polys, in, out, [](void* p, const uint64_t* d, uint64_t* o, uint64_t* s, void* st) {
    NttCall* c = (NttCall*)p;
    const size_t n = (size_t)1 << c->log_n;
    return p2e_ntt_dev(c->log_n, c->flag, c->n_polys, d, n, o, n, s, st);
  }, &c);
}

int p2e_coset_lde(int log_n, int rate_bits, int n_polys, const uint64_t* values, uint64_t* out) {
  if (bad_ntt_args(log_n, n_polys) || bad_ntt_args(log_n + rate_bits, n_polys)) return -1;
  if (rate_bits < 1 || rate_bits > 8) return set_error_msg("coset_lde: rate_bits must be 1..8");
  NttCall c = {log_n, rate_bits, n_polys};
  return ntt_host(log_n, log_n + rate_bits, rate_bits, n_polys, values, out,
                  [](void* p, const uint64_t* d, uint64_t* o, uint64_t* s, void* st) {
                    NttCall* c = (NttCall*)p;
                    const size_t n = (size_t)1 << c->log_n;
                    return p2e_coset_lde_dev(c->log_n, c->flag, c->n_polys, d, n, o, n << c->flag, s, st);
                  }, &c);
}

}  // extern "C"
