// p2e_fold.cuh -- pieces shared by the fused "selector filter + challenge fold" kernels (p2e_kernels.cu, p2e_ux.cu):
// un-reduced 96-bit accumulation of 64 x 64-bit products, its reduction, and plonky2's selector filter.
#pragma once
#include "p2e_field.cuh"

namespace p2e {

struct Acc3 {
  u32 w0, w1, w2;
};
__device__ __forceinline__ void mac3(Acc3& a, u32 x, u32 y) {
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;"
      : "+r"(a.w0), "+r"(a.w1), "+r"(a.w2)
      : "r"(x), "r"(y));
}
// (w0 + w1 2^32 + w2 2^64 + w3 2^96 + w4 2^128) mod P, w4 < 2^8, canonical result.
// 2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32 (mod P)  =>  V = (w1:w0) + (w2 << 32) - (w2 + w3 + (w4 << 32))
__device__ __forceinline__ u64 gl_reduce5(u32 w0, u32 w1, u32 w2, u32 w3, u32 w4) {
  u64 a = ((u64)w1 << 32) | w0;
  u64 b = (u64)w2 << 32;
  u64 c = (u64)w2 + (u64)w3 + ((u64)w4 << 32);       // < 2^41
  u64 t = a - c;
  if (a < c) t -= GL_EPS;                            // + P
  u64 r = t + b;
  if (r < b) r += GL_EPS;                            // - 2^64 + eps
  if (r >= GL_P) r -= GL_P;
  return r;
}

// (A + B*2^32 + C*2^64 + add) mod P for 96-bit accumulators A, B, C
__device__ __forceinline__ u64 reduce_abc(const Acc3& A, const Acc3& B, const Acc3& C, u64 add) {
  u64 t = (u64)A.w0 + (u32)add;
  u32 W0 = (u32)t;
  t = (t >> 32) + A.w1 + B.w0 + (add >> 32);
  u32 W1 = (u32)t;
  t = (t >> 32) + A.w2 + B.w1 + C.w0;
  u32 W2 = (u32)t;
  t = (t >> 32) + B.w2 + C.w1;
  u32 W3 = (u32)t;
  u32 W4 = (u32)(t >> 32) + C.w2;
  return gl_reduce5(W0, W1, W2, W3, W4);
}


constexpr int MAX_ALPHAS = 4;
struct FilterArgs {
  const u64* selector;       // [n] selector constants of the rows, or nullptr (filter = 1)
  int gate_index, group_start, group_end, many_selectors;
  int n_alphas, accumulate;
};

__device__ __forceinline__ u64 compute_filter(const FilterArgs& F, size_t i) {
  if (!F.selector) return 1;
  const u64 s = F.selector[i];
  u64 f = 1;
  for (int j = F.group_start; j < F.group_end; j++)
    if (j != F.gate_index) f = gl_mul(f, gl_sub((u64)j, s));
  if (F.many_selectors) f = gl_mul(f, gl_sub(0xFFFFFFFFull, s));
  return f;
}


// fused filter + fold of a plonky2_ux gate (p2e_ux.cu); returns 1 when the gate has no fused form
int ux_eval_filtered_fused(int gate, int bits, int lb, int count, int na, size_t n, const u64* wires, size_t is, const u64* pw,
                           const FilterArgs& F, u64* out, size_t os, cudaStream_t st);

}  // namespace p2e
