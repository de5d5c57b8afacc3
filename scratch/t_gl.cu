#include "../plonky2_ecdsa_b200/csrc/p2e_field.cuh"
using namespace p2e;
// current
__device__ __forceinline__ u64 rp3_cur(u64 v) {
  const u64 u = glm(v, gls(v, 7));
  const u64 z = gla(u, 6);
  const u64 zz = glm(z, z);
  u64 s4 = gl_red_lazy(z << 2, z >> 62);
  return glm(gls(zz, 36), gla(zz, glc(s4)));
}
// candidate: explicit 32-bit limb multiply + fused reduction
__device__ __forceinline__ u64 mulred(u64 a, u64 b) {
  const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  u32 x0, x1, x2, x3;
  // x = a*b as four words: p00 + (p01 + p10) << 32 + p11 << 64
  asm("{\n\t.reg .u32 m0, m1, m2, t;\n\t"
      "mul.lo.u32 %0, %4, %6;\n\tmul.hi.u32 %1, %4, %6;\n\t"          // p00
      "mul.lo.u32 %2, %5, %7;\n\tmul.hi.u32 %3, %5, %7;\n\t"          // p11
      "mad.lo.cc.u32 %1, %4, %7, %1;\n\tmadc.hi.cc.u32 %2, %4, %7, %2;\n\taddc.u32 %3, %3, 0;\n\t"   // + p01 << 32
      "mad.lo.cc.u32 %1, %5, %6, %1;\n\tmadc.hi.cc.u32 %2, %5, %6, %2;\n\taddc.u32 %3, %3, 0;\n\t"   // + p10 << 32
      "}" : "=&r"(x0), "=&r"(x1), "=&r"(x2), "=&r"(x3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  // reduce: x0 + x1 2^32 + x2 2^64 + x3 2^96 = (x1:x0) + x2 (2^32 - 1) - x3
  u64 lo = ((u64)x1 << 32) | x0;
  return gl_red_lazy(lo, ((u64)x3 << 32) | x2);
}
__device__ __forceinline__ u64 rp3_new(u64 v) {
  const u64 u = mulred(v, gls(v, 7));
  const u64 z = gla(u, 6);
  const u64 zz = mulred(z, z);
  u64 s4 = gl_red_lazy(z << 2, z >> 62);
  return mulred(gls(zz, 36), gla(zz, glc(s4)));
}
extern "C" __global__ void k_cur(const u64* a, u64* o) { o[threadIdx.x] = glc(rp3_cur(a[threadIdx.x])); }
extern "C" __global__ void k_new(const u64* a, u64* o) { o[threadIdx.x] = glc(rp3_new(a[threadIdx.x])); }
extern "C" __global__ void k_glm(const u64* a, const u64* b, u64* o) { o[threadIdx.x] = glm(a[threadIdx.x], b[threadIdx.x]); }
extern "C" __global__ void k_mulred(const u64* a, const u64* b, u64* o) { o[threadIdx.x] = mulred(a[threadIdx.x], b[threadIdx.x]); }
