#include <stdint.h>
typedef uint32_t u32; typedef uint64_t u64;
struct F8 { u32 v[8]; };

#define ROW_FIRST(e0,e1,e2,e3,e4,e5,e6,e7, a0,a1,a2,a3, b) \
  asm("mul.lo.u32 %0, %8, %12;\n\t mul.hi.u32 %1, %8, %12;\n\t" \
      "mul.lo.u32 %2, %9, %12;\n\t mul.hi.u32 %3, %9, %12;\n\t" \
      "mul.lo.u32 %4, %10, %12;\n\t mul.hi.u32 %5, %10, %12;\n\t" \
      "mul.lo.u32 %6, %11, %12;\n\t mul.hi.u32 %7, %11, %12;" \
      : "=r"(e0),"=r"(e1),"=r"(e2),"=r"(e3),"=r"(e4),"=r"(e5),"=r"(e6),"=r"(e7) : "r"(a0),"r"(a1),"r"(a2),"r"(a3),"r"(b))
#define ROW_NC(e0,e1,e2,e3,e4,e5,e6,e7, a0,a1,a2,a3, b) \
  asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t madc.hi.cc.u32 %1, %8, %12, %1;\n\t" \
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t madc.hi.cc.u32 %3, %9, %12, %3;\n\t" \
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t madc.hi.cc.u32 %5, %10, %12, %5;\n\t" \
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t madc.hi.u32 %7, %11, %12, %7;" \
      : "+r"(e0),"+r"(e1),"+r"(e2),"+r"(e3),"+r"(e4),"+r"(e5),"+r"(e6),"+r"(e7) : "r"(a0),"r"(a1),"r"(a2),"r"(a3),"r"(b))
#define ROW_C(e0,e1,e2,e3,e4,e5,e6,e7,e8, a0,a1,a2,a3, b) \
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t madc.hi.cc.u32 %1, %9, %13, %1;\n\t" \
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t madc.hi.cc.u32 %3, %10, %13, %3;\n\t" \
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t madc.hi.cc.u32 %5, %11, %13, %5;\n\t" \
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t madc.hi.cc.u32 %7, %12, %13, %7;\n\t" \
      "addc.u32 %8, 0, 0;" \
      : "+r"(e0),"+r"(e1),"+r"(e2),"+r"(e3),"+r"(e4),"+r"(e5),"+r"(e6),"+r"(e7),"=r"(e8) : "r"(a0),"r"(a1),"r"(a2),"r"(a3),"r"(b))

__device__ __forceinline__ void mul_wide(const u32 (&a)[8], const u32 (&b)[8], u32 (&t)[16]) {
  u32 e[16], o[16];
#pragma unroll
  for (int i = 8; i < 16; i++) { e[i] = 0; o[i] = 0; }
  ROW_FIRST(e[0],e[1],e[2],e[3],e[4],e[5],e[6],e[7], a[0],a[2],a[4],a[6], b[0]);
  ROW_FIRST(o[0],o[1],o[2],o[3],o[4],o[5],o[6],o[7], a[1],a[3],a[5],a[7], b[0]);
#pragma unroll
  for (int j = 1; j < 8; j++) {
    if (j & 1) {
      ROW_C(o[j-1],o[j],o[j+1],o[j+2],o[j+3],o[j+4],o[j+5],o[j+6],o[j+7], a[0],a[2],a[4],a[6], b[j]);
      ROW_NC(e[j+1],e[j+2],e[j+3],e[j+4],e[j+5],e[j+6],e[j+7],e[j+8], a[1],a[3],a[5],a[7], b[j]);
    } else {
      ROW_C(e[j],e[j+1],e[j+2],e[j+3],e[j+4],e[j+5],e[j+6],e[j+7],e[j+8], a[0],a[2],a[4],a[6], b[j]);
      ROW_NC(o[j],o[j+1],o[j+2],o[j+3],o[j+4],o[j+5],o[j+6],o[j+7], a[1],a[3],a[5],a[7], b[j]);
    }
  }
  t[0] = e[0];
  asm("add.cc.u32 %0, %1, %2;" : "=r"(t[1]) : "r"(e[1]), "r"(o[0]));
#pragma unroll
  for (int k = 2; k < 15; k++) asm("addc.cc.u32 %0, %1, %2;" : "=r"(t[k]) : "r"(e[k]), "r"(o[k-1]));
  asm("addc.u32 %0, %1, %2;" : "=r"(t[15]) : "r"(e[15]), "r"(o[14]));
}

// t[16] -> r[8] weakly reduced (< 2^256), p = 2^256 - 2^32 - 977
__device__ __forceinline__ void reduce16(const u32 (&t)[16], u32 (&r)[8]) {
  const u32 C = 977;
  u32 e[9], o[9];
  // even words of H times 977 added to L (aligned pairs), carry -> e[8]
  asm("mad.lo.cc.u32 %0, %9, %13, %14;\n\t madc.hi.cc.u32 %1, %9, %13, %15;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %16;\n\t madc.hi.cc.u32 %3, %10, %13, %17;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %18;\n\t madc.hi.cc.u32 %5, %11, %13, %19;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %20;\n\t madc.hi.cc.u32 %7, %12, %13, %21;\n\t"
      "addc.u32 %8, 0, 0;"
      : "=r"(e[0]),"=r"(e[1]),"=r"(e[2]),"=r"(e[3]),"=r"(e[4]),"=r"(e[5]),"=r"(e[6]),"=r"(e[7]),"=r"(e[8])
      : "r"(t[8]),"r"(t[10]),"r"(t[12]),"r"(t[14]),"r"(C),
        "r"(t[0]),"r"(t[1]),"r"(t[2]),"r"(t[3]),"r"(t[4]),"r"(t[5]),"r"(t[6]),"r"(t[7]));
  // odd words of H times 977, plus H itself (the 2^32 term), in the frame shifted by one word, carry -> o[8]
  asm("mad.lo.cc.u32 %0, %9, %13, %14;\n\t madc.hi.cc.u32 %1, %9, %13, %15;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %16;\n\t madc.hi.cc.u32 %3, %10, %13, %17;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %18;\n\t madc.hi.cc.u32 %5, %11, %13, %19;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %20;\n\t madc.hi.cc.u32 %7, %12, %13, %21;\n\t"
      "addc.u32 %8, 0, 0;"
      : "=r"(o[0]),"=r"(o[1]),"=r"(o[2]),"=r"(o[3]),"=r"(o[4]),"=r"(o[5]),"=r"(o[6]),"=r"(o[7]),"=r"(o[8])
      : "r"(t[9]),"r"(t[11]),"r"(t[13]),"r"(t[15]),"r"(C),
        "r"(t[8]),"r"(t[9]),"r"(t[10]),"r"(t[11]),"r"(t[12]),"r"(t[13]),"r"(t[14]),"r"(t[15]));
  // R = e + (o << 32): words 0..9
  u32 R[10];
  R[0] = e[0];
  asm("add.cc.u32 %0, %1, %2;" : "=r"(R[1]) : "r"(e[1]), "r"(o[0]));
#pragma unroll
  for (int k = 2; k < 9; k++) asm("addc.cc.u32 %0, %1, %2;" : "=r"(R[k]) : "r"(e[k]), "r"(o[k-1]));
  asm("addc.u32 %0, %1, 0;" : "=r"(R[9]) : "r"(o[8]));
  // second fold: tt = R[8] + 2^32 R[9] (R[9] <= 1):  A = tt * (2^32 + 977) as three words
  u32 A0, A1, A2;
  asm("mul.lo.u32 %0, %3, %5;\n\t mul.hi.u32 %1, %3, %5;\n\t"
      "add.cc.u32 %1, %1, %3;\n\t addc.u32 %2, %4, 0;\n\t"       // + R8 << 32, + R9 << 64
      "mad.lo.cc.u32 %1, %4, %5, %1;\n\t addc.u32 %2, %2, 0;"     // + R9 * 977 << 32
      : "=&r"(A0), "=&r"(A1), "=&r"(A2) : "r"(R[8]), "r"(R[9]), "r"(C));
  u32 c8;
  asm("add.cc.u32 %0, %9, %17;\n\t addc.cc.u32 %1, %10, %18;\n\t addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, 0;\n\t addc.cc.u32 %4, %13, 0;\n\t addc.cc.u32 %5, %14, 0;\n\t"
      "addc.cc.u32 %6, %15, 0;\n\t addc.cc.u32 %7, %16, 0;\n\t addc.u32 %8, 0, 0;"
      : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(c8)
      : "r"(R[0]),"r"(R[1]),"r"(R[2]),"r"(R[3]),"r"(R[4]),"r"(R[5]),"r"(R[6]),"r"(R[7]),"r"(A0),"r"(A1),"r"(A2));
  // third fold (c8 = 1 only when r < 2^67): cannot carry beyond word 2
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t addc.cc.u32 %1, %1, %3;\n\t addc.u32 %2, %2, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]) : "r"(c8), "r"(C));
}

extern "C" __global__ void k_f8mul(F8* a, F8* b, F8* r) {
  F8 x = a[threadIdx.x], y = b[threadIdx.x], z; u32 t[16];
  mul_wide(x.v, y.v, t); reduce16(t, z.v); r[threadIdx.x] = z; }
