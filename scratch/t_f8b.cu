#include "../plonky2_ecdsa_b200/csrc/p2e_f8.cuh"
using namespace p2e;
extern "C" __global__ void k_mul(F8* a, F8* b, F8* r) { F8 x = a[threadIdx.x], y = b[threadIdx.x], z; f8_mul(x.v, y.v, z.v); r[threadIdx.x] = z; }
extern "C" __global__ void k_sqr(F8* a, F8* r) { F8 x = a[threadIdx.x], z; f8_sqr(x.v, z.v); r[threadIdx.x] = z; }
extern "C" __global__ void k_add(F8* a, F8* b, F8* r) { F8 x = a[threadIdx.x], y = b[threadIdx.x], z; f8_add(x.v, y.v, z.v); r[threadIdx.x] = z; }
extern "C" __global__ void k_sub(F8* a, F8* b, F8* r) { F8 x = a[threadIdx.x], y = b[threadIdx.x], z; f8_sub(x.v, y.v, z.v); r[threadIdx.x] = z; }
extern "C" __global__ void k_ms3(F8* a, F8* r) { F8 x = a[threadIdx.x], z; f8_mul_small<3>(x.v, z.v); r[threadIdx.x] = z; }
extern "C" __global__ void k_ms8(F8* a, F8* r) { F8 x = a[threadIdx.x], z; f8_mul_small<8>(x.v, z.v); r[threadIdx.x] = z; }
extern "C" __global__ void k_canon29(F8* a, u32* r) { F8 x = a[threadIdx.x]; u32 l[9]; f8_to_canon_l29(x.v, l); for (int i=0;i<9;i++) r[i*32+threadIdx.x] = l[i]; }
extern "C" __global__ void k_from29(u32* a, F8* r) { u32 l[9]; for (int i=0;i<9;i++) l[i] = a[i*32+threadIdx.x]; F8 z; f8_from_l29(l, z.v); r[threadIdx.x] = z; }
