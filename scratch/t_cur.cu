#include "../plonky2_ecdsa_b200/csrc/p2e_field.cuh"
using namespace p2e;
struct I9 { i32 v[NL]; };
struct W9 { u32 v[NL]; };
extern "C" __global__ void k_lzmul(I9* a, I9* b, I9* r) { I9 x = a[threadIdx.x], y = b[threadIdx.x], z; lz_mul(x.v, y.v, z.v); r[threadIdx.x] = z; }
extern "C" __global__ void k_lzsqr(I9* a, I9* r) { I9 x = a[threadIdx.x], z; lz_sqr(x.v, z.v); r[threadIdx.x] = z; }
extern "C" __global__ void k_lzsub(I9* a, I9* b, I9* r) { I9 x = a[threadIdx.x], y = b[threadIdx.x], z; lz_sub(x.v, y.v, z.v); r[threadIdx.x] = z; }
extern "C" __global__ void k_lzcanon(I9* a, W9* r) { I9 x = a[threadIdx.x]; W9 z; lz_canon(x.v, z.v); r[threadIdx.x] = z; }
extern "C" __global__ void k_mulw(W9* a, W9* b, u64* o) { W9 x = a[threadIdx.x], y = b[threadIdx.x]; u32 r[NL], q[NL]; u64 cs[17], bb[16];
  u32 e = mul_witness_core<ModP>(x.v, y.v, r, q, cs, bb);
  for (int i = 0; i < 9; i++) { o[i*32+threadIdx.x] = r[i]; o[(9+i)*32+threadIdx.x] = q[i]; }
  for (int i = 0; i < 17; i++) o[(18+i)*32+threadIdx.x] = cs[i];
  for (int i = 0; i < 16; i++) o[(35+i)*32+threadIdx.x] = bb[i];
  o[51*32+threadIdx.x] = e; }
extern "C" __global__ void k_divmod(u64* a, W9* r) { u64 col[17]; for (int i = 0; i < 17; i++) col[i] = a[i*32+threadIdx.x]; W9 z; u32 q10[10]; divmod_cols<ModP,true>(col, z.v, q10); for (int i=0;i<9;i++) z.v[i]^=q10[i]; r[threadIdx.x] = z; }
