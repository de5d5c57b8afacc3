#include "../plonky2_ecdsa_b200/csrc/p2e_f8.cuh"
using namespace p2e;
struct W9 { u32 v[9]; };
extern "C" __global__ void k_addw(const W9* a, const W9* b, W9* r, u32* ov) { W9 x = a[threadIdx.x], y = b[threadIdx.x], z; ov[threadIdx.x] = add_witness_core<ModP>(x.v, y.v, z.v); r[threadIdx.x] = z; }
extern "C" __global__ void k_subw(const W9* a, const W9* b, W9* r, u32* ov) { W9 x = a[threadIdx.x], y = b[threadIdx.x], z; ov[threadIdx.x] = sub_witness_core<ModP>(x.v, y.v, z.v); r[threadIdx.x] = z; }
extern "C" __global__ void k_invk(const W9* a, const F8* b, W9* r, u32* e) {
  W9 xc = a[threadIdx.x]; F8 inv8 = b[threadIdx.x];
  u32 x8[8], t[16], r8[8], q8[8], div[9];
  f8_from_l29(xc.v, x8);
  f8_mul_wide(x8, inv8.v, t);
  u32 bad = f8_divmod16(t, r8, q8);
  bad |= r8[0] ^ 1u;
  for (int l = 1; l < 8; l++) bad |= r8[l];
  f8_to_l29(q8, div);
  W9 d; for (int l=0;l<9;l++) d.v[l]=div[l];
  r[threadIdx.x] = d; e[threadIdx.x] = bad;
}
extern "C" __global__ void k_divmod16(const u32* a, F8* r, F8* q) { u32 t[16]; for (int i=0;i<16;i++) t[i]=a[i*32+threadIdx.x]; F8 rr, qq; u32 b = f8_divmod16(t, rr.v, qq.v); rr.v[0]^=b; r[threadIdx.x]=rr; q[threadIdx.x]=qq; }
