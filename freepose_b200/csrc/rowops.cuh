// Row helpers shared by the score and retrieval kernels: 16-byte row loads (lane l owns elements c*256 + l*8 + j),
// the fixed-order fp32 reductions the CPU oracle restates (oracle/score.py _lane_reduce) and the bf16 F.normalize.
#pragma once

#include "common.cuh"

namespace fp {
namespace rowops {

constexpr int MAX_CHUNKS = 4;  // D <= 1024

__device__ __forceinline__ void load_row(const bf16* row, int lane, int chunks, uint4 (&u)[MAX_CHUNKS]) {
  const uint4* p = reinterpret_cast<const uint4*>(row);
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < chunks) u[c] = p[c * 32 + lane];
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}

// bf16(sqrt(sum x^2)) clamped below by bf16(eps), as float
__device__ __forceinline__ float row_norm(const uint4 (&u)[MAX_CHUNKS], int chunks) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < chunks) {
      float f[8];
      unpack8(u[c], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(f[j], f[j], acc);  // == fadd(acc, fmul(x, x)): the square of a bf16 is exact in fp32
    }
  acc = warp_sum(acc);
  const float nrm = bf16_round(__fsqrt_rn(acc));
  const float eps = bf16_round(1e-12f);
  return fmaxf(nrm, eps);
}

// fixed-order dot product of two rows held as packed bf16: acc = acc + a*b (a*b is exact in fp32), then the butterfly
__device__ __forceinline__ float row_dot(const uint4 (&a)[MAX_CHUNKS], const uint4 (&b)[MAX_CHUNKS], int chunks) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < chunks) {
      float af[8], bf[8];
      unpack8(a[c], af);
      unpack8(b[c], bf);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(af[j], bf[j], acc);  // == fadd(acc, fmul(a, b)): the product is exact
    }
  return warp_sum(acc);
}

// x / max(bf16(||x||), eps) rounded to bf16, element-wise on the packed row (F.normalize on a bf16 tensor)
__device__ __forceinline__ void normalise_row(uint4 (&u)[MAX_CHUNKS], int chunks) {
  const float nrm = row_norm(u, chunks);
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < chunks) {
      float f[8];
      unpack8(u[c], f);
      u[c].x = pack_bf16x2(__fdiv_rn(f[0], nrm), __fdiv_rn(f[1], nrm));
      u[c].y = pack_bf16x2(__fdiv_rn(f[2], nrm), __fdiv_rn(f[3], nrm));
      u[c].z = pack_bf16x2(__fdiv_rn(f[4], nrm), __fdiv_rn(f[5], nrm));
      u[c].w = pack_bf16x2(__fdiv_rn(f[6], nrm), __fdiv_rn(f[7], nrm));
    }
}

}  // namespace rowops
}  // namespace fp
