// Row helpers shared by the score and retrieval kernels: 16-byte row loads (lane l owns elements c*256 + l*8 + j),
// the fixed-order fp32 reductions the CPU oracle restates (oracle/score.py _lane_reduce) and the bf16 F.normalize.
//
// Reduction order over a D-long row (the contract oracle/score.py restates): lane l keeps TWO partial sums, one over
// its even elements j = 0,2,4,6 and one over its odd elements j = 1,3,5,7 (chunks c outer, words w = j/2 inner), adds
// them (even + odd), then the xor butterfly 16,8,4,2,1 across the lanes.  The two partial sums are the two halves of a
// packed fp32x2 register: a bf16x2 word unpacks into one (lo, hi) pair and one FFMA2 advances both sums -- half the
// FMA instructions and half the dependent-chain length of a single running sum.  Products of two bf16 values are exact
// in fp32, so fma(a, b, acc) == fadd(acc, fmul(a, b)) bit for bit.
#pragma once

#include "common.cuh"

namespace fp {
namespace rowops {

constexpr int MAX_CHUNKS = 4;  // D <= 1024

// ---- packed fp32x2 (sm_100 FFMA2 / FMUL2: two IEEE fp32 lanes per issue slot) ----------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// bf16x2 word -> its two values as an fp32 pair (lo = even element, hi = odd element)
__device__ __forceinline__ f32x2 word2(uint32_t v) { return pack2(bf16lo(v), bf16hi(v)); }
// sum of the two halves (even + odd)
__device__ __forceinline__ float hsum2(f32x2 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return __fadd_rn(lo, hi);
}

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
  f32x2 acc = pack2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < chunks) {
      const uint32_t w[4] = {u[c].x, u[c].y, u[c].z, u[c].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const f32x2 x = word2(w[i]);
        acc = fma2(x, x, acc);
      }
    }
  const float ss = warp_sum(hsum2(acc));
  const float nrm = bf16_round(__fsqrt_rn(ss));
  const float eps = bf16_round(1e-12f);
  return fmaxf(nrm, eps);
}

// fixed-order dot product of two rows held as packed bf16 (order: see the header)
__device__ __forceinline__ float row_dot(const uint4 (&a)[MAX_CHUNKS], const uint4 (&b)[MAX_CHUNKS], int chunks) {
  f32x2 acc = pack2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < chunks) {
      const uint32_t wa[4] = {a[c].x, a[c].y, a[c].z, a[c].w};
      const uint32_t wb[4] = {b[c].x, b[c].y, b[c].z, b[c].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) acc = fma2(word2(wa[i]), word2(wb[i]), acc);
    }
  return warp_sum(hsum2(acc));
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
