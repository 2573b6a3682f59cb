// LayerNorm of one row by one warp (contract: oracle/vit.py contract_layernorm, y = bf16(((x - mean) * rstd) * w + b),
// fp32 two-pass statistics).  Shared by layernorm_kernel (elementwise.cu) and by the LayerNorm warps fused into the
// residual GEMMs (gemm.cu), so that both produce the same bits: lane l owns elements c*256 + l*8 + j; every sum runs in
// two independent fp32x2 pair accumulators (pairs i even / i odd), halves added (even + odd elements), then the xor
// butterfly over the lanes.
#pragma once

#include "common.cuh"
#include "rowops.cuh"

namespace fp {
namespace lnrow {

using rowops::f32x2;

// v: the row's elements as fp32 pairs (NP = D / 64 per lane), normalised in place to (x - mean) * rstd
template <int NP>
__device__ __forceinline__ void normalise_pairs(f32x2 (&v)[NP], float inv_d, float eps) {
  using namespace rowops;
  f32x2 s0 = pack2(0.f, 0.f), s1 = s0;
#pragma unroll
  for (int i = 0; i < NP; i += 2) { s0 = add2(s0, v[i]); s1 = add2(s1, v[i + 1]); }
  const float mean = warp_sum(hsum2(add2(s0, s1))) * inv_d;
  const f32x2 nmean = pack2(-mean, -mean);
  f32x2 q0 = pack2(0.f, 0.f), q1 = q0;
#pragma unroll
  for (int i = 0; i < NP; i += 2) {
    v[i] = add2(v[i], nmean);
    v[i + 1] = add2(v[i + 1], nmean);
    q0 = fma2(v[i], v[i], q0);
    q1 = fma2(v[i + 1], v[i + 1], q1);
  }
  const float rstd = rsqrtf(warp_sum(hsum2(add2(q0, q1))) * inv_d + eps);
  const f32x2 rstd2 = pack2(rstd, rstd);
#pragma unroll
  for (int i = 0; i < NP; ++i) v[i] = mul2(v[i], rstd2);
}

// y = bf16(v * w + b) for one bf16x2 word
__device__ __forceinline__ uint32_t affine_word(f32x2 v, f32x2 w, f32x2 b) {
  float y0, y1;
  rowops::unpack2(rowops::fma2(v, w, b), y0, y1);
  return pack_bf16x2(y0, y1);
}

}  // namespace lnrow
}  // namespace fp
