// Device helpers shared by the attention kernels (attention.cu: all keys of a head resident, T <= 272;
// attention_long.cu: tiled keys with online softmax for larger crops).
#pragma once

#include "common.cuh"

namespace fp {
namespace attn {

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma_f2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t saddr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t saddr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
// D (16x8, fp32) += A (16x16 bf16, row) * B (16x8 bf16, col); used only for the <= 8 tail query rows
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// TMEM -> registers, 8 columns
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ unsigned long long add_f2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// 2^y for two values on the FMA / integer pipes (no MUFU): y is clamped to >= -125, split into round(y) + f with
// f in [-0.5, 0.5] (magic-number rounding), 2^f by a degree-4 minimax polynomial (max relative error 2.7e-6, i.e. 1/700
// of a bf16 ulp -- ex2.approx itself is good to 2^-22), and round(y) is added into the exponent field.  Used for a fixed
// subset of the softmax exponentials so that the 16-lane MUFU is not the only unit working (the FlashAttention-4 trick).
__device__ __forceinline__ void ex2_poly_x2(float y0, float y1, float& e0, float& e1) {
  const unsigned long long y = pack_f2(fmaxf(y0, -125.0f), fmaxf(y1, -125.0f));
  const unsigned long long t = add_f2(y, pack_f2(12582912.0f, 12582912.0f));            // 1.5 * 2^23: low bits = round(y)
  const unsigned long long n = add_f2(t, pack_f2(-12582912.0f, -12582912.0f));
  const unsigned long long f = fma_f2(n, pack_f2(-1.0f, -1.0f), y);
  unsigned long long q = fma_f2(pack_f2(0.009570101276040077f, 0.009570101276040077f), f,
                                pack_f2(0.05591785907745361f, 0.05591785907745361f));
  q = fma_f2(q, f, pack_f2(0.240247443318367f, 0.240247443318367f));
  q = fma_f2(q, f, pack_f2(0.6931217908859253f, 0.6931217908859253f));
  q = fma_f2(q, f, pack_f2(0.9999992847442627f, 0.9999992847442627f));
  float t0, t1, q0, q1;
  unpack_f2(t, t0, t1);
  unpack_f2(q, q0, q1);
  e0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  e1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

// Which of the 16 (even, odd) column pairs of a 32-logit group take the polynomial instead of MUFU.
#ifndef FP_ATTN_POLY_MASK
#define FP_ATTN_POLY_MASK 0x0000u
#endif

// exponentials of 32 logits -> 16 packed bf16x2 words; returns the fp32 sum of the unrounded values.
// MASKED: columns >= valid are forced to 0 (only the group that straddles T takes this path).
// Packed f32x2 FMA / ADD: 2.5 issue slots per element (FFMA2 .5, MUFU 1, FADD2 .5, F2FP .5) instead of 3.5.
template <bool MASKED, unsigned POLY = FP_ATTN_POLY_MASK>
__device__ __forceinline__ float exp_group(const uint32_t (&v)[32], float sl2, float msl, int valid,
                                           uint32_t (&packed)[16]) {
  const unsigned long long sl2_2 = pack_f2(sl2, sl2), nmsl_2 = pack_f2(-msl, -msl);
  unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};  // four independent (even, odd) chains
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float y0, y1;
    unpack_f2(fma_f2(pack_f2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sl2_2, nmsl_2), y0, y1);
    float e0, e1;
    if ((POLY >> (j >> 1)) & 1u) {
      ex2_poly_x2(y0, y1, e0, e1);
    } else {
      e0 = ex2(y0);
      e1 = ex2(y1);
    }
    if (MASKED) {
      e0 = j < valid ? e0 : 0.f;
      e1 = j + 1 < valid ? e1 : 0.f;
    }
    acc[(j >> 1) & 3] = add_f2(acc[(j >> 1) & 3], pack_f2(e0, e1));
    packed[j >> 1] = pack_bf16x2(e0, e1);
  }
  float s0, s1, s2, s3;
  unpack_f2(add_f2(acc[0], acc[1]), s0, s1);
  unpack_f2(add_f2(acc[2], acc[3]), s2, s3);
  return (s0 + s1) + (s2 + s3);
}

template <bool MASKED>
__device__ __forceinline__ float max_group(const uint32_t (&v)[32], int valid, float m) {
  float part[4] = {m, m, m, m};
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float x = (!MASKED || j < valid) ? __uint_as_float(v[j]) : -INFINITY;
    part[j & 3] = fmaxf(part[j & 3], x);
  }
  return fmaxf(fmaxf(part[0], part[1]), fmaxf(part[2], part[3]));
}


}  // namespace attn
}  // namespace fp
