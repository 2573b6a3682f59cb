// Persistent, warp-specialised tcgen05 GEMM for the ViT linear layers:
//
//     C[M, N] = epilogue( A[M, K] (bf16, K contiguous)  x  W[N, K]^T (bf16, K contiguous) ),  fp32 accumulate
//
// One CTA per SM, 128 x 256 output tiles, K consumed in 64-element (128-byte, SWIZZLE_128B) slabs:
//   warp 0        TMA producer      cp.async.bulk.tensor -> 4-stage smem ring (A 16 KB + W 32 KB per stage)
//   warp 1        MMA issuer        one thread issues tcgen05.mma (M128 N256 K16) into TMEM
//   warp 2        TMEM allocator    512 columns = two 128x256 fp32 accumulators (double buffered)
//   warps 4..11   epilogue          tcgen05.ld -> registers -> fused epilogue -> bf16 global stores;
//                                   overlaps the MMA of the next tile through the second accumulator
//
// Epilogues replicate the rounding points of PyTorch-eager bf16 (each ATen op rounds its output):
//   BIAS          out = bf16(acc + b)                                            (attn.qkv)
//   BIAS_GELU     out = bf16(gelu_erf(bf16(acc + b)))                            (mlp.fc1 + act)
//   BIAS_LS_RES   out = bf16(res + bf16(bf16(acc + b) * gamma))                  (attn.proj / mlp.fc2 + LayerScale + residual)
//   PATCH_EMBED   out[b*T + 1 + R + p] = bf16(bf16(acc + b) + pos[1 + p])        (patch-embed conv as GEMM + pos-embed)
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "ln_row.cuh"

namespace fp {

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int NUM_EPI_GROUPS = 3;                   // warps per TMEM lane quarter (column chunks are dealt round-robin)
constexpr int NUM_EPI_WARPS = 4 * NUM_EPI_GROUPS;   // 12: three per SM sub-partition hide the epilogue's ALU/MUFU latency
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int EPI_STAGE_BYTES = NUM_EPI_WARPS * 2 * 3 * 64;  // per warp: bias + gamma for its <= 3 column chunks
constexpr int OUT_STAGE_BYTES = NUM_EPI_WARPS * 32 * 64;     // per warp: one 32-row x 32-column bf16 output tile (TMA store source)
// layout after the operand ring: [barriers 256 B][bias/gamma staging][pad to 512 B][output staging]
constexpr int OUT_STAGE_OFF = (256 + EPI_STAGE_BYTES + 511) / 512 * 512;
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + OUT_STAGE_OFF + OUT_STAGE_BYTES;

struct Params {
  int M, N, K;
  bf16* out;
  int ldo;  // output row stride (elements)
  const bf16* bias;
  const bf16* gamma;
  const bf16* res;  // BIAS_LS_RES: residual [M, ldo] (may alias out); PATCH_EMBED: pos-embed [1 + P, N]
  int patches_per_img;
  int tokens_per_img;
  int token_offset;
  int debug;  // perf experiments only: 1 = skip epilogue stores, 2 = always load tile (0,0)
  // fused LayerNorm of the output rows (gemm2_kernel<EPI_BIAS_LS_RES, true>)
  const bf16* ln_w;
  const bf16* ln_b;
  bf16* ln_out;
  float ln_eps;
  unsigned* ln_counters;   // one word per 128 output rows: arrivals of the epilogue warps that stored a tile of them
  unsigned ln_expect;      // epoch * (epilogue warps * N tiles)
};

// ---- packed fp32x2 arithmetic (sm_100 FFMA2/FMUL2/FADD2: two fp32 lanes per issue slot) -------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 splat2(float c) { return pack2(c, c); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// bf16x2 bits -> the two values as an fp32 pair
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t v) { return pack2(bf16lo(v), bf16hi(v)); }
// round an fp32 pair to bf16 (RN-even) and return it again as fp32 pair (values now bf16-representable)
__device__ __forceinline__ f32x2 round2_bf16(f32x2 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return bf16x2_to_f32x2(pack_bf16x2(lo, hi));
}
__device__ __forceinline__ uint32_t f32x2_to_bf16x2(f32x2 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return pack_bf16x2(lo, hi);
}

// GELU(erf) of two values.  erf by Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7, far below half a bf16 ulp of
// the result) with approximate MUFU rcp / ex2, written so that no select is needed:
//   gelu(u) = u/2 + |u|/2 * erf(|u|/sqrt2),   erf(a) = 1 - (a1 t + ... + a5 t^5) exp(-a^2),  t = 1/(1 + p a)
// ~11.5 issue slots per element instead of ~39 for libdevice erff (the fc1 epilogue has to fit under the MMA time).
__device__ __forceinline__ f32x2 gelu2(f32x2 u) {
  // y = u * sqrt(log2 e / 2)  =>  exp(-(u/sqrt2)^2) = exp2(-y^2);  |u/sqrt2| = |y| / sqrt(log2 e)
  constexpr float kS = 0.84932180028801904f;             // sqrt(log2(e) / 2)
  constexpr float kP = 0.3275911f / 1.2011224087864498f;  // A&S p, rescaled to |y|
  const f32x2 y = mul2(u, splat2(kS));
  float y0, y1;
  unpack2(y, y0, y1);
  const f32x2 ay = pack2(fabsf(y0), fabsf(y1));
  float d0, d1;
  unpack2(fma2(ay, splat2(kP), splat2(1.0f)), d0, d1);
  const f32x2 t = pack2(rcp_fast(d0), rcp_fast(d1));
  // -(a1 t + ... + a5 t^5): coefficients negated so that erf(|z|) = fma(poly, e, 1)
  f32x2 pl = fma2(t, splat2(-1.061405429f), splat2(1.453152027f));
  pl = fma2(pl, t, splat2(-1.421413741f));
  pl = fma2(pl, t, splat2(0.284496736f));
  pl = fma2(pl, t, splat2(-0.254829592f));
  pl = mul2(pl, t);
  float e0, e1;
  unpack2(mul2(y, y), e0, e1);
  const f32x2 e = pack2(ex2_fast(-e0), ex2_fast(-e1));   // the negation folds into the MUFU operand
  const f32x2 erf_abs = fma2(pl, e, splat2(1.0f));       // erf(|u| / sqrt2)
  // gelu(u) = u/2 + |u|/2 * erf(|u|/sqrt2) = (y + |y| * erf_abs) / (2 kS)
  return mul2(fma2(ay, erf_abs, y), splat2(0.5f / kS));
}

// Per-tile operands that do not depend on the accumulator are fetched BEFORE the wait on the accumulator barrier,
// so their global-memory latency is hidden: bias (+ LayerScale gamma) of this warp's column chunks go to a
// warp-private smem staging area (read back as broadcast LDS), the first chunk's residual / pos-embed row segment
// goes to registers.
struct EpiPrefetch {
  uint4 x[4];  // residual (BIAS_LS_RES) or pos-embed (PATCH_EMBED) of the first chunk
};

template <int MODE>
__device__ __forceinline__ void epilogue_prefetch(const Params& p, uint4* stage, int grp, int lane, int row, int n_blk,
                                                  EpiPrefetch& pf) {
  // lanes 0..11: bias (3 chunks x 4 x 16 B), lanes 12..23: gamma
  const int which = lane / 12, l12 = lane - which * 12;
  const int ci = l12 >> 2, part = l12 & 3;
  const int chunk = grp + ci * NUM_EPI_GROUPS;
  if (lane < (MODE == EPI_BIAS_LS_RES ? 24 : 12) && chunk < BN / 32) {
    const bf16* src = (which == 0 ? p.bias : p.gamma) + n_blk * BN + chunk * 32 + part * 8;
    stage[which * 12 + l12] = __ldg(reinterpret_cast<const uint4*>(src));
  }
  if (MODE == EPI_BIAS_LS_RES || MODE == EPI_PATCH_EMBED) {
    const int gcol = n_blk * BN + grp * 32;
    if (row < p.M) {
      const bf16* src;
      if (MODE == EPI_BIAS_LS_RES) {
        src = p.res + size_t(row) * p.ldo + gcol;
      } else {
        const int img = row / p.patches_per_img;
        src = p.res + size_t(1 + row - img * p.patches_per_img) * p.N + gcol;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) pf.x[i] = reinterpret_cast<const uint4*>(src)[i];
    }
  }
  __syncwarp();
}

// Drains this warp's share (TMEM lane quarter q, column chunks grp, grp+G, ...) of one finished 128 x 256
// accumulator: tcgen05.ld -> fused epilogue in packed fp32x2 -> 64-byte bf16 row segments to HBM.
template <int MODE>
__device__ __forceinline__ void epilogue_tile(const Params& p, const CUtensorMap* tmOut, uint32_t tmem_base, int acc,
                                              int q, int grp, int row, int n_blk, const uint4* stage, EpiPrefetch& pf,
                                              uint8_t* out_stage) {
  // Contiguous-row outputs go through a swizzled smem tile and a TMA bulk store (full 32-byte sectors, asynchronous,
  // rows past M clipped by the hardware); the patch-embed scatter keeps direct stores (rows are remapped per image).
  constexpr bool kTmaStore = MODE != EPI_PATCH_EMBED;
  const int lane = threadIdx.x & 31;
  const bool row_ok = row < p.M;
  int out_row = row;
  int pos_row = 0;
  if (MODE == EPI_PATCH_EMBED) {
    const int img = row / p.patches_per_img;
    const int pidx = row - img * p.patches_per_img;
    out_row = img * p.tokens_per_img + p.token_offset + pidx;
    pos_row = 1 + pidx;
  }
  int ci = 0;
#pragma unroll 1
  for (int chunk = grp; chunk < BN / 32; chunk += NUM_EPI_GROUPS, ++ci) {
    const int col0 = chunk * 32;
    const int gcol = n_blk * BN + col0;
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN + col0), r);
    uint4 bv[4], gv[4], xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bv[i] = stage[ci * 4 + i];
      if (MODE == EPI_BIAS_LS_RES) gv[i] = stage[12 + ci * 4 + i];
      if (MODE == EPI_BIAS_LS_RES || MODE == EPI_PATCH_EMBED) xv[i] = pf.x[i];
    }
    // residual / pos-embed of the NEXT chunk: in flight while this chunk is computed
    if ((MODE == EPI_BIAS_LS_RES || MODE == EPI_PATCH_EMBED) && row_ok && chunk + NUM_EPI_GROUPS < BN / 32) {
      const int ncol = gcol + NUM_EPI_GROUPS * 32;
      const bf16* src = MODE == EPI_BIAS_LS_RES ? p.res + size_t(row) * p.ldo + ncol : p.res + size_t(pos_row) * p.N + ncol;
#pragma unroll
      for (int i = 0; i < 4; ++i) pf.x[i] = reinterpret_cast<const uint4*>(src)[i];
    }
    tmem_ld_wait();
    uint4 outv[4];
    if (kTmaStore || row_ok) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t bw[4] = {bv[i].x, bv[i].y, bv[i].z, bv[i].w};
        const uint32_t gw[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
        const uint32_t xw[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // two adjacent columns at a time in packed fp32x2
          f32x2 v = add2(pack2(__uint_as_float(r[i * 8 + j * 2]), __uint_as_float(r[i * 8 + j * 2 + 1])),
                         bf16x2_to_f32x2(bw[j]));
          if (MODE == EPI_BIAS_GELU) {
            v = gelu2(round2_bf16(v));
          } else if (MODE == EPI_BIAS_LS_RES) {
            v = round2_bf16(mul2(round2_bf16(v), bf16x2_to_f32x2(gw[j])));
            v = add2(v, bf16x2_to_f32x2(xw[j]));
          } else if (MODE == EPI_PATCH_EMBED) {
            v = add2(round2_bf16(v), bf16x2_to_f32x2(xw[j]));
          }
          ow[j] = f32x2_to_bf16x2(v);
        }
        outv[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
    if (kTmaStore) {
      // only now must the previous chunk's bulk store have finished reading the staging tile
      tma_store_wait_read();   // (a no-op for the lanes that never committed a bulk group)
      __syncwarp();
      // 64-byte swizzle (CU_TENSOR_MAP_SWIZZLE_64B): 16-byte chunk i of row `lane` sits at chunk i ^ ((lane>>1)&3)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint4*>(out_stage + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = outv[i];
    } else if (row_ok) {
      uint4* optr = reinterpret_cast<uint4*>(p.out + size_t(out_row) * p.ldo + gcol);
#pragma unroll
      for (int i = 0; i < 4; ++i) optr[i] = outv[i];
    }
    if (kTmaStore) {
      fence_proxy_async_smem();
      __syncwarp();
      if (p.debug != 1 && elect_one()) {
        tma_store_2d(tmOut, out_stage, gcol, row - lane);  // box = 32 rows of this warp's lane quarter x 32 columns
        tma_store_commit();
      }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmOut, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the __shared__ array keeps the shared address space: LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full_bar = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]   epilogue -> MMA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m_tiles = (p.M + BM - 1) / BM;
  const int num_n_tiles = p.N / BN;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int kblocks = p.K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n_tiles, n_blk = tile % num_n_tiles;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
          tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(sB + stage * B_STAGE_BYTES, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // (elect.sync, not `lane == 0`: the compiler then knows a single thread is active and emits the UTCHMMA stream
    // back to back instead of wrapping every instruction in an elect / branch loop)
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(sA + stage * A_STAGE_BYTES), 16, 1024);
          const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(sB + stage * B_STAGE_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // +32 bytes per UMMA_K inside the 128-byte swizzle row -> +2 in the (addr >> 4) field
            umma_bf16_ss(d_tmem, a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int grp = (warp - 4) >> 2;   // 0..NUM_EPI_GROUPS-1: which column chunks of the tile this warp drains
    uint4* epi_stage = reinterpret_cast<uint4*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256) + (warp - 4) * 24;
    uint8_t* out_stage = smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + OUT_STAGE_OFF + (warp - 4) * 2048;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n_tiles, n_blk = tile % num_n_tiles;
      const int row = m_blk * BM + q * 32 + lane;
      EpiPrefetch pf;
      epilogue_prefetch<MODE>(p, epi_stage, grp, lane, row, n_blk, pf);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<MODE>(p, &tmOut, tmem_base, acc, q, grp, row, n_blk, epi_stage, pf, out_stage);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (warp >= 4) tma_store_wait_all();  // smem staging must outlive the bulk stores reading it
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm fused into the residual GEMMs (attn.proj / mlp.fc2 -> norm2 / next block's norm1).
// The GEMM's tiles own 256 of the 1024 columns of a row, so the normalisation cannot happen in a tile's epilogue.  Instead
// the kernel carries one extra warpgroup: every epilogue warp, once the bulk stores of its part of a tile have completed,
// adds 1 to the counter of its 128-row block; the LayerNorm warps (4 per CTA, 592 on the chip) take units of 8 rows
// round-robin, wait until the block's counter says that all N tiles of those rows are in memory, read the rows back --
// from L2, they were written microseconds ago -- and write the normalised rows.  The standalone LayerNorm launch (557 MB
// through HBM, 0.1 ms, 44 times per step) disappears: the read hits L2 and the write overlaps the GEMM's tensor work.
// The arithmetic is lnrow::normalise_pairs / affine_word, the same code as layernorm_kernel: identical bits.
// ------------------------------------------------------------------------------------------------
constexpr int LN_WARPS_PER_CTA = 4;
constexpr int LN_UNIT_ROWS = 8;
constexpr int LN_AHEAD = 3;                   // rows in flight per LayerNorm warp
constexpr int LN_SMEM_BYTES = 2 * 1024 * 2;   // gamma | beta as bf16, D <= 1024

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int NCH>
__device__ __forceinline__ void layernorm_warps(const Params& p, uint8_t* ln_smem, int lw, int lane) {
  using namespace rowops;
  constexpr int D = NCH * 256, NP = NCH * 4;
  // gamma / beta -> shared memory (the four LayerNorm warps of this CTA only: named barrier 1)
  uint4* sgb = reinterpret_cast<uint4*>(ln_smem);
  for (int i = lw * 32 + lane; i < 2 * NCH * 32; i += LN_WARPS_PER_CTA * 32)
    sgb[i] = __ldg(reinterpret_cast<const uint4*>(i < NCH * 32 ? p.ln_w : p.ln_b) + (i < NCH * 32 ? i : i - NCH * 32));
  asm volatile("bar.sync 1, %0;" ::"n"(LN_WARPS_PER_CTA * 32) : "memory");
  if (p.debug & 4) return;   // perf experiments: no LayerNorm work at all (WRONG results)
  const long long units = ((long long)p.M + LN_UNIT_ROWS - 1) / LN_UNIT_ROWS;
  const long long stride = (long long)gridDim.x * LN_WARPS_PER_CTA;
  for (long long u = (long long)blockIdx.x * LN_WARPS_PER_CTA + lw; u < units; u += stride) {
    const int row0 = int(u) * LN_UNIT_ROWS;
    const int row1 = min(row0 + LN_UNIT_ROWS, p.M);
    if (lane == 0) {
      const unsigned* c = p.ln_counters + (row0 >> 7);
      unsigned spins = 0;
      while (ld_acquire_gpu(c) < p.ln_expect) {
        __nanosleep(64);
        if (++spins > (1u << 26)) __trap();   // a protocol bug must not hang the device
      }
    }
    __syncwarp();
    // rows come back from L2 (ld.cg: no L1, which may still hold the residual values this SM read before the update);
    // LN_AHEAD rows are in flight per warp: a row costs ~1 us of L2 latency under the GEMM's own traffic, and a warp owes
    // M / 592 rows per launch -- one row at a time made these warps the kernel's critical path
    uint4 ring[LN_AHEAD][NCH];
#pragma unroll
    for (int k = 0; k < LN_AHEAD; ++k)
      if (row0 + k < row1) {
        const uint4* xp = reinterpret_cast<const uint4*>(p.out + size_t(row0 + k) * p.ldo);
#pragma unroll
        for (int c = 0; c < NCH; ++c) ring[k][c] = __ldcg(xp + c * 32 + lane);
      }
#pragma unroll
    for (int k = 0; k < LN_UNIT_ROWS; ++k) {
      const int r = row0 + k;
      if (r >= row1) break;
      f32x2 v[NP];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const uint4 u = ring[k % LN_AHEAD][c];
        v[c * 4 + 0] = word2(u.x); v[c * 4 + 1] = word2(u.y); v[c * 4 + 2] = word2(u.z); v[c * 4 + 3] = word2(u.w);
      }
      if (r + LN_AHEAD < row1) {
        const uint4* xp = reinterpret_cast<const uint4*>(p.out + size_t(r + LN_AHEAD) * p.ldo);
#pragma unroll
        for (int c = 0; c < NCH; ++c) ring[k % LN_AHEAD][c] = __ldcg(xp + c * 32 + lane);
      }
      lnrow::normalise_pairs<NP>(v, 1.0f / D, p.ln_eps);
      uint4* op = reinterpret_cast<uint4*>(p.ln_out + size_t(r) * D);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const uint4 wu = sgb[c * 32 + lane], bu = sgb[NCH * 32 + c * 32 + lane];
        uint4 o;
        o.x = lnrow::affine_word(v[c * 4 + 0], word2(wu.x), word2(bu.x));
        o.y = lnrow::affine_word(v[c * 4 + 1], word2(wu.y), word2(bu.y));
        o.z = lnrow::affine_word(v[c * 4 + 2], word2(wu.z), word2(bu.z));
        o.w = lnrow::affine_word(v[c * 4 + 3], word2(wu.w), word2(bu.w));
        op[c * 32 + lane] = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): the two CTAs of a cluster drive ONE 256 x 256 tcgen05.mma.  Each CTA stages its own
// 128 rows of A and 128 of the 256 W rows per K slab (32 KB/stage instead of 48 KB: 6 stages, and half the operand
// traffic per SM), the leader CTA issues the MMAs, each CTA drains the 128 accumulator rows that live in its own
// TMEM.  Barriers: `full` lives in the leader (both CTAs' TMA bytes land on it); `empty` and `tfull` exist in both
// CTAs and are signalled by multicast tcgen05.commit; `tempty` lives in the leader and is arrived on remotely.
// ------------------------------------------------------------------------------------------------
constexpr int STAGES2 = 6;
constexpr int STAGE2_BYTES = 2 * BM * BK * 2;  // A half (16 KB) + W half (16 KB)
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + OUT_STAGE_OFF + OUT_STAGE_BYTES;
constexpr int NUM_THREADS_LN = NUM_THREADS + LN_WARPS_PER_CTA * 32;   // + the LayerNorm warpgroup (LN = true)
constexpr int UTIL_REGS = 40, LN_REGS = 96 + (96 - UTIL_REGS);        // setmaxnreg: 4 x 40 + 12 x 96 + 4 x 152 warps = 640 x 96

template <int MODE, bool LN = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LN ? NUM_THREADS_LN : NUM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmOut, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the __shared__ array keeps the shared address space: LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES2 * STAGE2_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES2]  (leader's copy is the live one)
  uint64_t* empty_bar = bars + STAGES2;            // [STAGES2]
  uint64_t* tfull_bar = bars + 2 * STAGES2;        // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES2 + 2;   // [2]        (leader's copy is the live one)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  constexpr int BM2 = 2 * BM;
  const int num_m_tiles = (p.M + BM2 - 1) / BM2;
  const int num_n_tiles = p.N / BN;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int kblocks = p.K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  cluster_sync_all();  // peer barriers are initialised before anyone signals them
  if (warp == 2) tmem_alloc_2cta(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) {
  // (LN: 640 threads x 96 registers at launch; the four utility warps need 40, the LayerNorm warpgroup takes what they
  // free.  Each setmaxnreg sits at the top of its warpgroup's own branch, which is what lets ptxas allocate the code it
  // dominates within the new limit.)
  if (LN) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(UTIL_REGS));
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = tile / num_n_tiles, n_blk = tile % num_n_tiles;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE2_BYTES);  // both CTAs' bytes
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          uint8_t* sA = smem + stage * STAGE2_BYTES;
          uint8_t* sB = sA + BM * BK * 2;
          const int mrow = p.debug == 2 ? int(rank) * BM : m_blk * BM2 + int(rank) * BM;
          const int nrow = p.debug == 2 ? int(rank) * (BN / 2) : n_blk * BN + int(rank) * (BN / 2);
          tma_load_2d_2cta(sA, &tmA, leader_full, p.debug == 2 ? 0 : kb * BK, mrow);
          tma_load_2d_2cta(sB, &tmB, leader_full, p.debug == 2 ? 0 : kb * BK, nrow);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM2, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * STAGE2_BYTES);
          const uint64_t a_desc = umma_smem_desc_sw128(sA, 16, 1024);
          const uint64_t b_desc = umma_smem_desc_sw128(sA + BM * BK * 2, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_ss_2cta(d_tmem, a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc, (kb | k) != 0);
          umma_commit_2cta(&empty_bar[stage], 0b11);  // frees this smem slot in BOTH CTAs
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta(&tfull_bar[acc], 0b11);  // accumulator (both halves) complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }
  } else if (LN && warp >= 4 + NUM_EPI_WARPS) {
    // ------------------------------------------------------------------ LayerNorm of finished rows (see above)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(LN_REGS));   // the registers warps 0..3 gave up
    uint8_t* ln_smem = smem + STAGES2 * STAGE2_BYTES + OUT_STAGE_OFF + OUT_STAGE_BYTES;
    if (p.N == 1024) layernorm_warps<4>(p, ln_smem, warp - 4 - NUM_EPI_WARPS, lane);
    else             layernorm_warps<3>(p, ln_smem, warp - 4 - NUM_EPI_WARPS, lane);
  } else {
    // ------------------------------------------------------------------ epilogue (each CTA drains its 128 rows)
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    uint4* epi_stage = reinterpret_cast<uint4*>(smem + STAGES2 * STAGE2_BYTES + 256) + (warp - 4) * 24;
    uint8_t* out_stage = smem + STAGES2 * STAGE2_BYTES + OUT_STAGE_OFF + (warp - 4) * 2048;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m_blk = tile / num_n_tiles, n_blk = tile % num_n_tiles;
      const int row = m_blk * BM2 + int(rank) * BM + q * 32 + lane;
      EpiPrefetch pf;
      epilogue_prefetch<MODE>(p, epi_stage, grp, lane, row, n_blk, pf);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<MODE>(p, &tmOut, tmem_base, acc, q, grp, row, n_blk, epi_stage, pf, out_stage);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (LN && !(p.debug & 8)) {   // (8: perf experiments, no completion signalling -- with 4)
        // this warp's part of the tile is in memory (bulk stores complete) -> one arrival on the 128-row block's counter
        tma_store_wait_all();
        asm volatile("fence.proxy.async;" ::: "memory");   // the async-proxy writes before the generic-proxy release below
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          red_release_gpu_add(p.ln_counters + (m_blk * 2 + int(rank)), 1u);
        }
      }
    }
  }

  if (warp >= 4 && warp < 4 + NUM_EPI_WARPS) tma_store_wait_all();
  tc_fence_before();
  cluster_sync_all();  // the peer may still be signalling barriers / reading operands that live in this CTA
  if (warp == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

static int prof_kind(int mode, int K) {
  return mode == EPI_PATCH_EMBED ? PROF_GEMM_PATCH
         : mode == EPI_BIAS_GELU ? PROF_GEMM_FC1
         : mode == EPI_BIAS      ? PROF_GEMM_QKV
         : (K > 1024 ? PROF_GEMM_FC2 : PROF_GEMM_PROJ);
}

template <int MODE>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const Params& p,
           cudaStream_t stream) {
  FP_ENSURE_DYN_SMEM(gemm_kernel<MODE>, SMEM_BYTES);
  const int tiles = ((p.M + BM - 1) / BM) * (p.N / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  ProfScope prof(prof_kind(MODE, p.K), 2.0 * double(p.M) * double(p.N) * double(p.K), 1, stream);
  gemm_kernel<MODE><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, tmOut, p);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int launch2_ln(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const Params& p,
               cudaStream_t stream) {
  auto kern = gemm2_kernel<EPI_BIAS_LS_RES, true>;
  FP_ENSURE_DYN_SMEM(kern, SMEM2_BYTES + LN_SMEM_BYTES);
  const int tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * (p.N / BN);
  const int clusters = tiles < sm_count() / 2 ? tiles : sm_count() / 2;
  ProfScope prof(prof_kind(EPI_BIAS_LS_RES, p.K), 2.0 * double(p.M) * double(p.N) * double(p.K), 1, stream);
  kern<<<2 * clusters, NUM_THREADS_LN, SMEM2_BYTES + LN_SMEM_BYTES, stream>>>(tmA, tmB, tmOut, p);
  FP_CUDA(cudaGetLastError());
  return 0;
}

template <int MODE>
int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const Params& p,
            cudaStream_t stream) {
  FP_ENSURE_DYN_SMEM(gemm2_kernel<MODE>, SMEM2_BYTES);
  const int tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * (p.N / BN);
  const int clusters = tiles < sm_count() / 2 ? tiles : sm_count() / 2;
  ProfScope prof(prof_kind(MODE, p.K), 2.0 * double(p.M) * double(p.N) * double(p.K), 1, stream);
  gemm2_kernel<MODE><<<2 * clusters, NUM_THREADS, SMEM2_BYTES, stream>>>(tmA, tmB, tmOut, p);
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

static bool two_cta_path(int M) {
  static const int force = [] { const char* e = getenv("FP_GEMM_CTAS"); return e ? atoi(e) : 0; }();
  return force ? force == 2 : M >= 2048;
}

bool gemm_fuses_layernorm(int M, int N) { return two_cta_path(M) && (N == 1024 || N == 768); }

int gemm_bf16(const GemmArgs& a, cudaStream_t stream) {
  FP_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  FP_REQUIRE(a.N % BN == 0, "gemm: N=%d must be a multiple of %d", a.N, BN);
  FP_REQUIRE(a.K % BK == 0, "gemm: K=%d must be a multiple of %d", a.K, BK);
  FP_REQUIRE(a.A && a.W && a.out && a.bias, "gemm: null operand");
  FP_REQUIRE(a.ldo % 8 == 0, "gemm: output row stride must be a multiple of 8 elements");
  // large problems: 2-CTA clusters (256 x 256 tiles); small ones (e.g. the single query crop): 1-CTA 128 x 256 tiles
  const bool two = two_cta_path(a.M);
  const bool fuse_ln = a.ln_out != nullptr;
  if (fuse_ln) {
    FP_REQUIRE(a.mode == EPI_BIAS_LS_RES && gemm_fuses_layernorm(a.M, a.N) && a.ldo == a.N,
               "gemm: fused LayerNorm needs the residual epilogue, M >= 2048, N = ldo = 1024 or 768");
    FP_REQUIRE(a.ln_w && a.ln_b && a.ln_counters && a.ln_epoch > 0, "gemm: fused LayerNorm needs weights, counters and an epoch");
    FP_REQUIRE(a.ln_out != a.out && a.ln_out != a.A, "gemm: the fused LayerNorm output must not alias the GEMM's operands");
  }
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap_2d_bf16(&tmA, a.A, uint64_t(a.M), uint64_t(a.K), uint64_t(a.lda), BM, BK)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmB, a.W, uint64_t(a.N), uint64_t(a.K), uint64_t(a.K), two ? BN / 2 : BN, BK)) return rc;
  CUtensorMap tmOut;
  {
    // store map over the output rows this GEMM may write (patch-embed scatters with direct stores instead)
    const uint64_t out_rows = a.mode == EPI_PATCH_EMBED ? 32 : uint64_t(a.M);
    if (int rc = make_tmap_2d_bf16_sw64(&tmOut, a.out, out_rows, uint64_t(a.N), uint64_t(a.ldo), 32)) return rc;
  }
  Params p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.out = a.out; p.ldo = a.ldo;
  p.bias = a.bias; p.gamma = a.gamma; p.res = a.res;
  p.patches_per_img = a.patches_per_img > 0 ? a.patches_per_img : 1;
  p.tokens_per_img = a.tokens_per_img;
  p.token_offset = a.token_offset;
  static const int dbg = [] { const char* e = getenv("FP_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
  p.debug = dbg;
  p.ln_w = a.ln_w; p.ln_b = a.ln_b; p.ln_out = a.ln_out; p.ln_eps = a.ln_eps; p.ln_counters = a.ln_counters;
  p.ln_expect = a.ln_epoch * unsigned(NUM_EPI_WARPS * (a.N / BN));
  if (fuse_ln) return launch2_ln(tmA, tmB, tmOut, p, stream);
  switch (a.mode) {
    case EPI_BIAS: return two ? launch2<EPI_BIAS>(tmA, tmB, tmOut, p, stream) : launch<EPI_BIAS>(tmA, tmB, tmOut, p, stream);
    case EPI_BIAS_GELU:
      return two ? launch2<EPI_BIAS_GELU>(tmA, tmB, tmOut, p, stream) : launch<EPI_BIAS_GELU>(tmA, tmB, tmOut, p, stream);
    case EPI_BIAS_LS_RES:
      FP_REQUIRE(a.gamma && a.res, "gemm: LayerScale/residual epilogue needs gamma and res");
      return two ? launch2<EPI_BIAS_LS_RES>(tmA, tmB, tmOut, p, stream) : launch<EPI_BIAS_LS_RES>(tmA, tmB, tmOut, p, stream);
    case EPI_PATCH_EMBED:
      FP_REQUIRE(a.res && a.tokens_per_img > 0, "gemm: patch-embed epilogue needs pos-embed and token layout");
      return two ? launch2<EPI_PATCH_EMBED>(tmA, tmB, tmOut, p, stream) : launch<EPI_PATCH_EMBED>(tmA, tmB, tmOut, p, stream);
  }
  set_error("gemm: unknown epilogue mode %d", a.mode);
  return -1;
}

}  // namespace fp
