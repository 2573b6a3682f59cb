// ViT self-attention for the headline shape (224^2 crops: T = 261 tokens = 2 x 128 query rows + 5) on tcgen05, with the
// key axis of every query tile split into TWO INDEPENDENT STREAMS:
//
//     stream 0 = keys [0, 144)      stream 1 = keys [144, 272)   (261 real keys, padded to 272)
//
// Each stream has its own softmax warpgroup (4 warps, one per TMEM lane quarter, a thread = a query row), its own row
// maximum m_h, row sum l_h and its own output accumulator O_h = P_h V_h in TMEM, P_h = bf16(exp2((s - m_h) c)).  Nothing
// is exchanged between the streams while a tile is in flight; the tile is finished by
//
//     O = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1),     a_h = exp2((m_h - max(m_0, m_1)) c)
//
// -- the split-key form of flash attention (same arithmetic as two key blocks of an online softmax, without the
// rescale in between).  Why: in the single-stream kernel (attention.cu) the two softmax warps of a scheduler work on
// the same rows, meet at a max exchange and therefore sit in the same phase (TMEM load / max / exponentials): the MUFU
// idles while both load, and the tensor pipe idles while both exponentiate (measured: 4900 cycles per tile against a
// MUFU floor of 2176 and a tensor floor of ~1650).  Here stream 0 starts as soon as ITS logits exist, the tensor core
// runs stream 0's P.V and the next tile's stream-0 logits while stream 1 is still exponentiating, and the two warps of
// a scheduler are half a tile apart: one of them has exponentials to issue at (almost) all times.
//
//   warp 0        TMA loader   K, V of an (image, head) pair (double buffered), Q tiles through a 2-slot ring, tail rows
//   warp 1        MMA issuer of stream 0: P.V of tile g, then the logits of tile g+1 (their TMEM columns are free once
//                              the P.V steps reading the P stored over them have been issued: tensor work of one
//                              issuing thread executes in issue order)
//   warp 2        MMA issuer of stream 1, the same for its keys.  One issuer per stream: neither ever waits for the
//                              other stream's softmax, so stream 0 may run a tile ahead (measured with a single issuer
//                              in fixed order: both streams exponentiate at the same time and idle at the same time)
//   warp 3        tail rows    the 5 leftover query rows (256..260) by mma.sync from the resident K/V, in the TRANSPOSED
//                              form S^T = K Q^T, O^T = V^T P^T: the 8-wide N dimension holds the queries, so no
//                              fragment rows are wasted (a third of the HMMAs of the row-major form, which was the
//                              bottleneck of this kernel: 408 legacy HMMAs per pair on two warps took two tile times)
//   warps 4..7    stream 0     TMEM -> registers, row max, exponentials, bf16 P stored over the logits, (m_0, l_0) to smem
//   warps 8..11   stream 1     the same for its keys
//                              + the PREVIOUS tile's epilogue between its two passes: combine O_0 / O_1, normalise,
//                              TMA store (stream 0 never takes part, so it may run ahead)
//
// TMEM: logits / P 272 columns, O_0 double buffered (2 x 64: stream 0 runs up to a tile ahead of the epilogue),
// O_1 single (64).  Arithmetic contract: oracle/vit.py contract_attention(streams=...).
#include <stdlib.h>

#include "attention_common.cuh"
#include "kernels.h"

namespace fp {

namespace {

using namespace attn;

constexpr int HD = 64;
constexpr int QT = 128;
constexpr int T = 261;
constexpr int TPAD = 272;
constexpr int W0 = 144;            // keys of stream 0
constexpr int W1 = TPAD - W0;      // 128 keys of stream 1 (117 real)
constexpr int VALID_LAST = T - (W0 + 96);   // 21 real keys in stream 1's last 32-column group
constexpr int ROW_BYTES = HD * 2;
constexpr int Q_TILE_BYTES = QT * ROW_BYTES;
constexpr int KV_BYTES = TPAD * ROW_BYTES;
constexpr int N_TAIL = T % QT;     // 5 rows for the tail warps
constexpr int TAIL_MAX = 8;
constexpr int TAIL_BOX = 16;
constexpr int TILES_PER_PAIR = T / QT;   // 2
constexpr int NUM_MMA_WARPS = 2;
constexpr int NUM_THREADS = 384;
constexpr int TMEM_COLS = 512;
constexpr int S_COL = 0;
constexpr int O0_COL = 272;        // 2 x 64
constexpr int O1_COL = 400;        // 64
constexpr int NUM_CHUNKS = 5;      // P chunks: stream 0 keys [0,64) [64,128) [128,144); stream 1 [144,208) [208,272)
constexpr int ML_SLOTS = 4;        // (m_0, l_0) exchange slots, indexed by tile & 3

constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + 2 * Q_TILE_BYTES;
constexpr int OFF_V = OFF_K + 2 * KV_BYTES;
constexpr int OFF_XCH = OFF_V + 2 * KV_BYTES;              // float [ML_SLOTS][m_0 c | l_0][128]
constexpr int OFF_TQ = OFF_XCH + ML_SLOTS * 2 * 128 * 4;
constexpr int OFF_OST = OFF_TQ + 2 * TAIL_BOX * ROW_BYTES; // 4 stream-1 warps x 2 x (32 rows x 64 B), 64B swizzle
static_assert(OFF_OST % 1024 == 0, "TMA store staging must keep the swizzle alignment");
constexpr int OFF_BAR = OFF_OST + 8 * 2048;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;

struct Params {
  bf16* out;
  int B, H;
  float sl2;   // scale * log2(e)
  long long* dbg;   // perf experiments: per-phase cycle counters of CTA 0 (nullptr = off), tests/dev_attn_phases.py
  int skip_tail;    // perf experiments only (FP_ATTN_NOTAIL=1): the tail warps do no work (WRONG results for 5 rows)
};

// exponentials of 16 logits -> 8 packed bf16x2 words; returns the fp32 sum of the unrounded values
__device__ __forceinline__ float exp_group16(const uint32_t (&v)[16], float sl2, float msl, uint32_t (&packed)[8]) {
  const unsigned long long sl2_2 = pack_f2(sl2, sl2), nmsl_2 = pack_f2(-msl, -msl);
  unsigned long long acc[2] = {0ull, 0ull};
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    float y0, y1;
    unpack_f2(fma_f2(pack_f2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sl2_2, nmsl_2), y0, y1);
    const float e0 = ex2(y0), e1 = ex2(y1);
    acc[(j >> 1) & 1] = add_f2(acc[(j >> 1) & 1], pack_f2(e0, e1));
    packed[j >> 1] = pack_bf16x2(e0, e1);
  }
  float s0, s1;
  unpack_f2(add_f2(acc[0], acc[1]), s0, s1);
  return s0 + s1;
}

template <unsigned POLY>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_split_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQt,
                       const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmOut,
                       const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_full = bars;          // [2]
  uint64_t* kv_empty = bars + 2;     // [2]  both MMA issuers + the tail warp
  uint64_t* q_full = bars + 4;       // [2]
  uint64_t* q_empty = bars + 6;      // [2]
  uint64_t* tq_full = bars + 8;      // [2]
  uint64_t* tq_empty = bars + 10;    // [2]
  uint64_t* s_full = bars + 12;      // [2]  logits of stream h are in TMEM
  uint64_t* p_full = bars + 14;      // [5]  bf16 P of a chunk stored (4 arrivals: the warps of the owning stream)
  uint64_t* o0_full = bars + 19;     // [2]  P_0 V of a tile complete in O_0[buf]
  uint64_t* o1_full = bars + 21;     //      P_1 V complete in O_1
  uint64_t* o0_empty = bars + 22;    // [2]  the epilogue (stream-1 warps) has read O_0[buf]
  uint64_t* o1_empty = bars + 24;    //      epilogue has read O_1
  uint64_t* ml_full = bars + 25;     // [4]  stream 0 published (m_0 c, l_0) of tile g in slot g & 3
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 25 + ML_SLOTS);
  float* xch = reinterpret_cast<float*>(smem + OFF_XCH);   // [slot][0: m_0 * c, 1: l_0][row]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = p.B * p.H;
  constexpr int half_rows = TPAD / 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmQt);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], NUM_MMA_WARPS + 1);
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], NUM_MMA_WARPS);
      mbar_init(&tq_full[i], 1); mbar_init(&tq_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&o0_empty[i], 4);
      mbar_init(&o0_full[i], 1);
    }
    for (int i = 0; i < NUM_CHUNKS; ++i) mbar_init(&p_full[i], 4);
    mbar_init(o1_full, 1);
    mbar_init(o1_empty, 4);
    for (int i = 0; i < ML_SLOTS; ++i) mbar_init(&ml_full[i], 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // ---- a tile's epilogue, run by the stream-1 warp of lane quarter q for its 32 rows x 64 columns:
  //   O = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1),  a_h = exp2(m_h c - max(m_0 c, m_1 c)).
  // It runs one tile late, between the two passes of the warp's next tile: by then both P.V of the tile have long
  // finished (waiting for them right after the last P chunk cost 850 cycles per tile), and stream 0 -- which never
  // takes part -- is free to run ahead.
  auto epilogue = [&](uint32_t ge, int b, int h, int t, int q, float msl1, float l1, uint8_t* out_stage) {
    const int r = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t ms = ge & (ML_SLOTS - 1), ob = ge & 1;
    mbar_wait(&ml_full[ms], (ge >> 2) & 1);
    const float msl0 = xch[ms * 256 + r], l0 = xch[ms * 256 + 128 + r];
    const float mm = fmaxf(msl0, msl1);
    const float a0 = ex2(msl0 - mm), a1 = ex2(msl1 - mm);
    const float inv = 1.0f / (l0 * a0 + l1 * a1);
    const float w0 = a0 * inv, w1 = a1 * inv;
    tma_store_wait_read();     // this warp's previous bulk stores have finished reading the staging tiles
    __syncwarp();
    mbar_wait(&o0_full[ob], (ge >> 1) & 1);
    mbar_wait(o1_full, ge & 1);
    tc_fence_after();
    // in 16-column pieces: this runs with 96 logits of the next tile live in registers
#pragma unroll
    for (int hx = 0; hx < 4; ++hx) {
      uint32_t o0[16], o1[16];
      tmem_ld_32x32b_x16(tmem_base + lane_addr + O0_COL + ob * HD + hx * 16, o0);
      tmem_ld_32x32b_x16(tmem_base + lane_addr + O1_COL + hx * 16, o1);
      tmem_ld_wait();
      if (hx == 3) {   // both accumulators are in registers: the next tiles' P.V may overwrite them
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&o0_empty[ob]); mbar_arrive(o1_empty); }
      }
#pragma unroll
      for (int jv = 0; jv < 2; ++jv) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = jv * 8 + e * 2;
          const float x0 = fmaf(__uint_as_float(o0[c]), w0, __uint_as_float(o1[c]) * w1);
          const float x1 = fmaf(__uint_as_float(o0[c + 1]), w0, __uint_as_float(o1[c + 1]) * w1);
          w[e] = pack_bf16x2(x0, x1);
        }
        // two 32-column staging tiles; 64-byte swizzle: 16-byte chunk i of row `lane` sits at chunk i ^ ((lane>>1)&3)
        const int ch = (hx & 1) * 2 + jv;
        *reinterpret_cast<uint4*>(out_stage + (hx >> 1) * 2048 + lane * 64 + ((ch ^ ((lane >> 1) & 3)) << 4)) =
            make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) {
      const int row = b * T + t * QT + q * 32;
      tma_store_2d(&tmOut, out_stage, h * HD, row);                // 32 rows x 32 columns each
      tma_store_2d(&tmOut, out_stage + 2048, h * HD + 32, row);
      tma_store_commit();
    }
  };

  if (warp == 0) {
    // ---------------------------------------------------------------------------- TMA loader
    if (elect_one()) {
      int it = 0;
      uint32_t qi = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
        const int b = pair / p.H, h = pair - b * p.H;
        const int row0 = b * T;
        const int buf = it & 1;
        mbar_wait(&kv_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[buf], 2 * TPAD * ROW_BYTES);
        uint8_t* sK = smem + OFF_K + buf * KV_BYTES;
        uint8_t* sV = smem + OFF_V + buf * KV_BYTES;
        const int kcol = p.H * HD + h * HD, vcol = 2 * p.H * HD + h * HD;
        tma_load_2d(sK, &tmKV, &kv_full[buf], kcol, row0);
        tma_load_2d(sK + half_rows * ROW_BYTES, &tmKV, &kv_full[buf], kcol, row0 + half_rows);
        tma_load_2d(sV, &tmKV, &kv_full[buf], vcol, row0);
        tma_load_2d(sV + half_rows * ROW_BYTES, &tmKV, &kv_full[buf], vcol, row0 + half_rows);
        mbar_wait(&tq_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&tq_full[buf], TAIL_BOX * ROW_BYTES);
        tma_load_2d(smem + OFF_TQ + buf * TAIL_BOX * ROW_BYTES, &tmQt, &tq_full[buf], h * HD, row0 + T - N_TAIL);
        for (int t = 0; t < TILES_PER_PAIR; ++t, ++qi) {
          const int slot = qi & 1;
          mbar_wait(&q_empty[slot], ((qi >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[slot], Q_TILE_BYTES);
          tma_load_2d(smem + OFF_Q + slot * Q_TILE_BYTES, &tmQ, &q_full[slot], h * HD, row0 + t * QT);
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ---------------------------------------------------------------------------- MMA issuers (one per stream)
    if (elect_one()) {
      const int st = warp - 1;                                       // stream
      constexpr uint32_t idesc_pv = umma_idesc_bf16(QT, HD, 0, 1);   // B (= V) is MN-major
      const uint32_t idesc_s = st ? umma_idesc_bf16(QT, W1, 0, 0) : umma_idesc_bf16(QT, W0, 0, 0);
      const int my_pairs = (npairs - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
      const uint32_t ntiles = uint32_t(my_pairs > 0 ? my_pairs : 0) * TILES_PER_PAIR;
      const uint64_t q_desc0 = umma_smem_desc_sw128(smem_u32(smem + OFF_Q), 16, 1024);
      const uint64_t k_desc0 = umma_smem_desc_sw128(smem_u32(smem + OFF_K), 16, 1024) + uint64_t(st ? ((W0 * ROW_BYTES) >> 4) : 0);
      const uint64_t v_desc0 = umma_smem_desc_sw128(smem_u32(smem + OFF_V), 1024, 1024);
      const uint32_t s_tmem = tmem_base + S_COL + (st ? W0 : 0);
      const bool timing = p.dbg != nullptr && blockIdx.x == 0;
      long long t_o = 0, t_pv = 0, t_s = 0;
      // this stream's logits of tile g (pair iteration `it`, tile `t` of the pair)
      auto issue_s = [&](uint32_t g, int it, int t) {
        const int buf = it & 1, slot = g & 1;
        if (t == 0) mbar_wait(&kv_full[buf], (it >> 1) & 1);
        mbar_wait(&q_full[slot], (g >> 1) & 1);
        tc_fence_after();
        const uint64_t q_desc = q_desc0 + uint64_t(slot * (Q_TILE_BYTES >> 4));
        const uint64_t k_desc = k_desc0 + uint64_t(buf * (KV_BYTES >> 4));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(s_tmem, q_desc + uint64_t(2 * k), k_desc + uint64_t(2 * k), idesc_s, k != 0);
        umma_commit(&s_full[st]);
        umma_commit(&q_empty[slot]);     // (two arrivals free the slot: one per stream)
      };
      if (ntiles > 0) issue_s(0, 0, 0);
      int it = 0, t = 0;
      for (uint32_t g = 0; g < ntiles; ++g) {
        int nit = it, nt = t + 1;
        if (nt == TILES_PER_PAIR) { nt = 0; ++nit; }
        const int buf = it & 1;
        const uint32_t ob = g & 1;
        const uint64_t v_desc = v_desc0 + uint64_t(buf * (KV_BYTES >> 4));
        long long c0 = 0, c1 = 0, c2 = 0;
        if (timing) c0 = clock64();
        if (st == 0) {
          // ---- O_0[ob] = P_0 V[0:144)
          mbar_wait(&o0_empty[ob], ((g >> 1) & 1) ^ 1);
          tc_fence_after();
          if (timing) c1 = clock64();
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            mbar_wait(&p_full[c], g & 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int key0 = c * 64 + k * 16;
              if (key0 < W0) {
                // P of keys [key0, key0+16): 8 packed columns at the start of the 32-column logit group they came from
                const uint32_t pcol = uint32_t((key0 & ~31) + ((key0 & 16) >> 1));
                umma_bf16_ts(tmem_base + O0_COL + ob * HD, tmem_base + S_COL + pcol,
                             v_desc + uint64_t(key0 * (ROW_BYTES >> 4)), idesc_pv, key0 != 0);
              }
            }
          }
          umma_commit(&o0_full[ob]);
        } else {
          // ---- O_1 = P_1 V[144:272)
          mbar_wait(o1_empty, (g & 1) ^ 1);
          tc_fence_after();
          if (timing) c1 = clock64();
#pragma unroll
          for (int c = 3; c < 5; ++c) {
            mbar_wait(&p_full[c], g & 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int rel = (c - 3) * 64 + k * 16;
              const int key0 = W0 + rel;
              const uint32_t pcol = uint32_t(W0 + (rel & ~31) + ((rel & 16) >> 1));
              umma_bf16_ts(tmem_base + O1_COL, tmem_base + S_COL + pcol, v_desc + uint64_t(key0 * (ROW_BYTES >> 4)),
                           idesc_pv, rel != 0);
            }
          }
          umma_commit(o1_full);
        }
        if (t == TILES_PER_PAIR - 1) umma_commit(&kv_empty[buf]);
        if (timing) c2 = clock64();
        if (g + 1 < ntiles) issue_s(g + 1, nit, nt);
        if (timing) { t_o += c1 - c0; t_pv += c2 - c1; t_s += clock64() - c2; }
        it = nit; t = nt;
      }
      if (timing) { p.dbg[10 + 3 * st] = t_o; p.dbg[11 + 3 * st] = t_pv; p.dbg[12 + 3 * st] = t_s; }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---------------------------------------------------------------------------- stream 0: keys [0, 144)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t sbase = tmem_base + (uint32_t(q * 32) << 16) + S_COL;
    auto publish = [&](int c) {   // P of a chunk is published one group late: its tcgen05.st completes under the next exponentials
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[c]);
    };
    uint32_t g = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      for (int t = 0; t < TILES_PER_PAIR; ++t, ++g) {
        uint32_t s0[32], s1[32], s2[32], s3[32], s4[16];
        const bool timing = p.dbg != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0;
        long long tk0 = 0, tk1 = 0, tk2 = 0;
        if (timing) tk0 = clock64();
        mbar_wait(&s_full[0], g & 1);
        tc_fence_after();
        if (timing) tk1 = clock64();
        // group 3 and the 16-column tail are reduced first and die (they are read a second time during the exponentials):
        // the three groups that stay in registers between the passes are loaded behind them (144 live logits do not fit
        // the 168-register budget)
        tmem_ld_32x32b_x32(sbase + 96, s3);
        tmem_ld_32x32b_x16(sbase + 128, s4);
        tmem_ld_wait();
        tmem_ld_32x32b_x32(sbase, s0);
        tmem_ld_32x32b_x32(sbase + 32, s1);
        tmem_ld_32x32b_x32(sbase + 64, s2);
        float m = max_group<false>(s3, 32, -INFINITY);
#pragma unroll
        for (int j = 0; j < 16; ++j) m = fmaxf(m, __uint_as_float(s4[j]));
        tmem_ld_wait();
        m = max_group<false>(s0, 32, m);
        m = max_group<false>(s1, 32, m);
        m = max_group<false>(s2, 32, m);
        const float msl = m * p.sl2;
        if (timing) tk2 = clock64();
        uint32_t pk[16];
        float l = exp_group<false, POLY>(s0, p.sl2, msl, 32, pk);
        tmem_st_32x32b_x16(sbase, pk);
        l += exp_group<false, POLY>(s1, p.sl2, msl, 32, pk);
        tmem_st_32x32b_x16(sbase + 32, pk);
        tmem_ld_32x32b_x32(sbase + 96, s3);              // second read of group 3 and the tail, under group 2's exponentials
        tmem_ld_32x32b_x16(sbase + 128, s4);
        l += exp_group<false, POLY>(s2, p.sl2, msl, 32, pk);
        publish(0);                                      // keys [0, 64)
        tmem_st_32x32b_x16(sbase + 64, pk);
        tmem_ld_wait();
        l += exp_group<false, POLY>(s3, p.sl2, msl, 32, pk);
        tmem_st_32x32b_x16(sbase + 96, pk);
        uint32_t pk8[8];
        l += exp_group16(s4, p.sl2, msl, pk8);
        publish(1);                                      // keys [64, 128)
        tmem_st_32x32b_x8(sbase + 128, pk8);
        publish(2);                                      // keys [128, 144)
        // (m_0 c, l_0) of this tile for the epilogue
        float* slot = xch + (g & (ML_SLOTS - 1)) * 256;
        slot[r] = msl;
        slot[128 + r] = l;
        __syncwarp();
        if (lane == 0) mbar_arrive(&ml_full[g & (ML_SLOTS - 1)]);
        if (timing) {
          p.dbg[0] += tk1 - tk0; p.dbg[1] += tk2 - tk1; p.dbg[2] += clock64() - tk2; p.dbg[3] += 1;
        }
      }
    }
  } else if (warp >= 8) {
    // ---------------------------------------------------------------------------- stream 1: keys [144, 272)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t sbase = tmem_base + (uint32_t(q * 32) << 16) + S_COL + W0;
    uint8_t* out_stage = smem + OFF_OST + (warp - 8) * 4096;   // two 32-row x 64-byte tiles
    auto publish = [&](int c) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[c]);
    };
    uint32_t g = 0;
    int pb = 0, ph = 0, pt = 0;          // coordinates of tile g - 1 (its epilogue is still owed) ...
    float pmsl = 0.f, pl = 0.f;          // ... and this stream's (m_1 c, l_1) of it
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int b = pair / p.H, h = pair - b * p.H;
      for (int t = 0; t < TILES_PER_PAIR; ++t, ++g) {
        const uint32_t par = g & 1;
        uint32_t s0[32], s1[32], s2[32], s3[32];
        const bool timing = p.dbg != nullptr && blockIdx.x == 0 && warp == 8 && lane == 0;
        long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
        if (timing) tk0 = clock64();
        mbar_wait(&s_full[1], par);
        tc_fence_after();
        if (timing) tk1 = clock64();
        tmem_ld_32x32b_x32(sbase + 96, s3);
        tmem_ld_32x32b_x32(sbase, s0);
        tmem_ld_32x32b_x32(sbase + 32, s1);
        tmem_ld_32x32b_x32(sbase + 64, s2);
        tmem_ld_wait();
        float m = max_group<true>(s3, VALID_LAST, -INFINITY);   // keys 240..260 are real, 261..271 padding
        m = max_group<false>(s0, 32, m);
        m = max_group<false>(s1, 32, m);
        m = max_group<false>(s2, 32, m);
        const float msl = m * p.sl2;
        if (timing) tk2 = clock64();
        // the previous tile's epilogue (this warp's half): its P.V finished while this tile's logits were produced
        if (g > 0) epilogue(g - 1, pb, ph, pt, q, pmsl, pl, out_stage);
        pb = b; ph = h; pt = t;
        if (timing) tk3 = clock64();
        uint32_t pk[16];
        float l = exp_group<false, POLY>(s0, p.sl2, msl, 32, pk);
        tmem_st_32x32b_x16(sbase, pk);
        tmem_ld_32x32b_x32(sbase + 96, s3);
        l += exp_group<false, POLY>(s1, p.sl2, msl, 32, pk);
        tmem_st_32x32b_x16(sbase + 32, pk);
        l += exp_group<false, POLY>(s2, p.sl2, msl, 32, pk);
        publish(3);                                      // keys [144, 208)
        tmem_st_32x32b_x16(sbase + 64, pk);
        tmem_ld_wait();
        l += exp_group<true, POLY>(s3, p.sl2, msl, VALID_LAST, pk);   // P of the padding keys = 0
        tmem_st_32x32b_x16(sbase + 96, pk);
        publish(4);                                      // keys [208, 272)
        pmsl = msl; pl = l;
        if (timing) {
          const long long tk4 = clock64();
          p.dbg[4] += tk1 - tk0; p.dbg[5] += tk2 - tk1; p.dbg[8] += tk3 - tk2; p.dbg[6] += tk4 - tk3; p.dbg[9] += 1;
        }
      }
    }
    if (g > 0) epilogue(g - 1, pb, ph, pt, q, pmsl, pl, out_stage);
    tma_store_wait_all();   // the staging tiles must outlive the bulk stores reading them
  } else if (warp == 3) {
    // ---------------------------------------------------------------------------- tail queries (5 rows), one warp
    // mma.sync.m16n8k16 in transposed form on the K/V tiles already in shared memory (ldmatrix understands the TMA 128B
    // swizzle: every 8x8 sub-matrix row is one 16-byte chunk):
    //   S^T[key, query] = K Q^T     A = K block (16 keys x 16 dims, row-major), B = Q^T (the 8 query columns = Q rows)
    //   O^T[dim, query] = V^T P^T   A = V^T (ldmatrix.trans of V), B = P^T: the bf16 pairs of the S^T accumulator fragment
    //                               (key = lane/4, queries 2(lane%4)+{0,1}) moved to the B layout by movmatrix.trans
    // All 17 key blocks of logits stay in registers (68 per thread): one pass.  These rows use ONE stream (global row
    // max).  Columns (queries) 5..7 of the fragments are ignored.
    const int gq = lane >> 2, tq = lane & 3;
    constexpr int NBLK = TPAD / 16;   // 17
    int it = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
      const int b = pair / p.H, h = pair - b * p.H;
      const int buf = it & 1;
      const uint32_t sK = smem_u32(smem + OFF_K + buf * KV_BYTES);
      const uint32_t sV = smem_u32(smem + OFF_V + buf * KV_BYTES);
      const uint32_t sTQ = smem_u32(smem + OFF_TQ + buf * TAIL_BOX * ROW_BYTES);
      mbar_wait(&kv_full[buf], (it >> 1) & 1);
      mbar_wait(&tq_full[buf], (it >> 1) & 1);
      if (p.skip_tail) {
        __syncwarp();
        if (lane == 0) { mbar_arrive(&kv_empty[buf]); mbar_arrive(&tq_empty[buf]); }
        continue;
      }
      // B fragments of Q^T for the four 16-dim k-steps: matrices (queries 0..7, dims 8c..8c+7), c = 0..7
      uint32_t qb[8];
      {
        const int row = lane & 7;
#pragma unroll
        for (int hx = 0; hx < 2; ++hx) {
          const int chunk = 4 * hx + (lane >> 3);
          uint32_t r4[4];
          ldmatrix_x4(sTQ + row * ROW_BYTES + ((chunk ^ (row & 7)) << 4), r4);
#pragma unroll
          for (int i = 0; i < 4; ++i) qb[4 * hx + i] = r4[i];
        }
      }
      // ---- logits: st[blk] = (key lane/4, queries 2tq, 2tq+1), (key lane/4 + 8, same queries)
      float sacc[NBLK][4];
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) {
#pragma unroll
        for (int e = 0; e < 4; ++e) sacc[blk][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t ka[4];   // A = K[keys blk*16 .. +15][dims 16ks .. +15]
          const int key = blk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int chunk = 2 * ks + (lane >> 4);
          ldmatrix_x4(sK + key * ROW_BYTES + ((chunk ^ (key & 7)) << 4), ka);
          mma_bf16_16816(sacc[blk], ka, qb[2 * ks], qb[2 * ks + 1]);
        }
      }
      // ---- row (= query) max over the keys: over this thread's blocks, then over the lanes sharing lane % 4
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) {
        const int k_lo = blk * 16 + gq, k_hi = k_lo + 8;
        if (k_lo < T) { m0 = fmaxf(m0, sacc[blk][0]); m1 = fmaxf(m1, sacc[blk][1]); }
        if (k_hi < T) { m0 = fmaxf(m0, sacc[blk][2]); m1 = fmaxf(m1, sacc[blk][3]); }
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
      }
      const float msl0 = m0 * p.sl2, msl1 = m1 * p.sl2;
      // ---- P = exp2(...), row sums, O^T += V^T P^T
      float oacc[4][4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int e = 0; e < 4; ++e) oacc[mt][e] = 0.f;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) {
        const int k_lo = blk * 16 + gq, k_hi = k_lo + 8;
        const float e00 = k_lo < T ? ex2(fmaf(sacc[blk][0], p.sl2, -msl0)) : 0.f;
        const float e01 = k_lo < T ? ex2(fmaf(sacc[blk][1], p.sl2, -msl1)) : 0.f;
        const float e10 = k_hi < T ? ex2(fmaf(sacc[blk][2], p.sl2, -msl0)) : 0.f;
        const float e11 = k_hi < T ? ex2(fmaf(sacc[blk][3], p.sl2, -msl1)) : 0.f;
        l0 += e00 + e10;
        l1 += e01 + e11;
        // (key, query pair) fragments -> B fragments of P^T: b0 = keys 0..7, b1 = keys 8..15 of the block
        uint32_t pb0, pb1;
        asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(pb0) : "r"(pack_bf16x2(e00, e01)));
        asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(pb1) : "r"(pack_bf16x2(e10, e11)));
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          uint32_t va[4];   // A = V^T[dims 16mt .. +15][keys blk*16 .. +15] = transposed 8x8 blocks of V
          const int key = blk * 16 + (lane & 7) + (lane >> 4) * 8;
          const int chunk = 2 * mt + ((lane >> 3) & 1);
          ldmatrix_x4_trans(sV + key * ROW_BYTES + ((chunk ^ (key & 7)) << 4), va);
          mma_bf16_16816(oacc[mt], va, pb0, pb1);
        }
      }
      // K, V and the tail Q rows of this pair are no longer needed by this warp
      __syncwarp();
      if (lane == 0) { mbar_arrive(&kv_empty[buf]); mbar_arrive(&tq_empty[buf]); }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        l0 += __shfl_xor_sync(0xffffffffu, l0, o);
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
      }
      const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
      // oacc[mt] = (dim 16mt + gq, queries 2tq, 2tq+1), (dim 16mt + gq + 8, same queries)
      const int q0 = 2 * tq;
      bf16* orow = p.out + (size_t(b) * T + (T - N_TAIL)) * (p.H * HD) + h * HD;
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int d0 = 16 * mt + gq;
        if (q0 < N_TAIL) {
          orow[size_t(q0) * (p.H * HD) + d0] = __float2bfloat16_rn(oacc[mt][0] * inv0);
          orow[size_t(q0) * (p.H * HD) + d0 + 8] = __float2bfloat16_rn(oacc[mt][2] * inv0);
        }
        if (q0 + 1 < N_TAIL) {
          orow[size_t(q0 + 1) * (p.H * HD) + d0] = __float2bfloat16_rn(oacc[mt][1] * inv1);
          orow[size_t(q0 + 1) * (p.H * HD) + d0 + 8] = __float2bfloat16_rn(oacc[mt][3] * inv1);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

// T must be 261 (224^2 crops).  poly_mask: which of the 16 (even, odd) column pairs of every 32-logit group take the
// FMA-pipe polynomial exp2 instead of the MUFU (0 = none, 0x1111 = 25 %, 0x5555 = 50 %).
int attention_split_bf16(const bf16* qkv, bf16* out, int B, int T_, int H, float scale, unsigned poly_mask,
                         cudaStream_t stream) {
  FP_REQUIRE(T_ == T, "attention_split: built for %d tokens, got %d", T, T_);
  FP_REQUIRE(B > 0 && H > 0, "attention: empty problem");
  const int C = 3 * H * HD;
  CUtensorMap tmQ, tmQt, tmKV, tmOut;
  const uint64_t rows = uint64_t(B) * T;
  if (int rc = make_tmap_2d_bf16(&tmQ, qkv, rows, uint64_t(C), uint64_t(C), QT, HD)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmQt, qkv, rows, uint64_t(C), uint64_t(C), TAIL_BOX, HD)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmKV, qkv, rows, uint64_t(C), uint64_t(C), uint32_t(TPAD / 2), HD)) return rc;
  if (int rc = make_tmap_2d_bf16_sw64(&tmOut, out, rows, uint64_t(H) * HD, uint64_t(H) * HD, 32)) return rc;
  Params p;
  p.out = out; p.B = B; p.H = H;
  p.sl2 = scale * 1.4426950408889634f;
  p.skip_tail = getenv("FP_ATTN_NOTAIL") ? atoi(getenv("FP_ATTN_NOTAIL")) : 0;
  p.dbg = getenv("FP_ATTN_DBG") ? reinterpret_cast<long long*>(strtoull(getenv("FP_ATTN_DBG"), nullptr, 0)) : nullptr;
  const int npairs = B * H;
  const int grid = npairs < sm_count() ? npairs : sm_count();
  ProfScope prof(PROF_ATTENTION, 4.0 * double(B) * H * double(T) * T * HD, 1, stream);
#define FP_LAUNCH_SPLIT(MASK_)                                                          \
  do {                                                                                  \
    auto kern = attention_split_kernel<MASK_>;                                          \
    FP_ENSURE_DYN_SMEM(kern, SMEM_BYTES);                                               \
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmQt, tmKV, tmOut, p);         \
  } while (0)
  switch (poly_mask) {
    case 0x1111u: FP_LAUNCH_SPLIT(0x1111u); break;
    case 0x5555u: FP_LAUNCH_SPLIT(0x5555u); break;
    default: FP_LAUNCH_SPLIT(0u); break;
  }
#undef FP_LAUNCH_SPLIT
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fp
