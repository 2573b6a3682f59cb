// ViT self-attention for crops above 224^2 (reference-native 420^2 -> T = 905 tokens, the refiner's 518^2 -> 1374;
// SURVEY.md section 0.1): TWO query tiles in flight per CTA, each with its own softmax warpgroup, its own MMA issuer and
// DOUBLE-BUFFERED logits in TMEM, walking over key blocks of 96 with an online softmax.
//
// Why (measured on the previous kernel, attention_long.cu: one query tile, 256-key blocks, two warps per row meeting at a
// max exchange, a single logit buffer: 268 TFLOP/s in-step = 0.19 of the tensor roofline, 41 % of the 420^2 step): a
// single stream is a dependency chain  S -> max -> exp -> P.V -> S ...  in which the MUFU waits for the tensor core
// and the tensor core for the MUFU.  The softmax here is MUFU-bound (16 exponentials per clock and SM against 256
// FLOPs of tensor work per exponential at head dim 64 = half the tensor roofline), so the design goal is to keep the
// MUFU fed:
//   * a thread owns a whole query row of a 96-key block (three 32-column TMEM loads, all kept in registers): no exchange
//     of row maxima between warps, no second read of the logits;
//   * the two warps of a scheduler belong to DIFFERENT query tiles (streams) and drift freely: while one loads / takes
//     the max / rescales O, the other exponentiates;
//   * each stream has two logit buffers: S of block j+2 is issued as soon as P.V of block j (which read the P stored over
//     that buffer) has been issued, so a stream never waits for its logits inside an item;
//   * one MMA-issuing warp per stream (tensor work of one issuing thread executes in issue order; the two issuers never
//     wait on each other's softmax).
//
//   warp 0        TMA loader   Q tile per stream and item (2 slots each); K blocks through a 5-stage ring, V blocks through
//                              a 4-stage ring, shared by both streams (a stage is free when both issuers have released it)
//   warp 1 / 2    MMA issuers  stream 0 / 1:  P.V of block j (P read from TMEM, V as MN-major smem operand), then S of
//                              block j+2
//   warp 3        TMEM allocator
//   warps 4..7    softmax of stream 0 (query tile 2i),  warps 8..11  softmax of stream 1 (query tile 2i+1):
//                              row max, exponentials against the running max, bf16 P stored over the consumed logits and
//                              published per 32-key group, O rescaled in TMEM when a row of the warp raised its max
//                              (skipped otherwise: a factor of exactly 1), item epilogue O / l -> bf16 -> HBM deferred
//                              into the next item's first block
//
// TMEM (512 columns): S[stream][buffer] 4 x 96, O[stream] 2 x 64.
// Work item = (image, head, pair of query tiles).  Arithmetic contract: oracle/vit.py contract_attention with
// key_block = 96, lazy_tau = 8 -- block-wise flash attention: P of block j is exp2((s - m_j) c) rounded to bf16, m_j the
// row's reference maximum after block j (the running maximum, updated lazily: see LAZY_TAU).
#include <stdlib.h>

#include "attention_common.cuh"
#include "kernels.h"

namespace fp {

namespace {

using namespace attn;

constexpr int HD = 64;
constexpr int QT = 128;
constexpr int KB = 96;                        // keys per block
constexpr int ROW_BYTES = HD * 2;
constexpr int Q_TILE_BYTES = QT * ROW_BYTES;  // 16 KB
constexpr int KV_BLOCK_BYTES = KB * ROW_BYTES;  // 12 KB
constexpr int K_STAGES = 5, V_STAGES = 4;
constexpr int NUM_THREADS = 384;
constexpr int TMEM_COLS = 512;
constexpr int O_COL = 4 * KB;                 // 384
constexpr int NGROUPS = KB / 32;              // 3 groups of 32 keys = P chunks
constexpr float LAZY_TAU = 8.0f;              // log2 of the largest un-normalised P before the reference max moves

constexpr int OFF_Q = 0;                                   // [2 streams][2 slots]
constexpr int OFF_K = OFF_Q + 4 * Q_TILE_BYTES;
constexpr int OFF_V = OFF_K + K_STAGES * KV_BLOCK_BYTES;
constexpr int OFF_OST = (OFF_V + V_STAGES * KV_BLOCK_BYTES + 1023) / 1024 * 1024;   // 8 warps x 2 x (32 rows x 64 B)
constexpr int OFF_BAR = OFF_OST + 8 * 4096;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;

struct Params {
  bf16* out;
  int B, T, H;
  int tpad;   // keys padded to a multiple of 16
  int nq;     // query tiles per (image, head)
  int npq;    // query tile pairs per (image, head)
  int nkb;    // key blocks
  float sl2;  // scale * log2(e)
};

template <unsigned POLY>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                      const __grid_constant__ CUtensorMap tmOut, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* k_full = bars;            // [5]
  uint64_t* k_empty = bars + 5;       // [5]  both issuers
  uint64_t* v_full = bars + 10;       // [4]
  uint64_t* v_empty = bars + 14;      // [4]  both issuers
  uint64_t* q_full = bars + 18;       // [2 streams][2 slots]
  uint64_t* q_empty = bars + 22;      // [2][2]
  uint64_t* s_full = bars + 26;       // [2 streams][2 buffers]
  // Every barrier between a stream's softmax warps and its issuer exists TWICE, alternating with the block parity: a
  // parity wait is only meaningful within one phase of its barrier, and with the lazy rescale the softmax warps no
  // longer wait for P.V(j-1), so they (or, for rows past the end of the image, which have nothing to compute, certainly)
  // can be two blocks ahead of the issuer.  With alternating barriers the arrivals for block j+2 come behind S(j+2),
  // which the issuer only issues after it has consumed the barriers of block j.
  uint64_t* p_full = bars + 30;       // [2 streams][2 parities][3 groups]  (4 arrivals: the stream's warps)
  // [2 streams][2]: P.V of key block j complete, on barrier j & 1.  Two barriers per stream because the softmax warps wait
  // for it only when they have to rescale O: a parity wait is meaningful only within one phase of its barrier, and with
  // alternating barriers "block j-1" is always either the pending or the last completed phase of its barrier (S of block
  // j, which the warp has just waited for, was issued behind P.V of block j-2).
  uint64_t* o_full = bars + 42;
  uint64_t* o_ready = bars + 46;      // [2 streams][2 parities]  O rescaled / read out: this block's P.V may accumulate (4 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 50);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = p.B * p.H * p.npq;
  const int my_items = nitems > int(blockIdx.x) ? (nitems - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < K_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); }
    for (int i = 0; i < V_STAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2); }
    for (int i = 0; i < 4; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_full[i], 1); }
    for (int i = 0; i < 12; ++i) mbar_init(&p_full[i], 4);
    for (int i = 0; i < 4; ++i) { mbar_init(&o_full[i], 1); mbar_init(&o_ready[i], 4); }
    fence_barrier_init();
  }
  if (warp == 3) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  auto block_keys = [&](int kb) { return (p.tpad - kb * KB) < KB ? (p.tpad - kb * KB) : KB; };   // multiple of 16

  if (warp == 0) {
    // ---------------------------------------------------------------------------- TMA loader
    if (elect_one()) {
      uint32_t j = 0, qi = 0;   // block / item counters of this CTA
      int ks = 0, vs = 0;
      uint32_t kphase = 0, vphase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++qi) {
        const int pair = item / p.npq, tp = item - pair * p.npq;
        const int b = pair / p.H, h = pair - b * p.H;
        const int row0 = b * p.T;
        const int slot = qi & 1;
#pragma unroll
        for (int st = 0; st < 2; ++st) {
          // (a tile past the last one loads whatever rows follow: they are computed and never stored)
          mbar_wait(&q_empty[st * 2 + slot], ((qi >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[st * 2 + slot], Q_TILE_BYTES);
          tma_load_2d(smem + OFF_Q + (st * 2 + slot) * Q_TILE_BYTES, &tmQ, &q_full[st * 2 + slot], h * HD,
                      row0 + (2 * tp + st) * QT);
        }
        const int kcol = p.H * HD + h * HD, vcol = 2 * p.H * HD + h * HD;
        for (int kb = 0; kb < p.nkb; ++kb, ++j) {
          const int krow = row0 + kb * KB;
          mbar_wait(&k_empty[ks], kphase ^ 1);
          mbar_arrive_expect_tx(&k_full[ks], KV_BLOCK_BYTES);
          tma_load_2d(smem + OFF_K + ks * KV_BLOCK_BYTES, &tmKV, &k_full[ks], kcol, krow);
          mbar_wait(&v_empty[vs], vphase ^ 1);
          mbar_arrive_expect_tx(&v_full[vs], KV_BLOCK_BYTES);
          tma_load_2d(smem + OFF_V + vs * KV_BLOCK_BYTES, &tmKV, &v_full[vs], vcol, krow);
          if (++ks == K_STAGES) { ks = 0; kphase ^= 1; }
          if (++vs == V_STAGES) { vs = 0; vphase ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ---------------------------------------------------------------------------- MMA issuers (one per stream)
    if (elect_one()) {
      const int st = warp - 1;
      constexpr uint32_t idesc_pv = umma_idesc_bf16(QT, HD, 0, 1);
      const uint64_t q_desc0 = umma_smem_desc_sw128(smem_u32(smem + OFF_Q + st * 2 * Q_TILE_BYTES), 16, 1024);
      const uint64_t k_desc0 = umma_smem_desc_sw128(smem_u32(smem + OFF_K), 16, 1024);
      const uint64_t v_desc0 = umma_smem_desc_sw128(smem_u32(smem + OFF_V), 1024, 1024);
      const uint32_t s_tmem0 = tmem_base + uint32_t(st * 2 * KB);
      const uint32_t o_tmem = tmem_base + O_COL + st * HD;
      const uint32_t nblocks = uint32_t(my_items) * uint32_t(p.nkb);
      // state of the block whose S is issued next ("n")
      uint32_t n_j = 0, n_qi = 0;
      int n_kb = 0, n_ks = 0;
      uint32_t n_kphase = 0;
      auto issue_s = [&]() {
        const int nkeys = block_keys(n_kb);
        const int slot = n_qi & 1;
        if (n_kb == 0) mbar_wait(&q_full[st * 2 + slot], (n_qi >> 1) & 1);
        mbar_wait(&k_full[n_ks], n_kphase);
        tc_fence_after();
        const uint64_t q_desc = q_desc0 + uint64_t(slot * (Q_TILE_BYTES >> 4));
        const uint64_t k_desc = k_desc0 + uint64_t(n_ks * (KV_BLOCK_BYTES >> 4));
        const uint32_t idesc = umma_idesc_bf16(QT, nkeys, 0, 0);
        const uint32_t d = s_tmem0 + (n_j & 1) * KB;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(d, q_desc + uint64_t(2 * k), k_desc + uint64_t(2 * k), idesc, k != 0);
        umma_commit(&s_full[st * 2 + (n_j & 1)]);
        umma_commit(&k_empty[n_ks]);
        if (n_kb == p.nkb - 1) umma_commit(&q_empty[st * 2 + slot]);
        if (++n_ks == K_STAGES) { n_ks = 0; n_kphase ^= 1; }
        if (++n_kb == p.nkb) { n_kb = 0; ++n_qi; }
        ++n_j;
      };
      if (nblocks > 0) issue_s();
      if (nblocks > 1) issue_s();
      int kb = 0, vs = 0;
      uint32_t vphase = 0;
      for (uint32_t j = 0; j < nblocks; ++j) {
        const int nkeys = block_keys(kb);
        const uint64_t v_desc = v_desc0 + uint64_t(vs * (KV_BLOCK_BYTES >> 4));
        const uint32_t p_tmem = s_tmem0 + (j & 1) * KB;
        mbar_wait(&v_full[vs], vphase);
        mbar_wait(&o_ready[st * 2 + (j & 1)], (j >> 1) & 1);   // O rescaled for this block's reference max (or read out by the previous item)
        tc_fence_after();
#pragma unroll
        for (int g = 0; g < NGROUPS; ++g) {
          mbar_wait(&p_full[(st * 2 + (j & 1)) * 3 + g], (j >> 1) & 1);   // (every group barrier completes one phase per use, keys or not)
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int key0 = g * 32 + k * 16;
            if (key0 < nkeys)   // P of keys [key0, key0+16): 8 packed columns at the start of their 32-column logit group
              umma_bf16_ts(o_tmem, p_tmem + uint32_t(g * 32 + k * 8), v_desc + uint64_t(key0 * (ROW_BYTES >> 4)), idesc_pv,
                           (kb | key0) != 0);
          }
        }
        umma_commit(&o_full[st * 2 + (j & 1)]);
        umma_commit(&v_empty[vs]);
        if (++vs == V_STAGES) { vs = 0; vphase ^= 1; }
        if (++kb == p.nkb) kb = 0;
        if (j + 2 < nblocks) issue_s();        // S of block j+2 over the P just consumed
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------------------- softmax + epilogue of one stream
    const int st = (warp - 4) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t s_tmem0 = tmem_base + lane_addr + uint32_t(st * 2 * KB);
    const uint32_t o_tmem = tmem_base + lane_addr + O_COL + st * HD;
    uint8_t* out_stage = smem + OFF_OST + (warp - 4) * 4096;   // two 32-row x 64-byte tiles, 64B swizzle
    uint32_t j = 0;        // block counter of this CTA
    // the previous item's normalisation + store is owed until the next item's first block (or the end)
    float prev_l = 1.f;
    int prev_b = 0, prev_h = 0, prev_t = 0;

    // O / l -> bf16 -> HBM for query tile t of (b, h): 32 rows x 64 columns per warp
    auto store_item = [&](int b, int h, int t, float l) {
      const bool warp_active = t * QT + q * 32 < p.T;
      if (!warp_active) return;
      const bool full_rows = t * QT + q * 32 + 32 <= p.T;
      const int tok = t * QT + r;
      const float inv = 1.0f / l;
      if (full_rows) {
        tma_store_wait_read();     // this warp's previous bulk stores have finished reading the staging tiles
        __syncwarp();
      }
#pragma unroll
      for (int hx = 0; hx < 2; ++hx) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(o_tmem + hx * 32, o);
        tmem_ld_wait();
        uint4 w[4];
#pragma unroll
        for (int jv = 0; jv < 4; ++jv) {
          w[jv].x = pack_bf16x2(__uint_as_float(o[jv * 8 + 0]) * inv, __uint_as_float(o[jv * 8 + 1]) * inv);
          w[jv].y = pack_bf16x2(__uint_as_float(o[jv * 8 + 2]) * inv, __uint_as_float(o[jv * 8 + 3]) * inv);
          w[jv].z = pack_bf16x2(__uint_as_float(o[jv * 8 + 4]) * inv, __uint_as_float(o[jv * 8 + 5]) * inv);
          w[jv].w = pack_bf16x2(__uint_as_float(o[jv * 8 + 6]) * inv, __uint_as_float(o[jv * 8 + 7]) * inv);
        }
        if (full_rows) {
          // 64-byte swizzle (CU_TENSOR_MAP_SWIZZLE_64B): 16-byte chunk i of row `lane` sits at chunk i ^ ((lane>>1)&3)
#pragma unroll
          for (int jv = 0; jv < 4; ++jv)
            *reinterpret_cast<uint4*>(out_stage + hx * 2048 + lane * 64 + ((jv ^ ((lane >> 1) & 3)) << 4)) = w[jv];
        } else if (tok < p.T) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t(b) * p.T + tok) * (p.H * HD) + h * HD + hx * 32);
#pragma unroll
          for (int jv = 0; jv < 4; ++jv) dst[jv] = w[jv];
        }
      }
      if (full_rows) {
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          const int row = b * p.T + t * QT + q * 32;
          tma_store_2d(&tmOut, out_stage, h * HD, row);                // 32 rows x 32 columns each
          tma_store_2d(&tmOut, out_stage + 2048, h * HD + 32, row);
          tma_store_commit();
        }
      }
    };
    // P of a group is published one group late: its tcgen05.st completes under the next group's exponentials
    auto publish = [&](uint32_t jj, int g) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[(st * 2 + (jj & 1)) * 3 + g]);
    };

    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int pair = item / p.npq, tp = item - pair * p.npq;
      const int b = pair / p.H, h = pair - b * p.H;
      const int t = 2 * tp + st;
      const bool warp_active = t * QT + q * 32 < p.T;   // (false for a tile past the last one and for empty row quarters)
      float m_run = -INFINITY, l_run = 0.f;
      for (int kb = 0; kb < p.nkb; ++kb, ++j) {
        const uint32_t buf = j & 1;
        const uint32_t sbase = s_tmem0 + buf * KB;
        const int key_base = kb * KB;
        const int nkeys = block_keys(kb);
        const int valid = p.T - key_base;   // keys of this block below T (may exceed nkeys)
        const bool full_block = nkeys == KB && valid >= KB;
        uint32_t s0[32], s1[32], s2[32];
        float m = -INFINITY;
        // ---- pass 1: block row max
        mbar_wait(&s_full[st * 2 + buf], (j >> 1) & 1);
        tc_fence_after();
        if (warp_active) {
          if (full_block) {
            tmem_ld_32x32b_x32(sbase, s0);
            tmem_ld_32x32b_x32(sbase + 32, s1);
            tmem_ld_32x32b_x32(sbase + 64, s2);
            tmem_ld_wait();
            m = max_group<false>(s0, 32, m);
            m = max_group<false>(s1, 32, m);
            m = max_group<false>(s2, 32, m);
          } else {
            // last block of the sequence: 16 .. 96 keys of which `valid` are real; absent groups hold -inf
            auto load_group = [&](int g, uint32_t (&v)[32]) {
              const int c0 = g * 32;
              if (c0 + 32 <= nkeys) {
                tmem_ld_32x32b_x32(sbase + c0, v);
              } else if (c0 < nkeys) {
                tmem_ld_32x32b_x16(sbase + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
#pragma unroll
                for (int i = 16; i < 32; ++i) v[i] = 0xff800000u;
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0xff800000u;
              }
            };
            load_group(0, s0); load_group(1, s1); load_group(2, s2);
            tmem_ld_wait();
            m = max_group<true>(s0, valid, m);
            m = max_group<true>(s1, valid - 32, m);
            m = max_group<true>(s2, valid - 64, m);
          }
        }
        // LAZY reference maximum (the FlashAttention-4 rule): the row's reference m_run moves (and O / l are rescaled) only
        // when the block maximum exceeds it by more than 2^LAZY_TAU in the exponent; until then P = exp2((s - m_run) c) may
        // reach 2^LAZY_TAU instead of 1 -- the same relative bf16 precision, and O / l at the end does not care which
        // reference the sums were taken against.  Rescaling needs the previous P.V to have COMPLETED, a stall of a few
        // hundred cycles per block on the softmax warps' critical path; with the threshold it is paid a few times per row
        // instead of in almost every block (32 rows per warp: some row raises its maximum in nearly every block).
        const bool grow = (m - m_run) * p.sl2 > LAZY_TAU;      // true for the first block (m_run = -inf)
        const float m_new = grow ? m : m_run;
        const float alpha = grow ? ex2((m_run - m_new) * p.sl2) : 1.0f;   // 0 for the first block
        const float msl = m_new * p.sl2;
        const bool rescale = kb > 0 && warp_active && __any_sync(0xffffffffu, grow);
        // ---- O slot: rescale by alpha (kb > 0, only when a row of this warp moved its reference) or hand the previous
        //      item over to HBM (kb == 0)
        if (j > 0 && (kb == 0 || rescale)) {
          mbar_wait(&o_full[st * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
          tc_fence_after();
          if (kb > 0) {
            {
#pragma unroll
              for (int hx = 0; hx < 2; ++hx) {
                uint32_t o[32];
                tmem_ld_32x32b_x32(o_tmem + hx * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st_32x32b_x16(o_tmem + hx * 32, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
                tmem_st_32x32b_x16(o_tmem + hx * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&o[16]));
              }
              tmem_st_wait();
            }
          } else {
            store_item(prev_b, prev_h, prev_t, prev_l);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_ready[st * 2 + (j & 1)]);
        // ---- pass 2: exponentials against the running max, row sum, bf16 P into TMEM over the consumed logits
        float l = 0.f;
        uint32_t pk[16];
        if (!warp_active) {
          // nothing to compute for these rows: the issuer still waits for the group barriers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int g = 0; g < NGROUPS; ++g) mbar_arrive(&p_full[(st * 2 + (j & 1)) * 3 + g]);
          }
        } else if (full_block) {
          l += exp_group<false, POLY>(s0, p.sl2, msl, 32, pk);
          tmem_st_32x32b_x16(sbase, pk);
          l += exp_group<false, POLY>(s1, p.sl2, msl, 32, pk);
          publish(j, 0);
          tmem_st_32x32b_x16(sbase + 32, pk);
          l += exp_group<false, POLY>(s2, p.sl2, msl, 32, pk);
          publish(j, 1);
          tmem_st_32x32b_x16(sbase + 64, pk);
          publish(j, 2);
        } else {
#pragma unroll
          for (int g = 0; g < NGROUPS; ++g) {
            const int c0 = g * 32;
            if (c0 < nkeys) {
              uint32_t(&v)[32] = g == 0 ? s0 : (g == 1 ? s1 : s2);
              l += exp_group<true>(v, p.sl2, msl, valid - c0, pk);
              if (c0 + 32 <= nkeys) {
                tmem_st_32x32b_x16(sbase + c0, pk);
              } else {
                tmem_st_32x32b_x8(sbase + c0, *reinterpret_cast<uint32_t(*)[8]>(&pk[0]));
              }
            }
            publish(j, g);   // (also for a group without keys: its barrier completes one phase per use)
          }
        }
        l_run = l_run * alpha + l;
        m_run = m_new;
      }
      prev_l = l_run; prev_b = b; prev_h = h; prev_t = t;
    }
    if (j > 0) {
      mbar_wait(&o_full[st * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
      tc_fence_after();
      store_item(prev_b, prev_h, prev_t, prev_l);
    }
    tma_store_wait_all();   // the staging tiles must outlive the bulk stores reading them
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

int attention_pair_bf16(const bf16* qkv, bf16* out, int B, int T, int H, float scale, cudaStream_t stream) {
  FP_REQUIRE(B > 0 && H > 0 && T > 0, "attention: empty problem");
  const int tpad = (T + 15) / 16 * 16;
  const int C = 3 * H * HD;
  CUtensorMap tmQ, tmKV, tmOut;
  const uint64_t rows = uint64_t(B) * T;
  if (int rc = make_tmap_2d_bf16(&tmQ, qkv, rows, uint64_t(C), uint64_t(C), QT, HD)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmKV, qkv, rows, uint64_t(C), uint64_t(C), KB, HD)) return rc;
  if (int rc = make_tmap_2d_bf16_sw64(&tmOut, out, rows, uint64_t(H) * HD, uint64_t(H) * HD, 32)) return rc;
  Params p;
  p.out = out; p.B = B; p.T = T; p.H = H;
  p.tpad = tpad;
  p.nq = (T + QT - 1) / QT;
  p.npq = (p.nq + 1) / 2;
  p.nkb = (tpad + KB - 1) / KB;
  p.sl2 = scale * 1.4426950408889634f;
  const long long nitems = (long long)B * H * p.npq;
  const int grid = nitems < sm_count() ? int(nitems) : sm_count();
  ProfScope prof(PROF_ATTENTION, 4.0 * double(B) * H * double(T) * T * HD, 1, stream);
  // FP_ATTN_POLY=0x1111 / 0x5555: 25 / 50 % of the exponentials on the FMA pipe (degree-4 polynomial) instead of the MUFU
  static const unsigned poly = [] { const char* e = getenv("FP_ATTN_POLY"); return e ? unsigned(strtoul(e, nullptr, 0)) : 0u; }();
#define FP_LAUNCH_PAIR(MASK_)                                                           \
  do {                                                                                  \
    auto kern = attention_pair_kernel<MASK_>;                                           \
    FP_ENSURE_DYN_SMEM(kern, SMEM_BYTES);                                               \
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, tmOut, p);               \
  } while (0)
  switch (poly) {
    case 0x1111u: FP_LAUNCH_PAIR(0x1111u); break;
    case 0x5555u: FP_LAUNCH_PAIR(0x5555u); break;
    case 0x0101u: FP_LAUNCH_PAIR(0x0101u); break;
    default: FP_LAUNCH_PAIR(0u); break;
  }
#undef FP_LAUNCH_PAIR
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fp
