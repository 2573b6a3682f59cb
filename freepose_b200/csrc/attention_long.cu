// ViT self-attention for crops above 224^2 (reference-native 420^2 -> T = 905 tokens, the refiner's 518^2 -> 1374;
// SURVEY.md section 0.1): keys no longer fit one TMEM accumulator, so every 128-row query tile walks over key blocks of
// 256 with an ONLINE softmax (running row max m and row sum l; the O accumulator in TMEM is rescaled by
// exp2((m_old - m_new) * c) between blocks).
//
//   warp 0       TMA loader   Q tile per work item; K blocks through a 3-stage ring, V blocks through a 2-stage ring
//                             (K of block j+1 is needed while block j is still being exponentiated)
//   warp 1       MMA issuer   S = Q K_blk^T in two key parts (M128 x N128 x K64 each): a part of the NEXT block is issued as
//                             soon as the P.V steps reading the P stored over its columns have been issued, so the next
//                             block's logits are ready when the softmax warps finish the current one
//                             O += P V_blk (P read straight from TMEM -- "TS" form --, V as MN-major smem operand)
//   warp 2       TMEM allocator
//   warps 4..11  softmax      two warps per TMEM lane quarter split the key columns of a row.  Full blocks: the warp's 128
//                             logits are read once, three of the four 32-column groups stay in registers between the max
//                             and the exponential pass; bf16 P is stored over the logits it came from and published one
//                             chunk late.  The O rescale of block b (or, for the first block of an item, the previous
//                             item's normalisation + store) sits between the two passes, where P.V of the previous
//                             block has had a whole max pass to finish.  Row sums stay per warp until the item ends.
//
// Work item = (image, head, query tile).  Arithmetic contract: oracle/vit.py contract_attention with key_block = 256 --
// block-wise flash attention: P of block b is exp2((s - m_b) c) rounded to bf16, where m_b is the running max after
// block b.  (The rescale is skipped when no row of the warp raised its max: multiplying by exactly 1 changes nothing.)
#include "attention_common.cuh"
#include "kernels.h"

namespace fp {

namespace {

using namespace attn;

constexpr int HD = 64;
constexpr int QT = 128;
constexpr int KB = 256;                       // keys per block
constexpr int ROW_BYTES = HD * 2;
constexpr int Q_TILE_BYTES = QT * ROW_BYTES;  // 16 KB
constexpr int KV_BLOCK_BYTES = KB * ROW_BYTES;  // 32 KB
constexpr int K_STAGES = 3, V_STAGES = 2;
constexpr int NUM_SOFTMAX_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_SOFTMAX_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int S_COL = 0;     // 256 fp32 logit columns; bf16x2 P overwrites the first half of every consumed 32-column group
constexpr int O_COL = 256;   // 64 fp32 columns
constexpr int S_PART = 128;  // S is issued as keys [0,128) and [128,256) of the block
constexpr int NCHUNKS = KB / 64;

constexpr int OFF_Q = 0;                                   // 2 slots
constexpr int OFF_K = OFF_Q + 2 * Q_TILE_BYTES;            // 3 stages
constexpr int OFF_V = OFF_K + K_STAGES * KV_BLOCK_BYTES;   // 2 stages
constexpr int OFF_XCH = OFF_V + V_STAGES * KV_BLOCK_BYTES; // float [2 parities][2 halves][128] max + the same for the item sums
constexpr int OFF_BAR = OFF_XCH + 8 * 128 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;

struct Params {
  bf16* out;
  int B, T, H;
  int tpad;   // keys padded to a multiple of 16
  int nq;     // query tiles per (image, head)
  int nkb;    // key blocks
  float sl2;  // scale * log2(e)
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_long_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* k_full = bars;         // [3]
  uint64_t* k_empty = bars + 3;    // [3]
  uint64_t* v_full = bars + 6;     // [2]
  uint64_t* v_empty = bars + 8;    // [2]
  uint64_t* q_full = bars + 10;    // [2]
  uint64_t* q_empty = bars + 12;   // [2]
  uint64_t* s_full = bars + 14;    // [2] key parts of S
  uint64_t* o_full = bars + 16;    // MMA -> softmax: P.V of a key block complete
  uint64_t* o_ready = bars + 17;   // softmax -> MMA: O rescaled / read out, this block's P.V may accumulate
  uint64_t* p_full = bars + 18;    // [NCHUNKS]; all eight arrivals on chunk c also mean "S columns of chunk c are consumed"
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 18 + NCHUNKS);
  float* xch_max = reinterpret_cast<float*>(smem + OFF_XCH);   // [2][2][128]
  float* xch_sum = xch_max + 4 * 128;                           // [2][2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = p.B * p.H * p.nq;
  const int my_items = (nitems - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < K_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(o_full, 1); mbar_init(o_ready, NUM_SOFTMAX_WARPS);
    for (int i = 0; i < NCHUNKS; ++i) mbar_init(&p_full[i], NUM_SOFTMAX_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ---------------------------------------------------------------------------- TMA loader
    if (elect_one()) {
      uint32_t j = 0, qi = 0;   // block / item counters of this CTA
      int ks = 0;               // j % K_STAGES
      uint32_t kphase = 0;      // (j / K_STAGES) & 1
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++qi) {
        const int pair = item / p.nq, t = item - pair * p.nq;
        const int b = pair / p.H, h = pair - b * p.H;
        const int row0 = b * p.T;
        const int slot = qi & 1;
        mbar_wait(&q_empty[slot], ((qi >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[slot], Q_TILE_BYTES);
        tma_load_2d(smem + OFF_Q + slot * Q_TILE_BYTES, &tmQ, &q_full[slot], h * HD, row0 + t * QT);
        const int kcol = p.H * HD + h * HD, vcol = 2 * p.H * HD + h * HD;
        for (int kb = 0; kb < p.nkb; ++kb, ++j) {
          const int krow = row0 + kb * KB;
          mbar_wait(&k_empty[ks], kphase ^ 1);
          mbar_arrive_expect_tx(&k_full[ks], KV_BLOCK_BYTES);
          uint8_t* sK = smem + OFF_K + ks * KV_BLOCK_BYTES;
          tma_load_2d(sK, &tmKV, &k_full[ks], kcol, krow);
          tma_load_2d(sK + 128 * ROW_BYTES, &tmKV, &k_full[ks], kcol, krow + 128);
          const int vs = j & 1;
          mbar_wait(&v_empty[vs], ((j >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[vs], KV_BLOCK_BYTES);
          uint8_t* sV = smem + OFF_V + vs * KV_BLOCK_BYTES;
          tma_load_2d(sV, &tmKV, &v_full[vs], vcol, krow);
          tma_load_2d(sV + 128 * ROW_BYTES, &tmKV, &v_full[vs], vcol, krow + 128);
          if (++ks == K_STAGES) { ks = 0; kphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------- MMA issuer
    if (elect_one()) {   // (see attention.cu: a single elected thread lets the compiler emit UTCHMMA back to back)
      const uint32_t idesc_pv = umma_idesc_bf16(QT, HD, 0, 1);
      const uint64_t q_desc0 = umma_smem_desc_sw128(smem_u32(smem + OFF_Q), 16, 1024);
      const uint64_t k_desc0 = umma_smem_desc_sw128(smem_u32(smem + OFF_K), 16, 1024);
      const uint64_t v_desc0 = umma_smem_desc_sw128(smem_u32(smem + OFF_V), 1024, 1024);
      const uint32_t nblocks = uint32_t(my_items) * uint32_t(p.nkb);
      // state of the block whose S is issued next ("n"): item counter, key block, K stage / phase
      uint32_t n_qi = 0;
      int n_kb = 0, n_ks = 0;
      uint32_t n_kphase = 0;
      auto block_keys = [&](int kb) { return (p.tpad - kb * KB) < KB ? (p.tpad - kb * KB) : KB; };   // multiple of 16
      // S part `part` of block "n"; after the block's last part the state advances to the following block
      auto issue_s = [&](int part) {
        const int nkeys = block_keys(n_kb);
        const int n = nkeys - part * S_PART < S_PART ? nkeys - part * S_PART : S_PART;   // keys of this part (may be <= 0)
        const int slot = n_qi & 1;
        if (part == 0) {
          if (n_kb == 0) mbar_wait(&q_full[slot], (n_qi >> 1) & 1);
          mbar_wait(&k_full[n_ks], n_kphase);
          tc_fence_after();
        }
        if (n > 0) {
          const uint64_t q_desc = q_desc0 + uint64_t(slot * (Q_TILE_BYTES >> 4));
          const uint64_t k_desc = k_desc0 + uint64_t(n_ks * (KV_BLOCK_BYTES >> 4) + part * ((S_PART * ROW_BYTES) >> 4));
          const uint32_t idesc = umma_idesc_bf16(QT, n, 0, 0);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_bf16_ss(tmem_base + S_COL + part * S_PART, q_desc + uint64_t(2 * k), k_desc + uint64_t(2 * k), idesc, k != 0);
        }
        umma_commit(&s_full[part]);
        if (part == 1) {
          umma_commit(&k_empty[n_ks]);                              // K of this block is no longer needed
          if (n_kb == p.nkb - 1) umma_commit(&q_empty[slot]);       // nor is the item's Q tile
          if (++n_ks == K_STAGES) { n_ks = 0; n_kphase ^= 1; }
          if (++n_kb == p.nkb) { n_kb = 0; ++n_qi; }
        }
      };
      if (nblocks > 0) { issue_s(0); issue_s(1); }
      int kb = 0;
      for (uint32_t j = 0; j < nblocks; ++j) {
        const int vs = j & 1;
        const int nkeys = block_keys(kb);
        const bool has_next = j + 1 < nblocks;
        const uint64_t v_desc = v_desc0 + uint64_t(vs * (KV_BLOCK_BYTES >> 4));
        mbar_wait(&v_full[vs], (j >> 1) & 1);
        mbar_wait(o_ready, j & 1);        // O rescaled for this block's running max (or read out by the previous item)
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < NCHUNKS; ++c) {
          mbar_wait(&p_full[c], j & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int key0 = c * 64 + k * 16;
            if (key0 < nkeys) {
              // P of keys [key0, key0+16): 8 packed columns at the start of the 32-column logit group they came from
              const uint32_t pcol = uint32_t((key0 & ~31) + ((key0 & 16) >> 1));
              umma_bf16_ts(tmem_base + O_COL, tmem_base + S_COL + pcol, v_desc + uint64_t(key0 * (ROW_BYTES >> 4)), idesc_pv,
                           (kb | key0) != 0);
            }
          }
          if (c == NCHUNKS - 1) {
            umma_commit(o_full);
            umma_commit(&v_empty[vs]);
          }
          if (has_next && (c & 1)) issue_s(c >> 1);   // after chunks 1 / 3: key part 0 / 1 of the next block
        }
        if (++kb == p.nkb) kb = 0;
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------------------- softmax + epilogue
    const int q = warp & 3;
    const int hf = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t sbase = tmem_base + lane_addr + S_COL + hf * 32;
    const uint32_t obase = tmem_base + lane_addr + O_COL + hf * 32;
    uint32_t j = 0;        // block counter of this CTA
    uint32_t it = 0;       // item counter
    // the previous item's normalisation + store is owed until the next item's first block (or the end)
    float prev_l = 0.f;
    int prev_b = 0, prev_h = 0, prev_t = 0;

    // O / l -> bf16 -> HBM for item (b, h, t): this warp writes 32 of the 64 output columns of its 32 rows
    auto store_item = [&](int b, int h, int t, float l_own, uint32_t item_par) {
      const float l = l_own + xch_sum[item_par * 256 + (hf ^ 1) * 128 + r];
      const bool warp_active = t * QT + q * 32 < p.T;
      const int tok = t * QT + r;
      if (warp_active) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(obase, o);
        tmem_ld_wait();
        if (tok < p.T) {
          const float inv = 1.0f / l;
          uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t(b) * p.T + tok) * (p.H * HD) + h * HD + hf * 32);
#pragma unroll
          for (int jv = 0; jv < 4; ++jv) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[jv * 8 + 0]) * inv, __uint_as_float(o[jv * 8 + 1]) * inv);
            w.y = pack_bf16x2(__uint_as_float(o[jv * 8 + 2]) * inv, __uint_as_float(o[jv * 8 + 3]) * inv);
            w.z = pack_bf16x2(__uint_as_float(o[jv * 8 + 4]) * inv, __uint_as_float(o[jv * 8 + 5]) * inv);
            w.w = pack_bf16x2(__uint_as_float(o[jv * 8 + 6]) * inv, __uint_as_float(o[jv * 8 + 7]) * inv);
            dst[jv] = w;
          }
        }
      }
    };
    // P of a chunk is published one chunk late: its tcgen05.st completes under the next chunk's exponentials
    auto publish = [&](int c) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[c]);
    };

    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      const int pair = item / p.nq, t = item - pair * p.nq;
      const int b = pair / p.H, h = pair - b * p.H;
      const bool warp_active = t * QT + q * 32 < p.T;
      float m_run = -INFINITY, l_run = 0.f;   // l_run: this warp's columns only; the halves meet when the item is stored
      for (int kb = 0; kb < p.nkb; ++kb, ++j) {
        const uint32_t par = j & 1;
        const int key_base = kb * KB;
        const int nkeys = (p.tpad - key_base) < KB ? (p.tpad - key_base) : KB;
        const int valid = p.T - key_base;  // keys of this block below T (may exceed nkeys)
        const bool full_block = nkeys == KB && valid >= KB && warp_active;   // warp-uniform
        uint32_t s0[32], s1[32], s2[32], s3[32];
        // ---- pass 1: block row max
        float m = -INFINITY;
        mbar_wait(&s_full[0], par);
        tc_fence_after();
        if (full_block) {
          tmem_ld_32x32b_x32(sbase, s0);
          tmem_ld_32x32b_x32(sbase + 64, s1);
          mbar_wait(&s_full[1], par);
          tc_fence_after();
          tmem_ld_32x32b_x32(sbase + 128, s2);
          tmem_ld_32x32b_x32(sbase + 192, s3);
          tmem_ld_wait();
          m = max_group<false>(s3, 32, m);      // group 3 is read a second time in pass 2
          m = max_group<false>(s0, 32, m);
          m = max_group<false>(s1, 32, m);
          m = max_group<false>(s2, 32, m);
        } else {
          mbar_wait(&s_full[1], par);
          tc_fence_after();
          if (warp_active) {
            for (int g = hf; g * 32 < nkeys; g += 2) {
              const int c0 = g * 32;
              uint32_t v[32];
              if (nkeys - c0 >= 32) {
                tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL + c0, v);
              } else {
                uint32_t w16[16];
                tmem_ld_32x32b_x16(tmem_base + lane_addr + S_COL + c0, w16);
#pragma unroll
                for (int i = 0; i < 16; ++i) { v[i] = w16[i]; v[16 + i] = 0xff800000u; }
              }
              tmem_ld_wait();
              if (c0 + 32 <= valid) m = max_group<false>(v, 32, m);
              else                  m = max_group<true>(v, valid - c0, m);
            }
          }
        }
        xch_max[par * 256 + hf * 128 + r] = m;
        named_bar_sync(1 + q, 64);
        m = fmaxf(xch_max[par * 256 + r], xch_max[par * 256 + 128 + r]);
        const float m_new = fmaxf(m_run, m);
        const float alpha = ex2((m_run - m_new) * p.sl2);  // 0 for the first block (m_run = -inf)
        const float msl = m_new * p.sl2;
        // ---- O slot: rescale by alpha (kb > 0) or hand the previous item over to HBM (kb == 0); either way the previous
        //      P.V has had the whole max pass to finish
        if (j > 0) {
          mbar_wait(o_full, par ^ 1);
          tc_fence_after();
          if (kb > 0) {
            if (warp_active && __any_sync(0xffffffffu, alpha != 1.0f)) {
              uint32_t o[32];
              tmem_ld_32x32b_x32(obase, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32b_x16(obase, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
              tmem_st_32x32b_x16(obase + 16, *reinterpret_cast<uint32_t(*)[16]>(&o[16]));
              tmem_st_wait();
            }
          } else {
            store_item(prev_b, prev_h, prev_t, prev_l, (it - 1) & 1);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_ready);
        // ---- pass 2: exponentials against the running max, row sum, bf16 P into TMEM over the consumed logits
        float l = 0.f;
        if (full_block) {
          uint32_t pk[16];
          l += exp_group<false>(s0, p.sl2, msl, 32, pk);
          tmem_st_32x32b_x16(sbase, pk);
          tmem_ld_32x32b_x32(sbase + 192, s3);          // second read of group 3, under the exponentials of 1 and 2
          l += exp_group<false>(s1, p.sl2, msl, 32, pk);
          publish(0);
          tmem_st_32x32b_x16(sbase + 64, pk);
          l += exp_group<false>(s2, p.sl2, msl, 32, pk);
          publish(1);
          tmem_st_32x32b_x16(sbase + 128, pk);
          tmem_ld_wait();
          l += exp_group<false>(s3, p.sl2, msl, 32, pk);
          publish(2);
          tmem_st_32x32b_x16(sbase + 192, pk);
          publish(3);
        } else {
          for (int c = 0; c < NCHUNKS; ++c) {
            const int c0 = c * 64 + hf * 32;
            if (warp_active && c0 < nkeys) {
              const int width = nkeys - c0 >= 32 ? 32 : 16;
              uint32_t v[32], pk[16];
              if (width == 32) {
                tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL + c0, v);
              } else {
                uint32_t w16[16];
                tmem_ld_32x32b_x16(tmem_base + lane_addr + S_COL + c0, w16);
#pragma unroll
                for (int i = 0; i < 16; ++i) { v[i] = w16[i]; v[16 + i] = 0xff800000u; }
              }
              tmem_ld_wait();
              if (c0 + 32 <= valid) l += exp_group<false>(v, p.sl2, msl, 32, pk);
              else                  l += exp_group<true>(v, p.sl2, msl, valid - c0, pk);
              if (width == 32) {
                tmem_st_32x32b_x16(tmem_base + lane_addr + S_COL + c0, pk);
              } else {
                uint32_t pk8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) pk8[i] = pk[i];
                tmem_st_32x32b_x8(tmem_base + lane_addr + S_COL + c0, pk8);
              }
              tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[c]);
          }
        }
        l_run = l_run * alpha + l;
        m_run = m_new;
      }
      // the item's row sums meet in the next item's O slot (after that block's max-exchange barrier) or after the loop
      xch_sum[(it & 1) * 256 + hf * 128 + r] = l_run;
      prev_l = l_run; prev_b = b; prev_h = h; prev_t = t;
    }
    if (j > 0) {
      named_bar_sync(1 + q, 64);              // the partner's row sums of the last item
      mbar_wait(o_full, (j - 1) & 1);
      tc_fence_after();
      store_item(prev_b, prev_h, prev_t, prev_l, (it - 1) & 1);
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

int attention_long_bf16(const bf16* qkv, bf16* out, int B, int T, int H, float scale, cudaStream_t stream) {
  FP_REQUIRE(B > 0 && H > 0 && T > 0, "attention: empty problem");
  const int tpad = (T + 15) / 16 * 16;
  const int C = 3 * H * HD;
  CUtensorMap tmQ, tmKV;
  const uint64_t rows = uint64_t(B) * T;
  if (int rc = make_tmap_2d_bf16(&tmQ, qkv, rows, uint64_t(C), uint64_t(C), QT, HD)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmKV, qkv, rows, uint64_t(C), uint64_t(C), 128, HD)) return rc;
  Params p;
  p.out = out; p.B = B; p.T = T; p.H = H;
  p.tpad = tpad;
  p.nq = (T + QT - 1) / QT;
  p.nkb = (tpad + KB - 1) / KB;
  p.sl2 = scale * 1.4426950408889634f;
  FP_ENSURE_DYN_SMEM(attention_long_kernel, SMEM_BYTES);
  const long long nitems = (long long)B * H * p.nq;
  const int grid = nitems < sm_count() ? int(nitems) : sm_count();
  ProfScope prof(PROF_ATTENTION, 4.0 * double(B) * H * double(T) * T * HD, 1, stream);
  attention_long_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, p);
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fp
