// ViT self-attention for crops above 224^2 (reference-native 420^2 -> T = 905 tokens): keys no
// longer fit one TMEM accumulator, so every 128-row query tile walks over key blocks of 256 with an ONLINE softmax
// (running row max m and row sum l; the O accumulator in TMEM is rescaled by exp2((m_old - m_new) * c) between blocks).
//
//   warp 0       TMA loader   Q tile per work item, K/V blocks through a 2-stage ring (64 KB per stage)
//   warp 1       MMA issuer   S = Q K_blk^T (M128 x N<=256 x K64) ; O += P V_blk (P from TMEM, V MN-major smem)
//   warp 2       TMEM allocator
//   warps 4..11  softmax      as in attention.cu (two warps per TMEM lane quarter), plus the O rescale
//
// Work item = (image, head, query tile).  Arithmetic contract: oracle/vit.py contract_attention with
// key_block = 256 -- block-wise flash attention: P of block b is exp2((s - m_b) c) rounded to bf16, where m_b is the
// running max after block b.
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
constexpr int NUM_SOFTMAX_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_SOFTMAX_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int S_COL = 0;     // 256 fp32 columns
constexpr int P_COL = 256;   // 128 columns of packed bf16x2
constexpr int O_COL = 384;   // 64 fp32 columns
constexpr int NCHUNKS = KB / 64;

constexpr int OFF_Q = 0;                             // 2 slots
constexpr int OFF_K = OFF_Q + 2 * Q_TILE_BYTES;      // 2 stages
constexpr int OFF_V = OFF_K + 2 * KV_BLOCK_BYTES;    // 2 stages
constexpr int OFF_XCH = OFF_V + 2 * KV_BLOCK_BYTES;  // float [2][128] max + [2][128] sum
constexpr int OFF_BAR = OFF_XCH + 4 * 128 * 4;
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
  uint64_t* kv_full = bars;        // [2]
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* q_full = bars + 4;     // [2]
  uint64_t* q_empty = bars + 6;    // [2]
  uint64_t* s_full = bars + 8;
  uint64_t* s_empty = bars + 9;
  uint64_t* o_full = bars + 10;    // MMA -> softmax: P.V of a key block complete
  uint64_t* o_ready = bars + 11;   // softmax -> MMA: O rescaled (or nothing to rescale), next P.V may accumulate
  uint64_t* p_full = bars + 12;    // [NCHUNKS]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12 + NCHUNKS);
  float* xch_max = reinterpret_cast<float*>(smem + OFF_XCH);
  float* xch_sum = xch_max + 2 * 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = p.B * p.H * p.nq;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1); mbar_init(s_empty, NUM_SOFTMAX_WARPS);
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
      uint32_t kvi = 0, qi = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++qi) {
        const int pair = item / p.nq, t = item - pair * p.nq;
        const int b = pair / p.H, h = pair - b * p.H;
        const int row0 = b * p.T;
        const int slot = qi & 1;
        mbar_wait(&q_empty[slot], ((qi >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[slot], Q_TILE_BYTES);
        tma_load_2d(smem + OFF_Q + slot * Q_TILE_BYTES, &tmQ, &q_full[slot], h * HD, row0 + t * QT);
        for (int kb = 0; kb < p.nkb; ++kb, ++kvi) {
          const int st = kvi & 1;
          mbar_wait(&kv_empty[st], ((kvi >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[st], 2 * KV_BLOCK_BYTES);
          uint8_t* sK = smem + OFF_K + st * KV_BLOCK_BYTES;
          uint8_t* sV = smem + OFF_V + st * KV_BLOCK_BYTES;
          const int krow = row0 + kb * KB;
          const int kcol = p.H * HD + h * HD, vcol = 2 * p.H * HD + h * HD;
          tma_load_2d(sK, &tmKV, &kv_full[st], kcol, krow);
          tma_load_2d(sK + 128 * ROW_BYTES, &tmKV, &kv_full[st], kcol, krow + 128);
          tma_load_2d(sV, &tmKV, &kv_full[st], vcol, krow);
          tma_load_2d(sV + 128 * ROW_BYTES, &tmKV, &kv_full[st], vcol, krow + 128);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------- MMA issuer
    if (elect_one()) {   // (see attention.cu: a single elected thread lets the compiler emit UTCHMMA back to back)
      const uint32_t idesc_pv = umma_idesc_bf16(QT, HD, 0, 1);
      uint32_t kvi = 0, qi = 0, blk_iter = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++qi) {
        const int slot = qi & 1;
        const uint32_t sQ = smem_u32(smem + OFF_Q + slot * Q_TILE_BYTES);
        mbar_wait(&q_full[slot], (qi >> 1) & 1);
        for (int kb = 0; kb < p.nkb; ++kb, ++kvi, ++blk_iter) {
          const int st = kvi & 1;
          const uint32_t sK = smem_u32(smem + OFF_K + st * KV_BLOCK_BYTES);
          const uint32_t sV = smem_u32(smem + OFF_V + st * KV_BLOCK_BYTES);
          const int nkeys = (p.tpad - kb * KB) < KB ? (p.tpad - kb * KB) : KB;  // multiple of 16
          mbar_wait(&kv_full[st], (kvi >> 1) & 1);
          mbar_wait(s_empty, (blk_iter & 1) ^ 1);
          tc_fence_after();
          const uint32_t idesc_s = umma_idesc_bf16(QT, nkeys, 0, 0);
          const uint64_t q_desc = umma_smem_desc_sw128(sQ, 16, 1024);
          const uint64_t k_desc = umma_smem_desc_sw128(sK, 16, 1024);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_bf16_ss(tmem_base + S_COL, q_desc + uint64_t(2 * k), k_desc + uint64_t(2 * k), idesc_s, k != 0);
          umma_commit(s_full);
          if (kb == p.nkb - 1) umma_commit(&q_empty[slot]);
          // ---- O (+)= P V_blk once the softmax warps have rescaled O for this block's running max
          mbar_wait(o_ready, blk_iter & 1);
          tc_fence_after();
          for (int c = 0; c < NCHUNKS; ++c) {
            mbar_wait(&p_full[c], blk_iter & 1);
            tc_fence_after();
            const int keys = (nkeys - c * 64) < 64 ? (nkeys - c * 64) : 64;
            for (int k = 0; k * 16 < keys; ++k) {
              const int key0 = c * 64 + k * 16;
              const uint64_t v_desc = umma_smem_desc_sw128(sV + uint32_t(key0) * ROW_BYTES, 1024, 1024);
              umma_bf16_ts(tmem_base + O_COL, tmem_base + P_COL + uint32_t(key0 >> 1), v_desc, idesc_pv,
                           (kb | key0) != 0);
            }
          }
          umma_commit(o_full);
          umma_commit(&kv_empty[st]);
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------------------- softmax + epilogue
    const int q = warp & 3;
    const int hf = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    uint32_t blk_iter = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int pair = item / p.nq, t = item - pair * p.nq;
      const int b = pair / p.H, h = pair - b * p.H;
      const bool warp_active = t * QT + q * 32 < p.T;
      const int tok = t * QT + r;
      float m_run = -INFINITY, l_run = 0.f;
      for (int kb = 0; kb < p.nkb; ++kb, ++blk_iter) {
        const uint32_t par = blk_iter & 1;
        const int key_base = kb * KB;
        const int nkeys = (p.tpad - key_base) < KB ? (p.tpad - key_base) : KB;
        const int valid = p.T - key_base;  // keys of this block below T (may exceed nkeys)
        mbar_wait(s_full, par);
        tc_fence_after();
        // ---- pass 1: block row max
        float m = -INFINITY;
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
              for (int j = 0; j < 16; ++j) { v[j] = w16[j]; v[16 + j] = 0xff800000u; }
            }
            tmem_ld_wait();
            if (c0 + 32 <= valid) m = max_group<false>(v, 32, m);
            else                  m = max_group<true>(v, valid - c0, m);
          }
        }
        xch_max[hf * 128 + r] = m;
        named_bar_sync(1 + q, 64);
        m = fmaxf(xch_max[r], xch_max[128 + r]);
        const float m_new = fmaxf(m_run, m);
        const float alpha = ex2((m_run - m_new) * p.sl2);  // 0 for the first block (m_run = -inf)
        const float msl = m_new * p.sl2;
        // ---- rescale O by alpha (its previous P.V must have finished), then let the MMA warp accumulate
        if (kb > 0) {
          mbar_wait(o_full, par ^ 1);  // completion of the previous block's P.V
          tc_fence_after();
          if (warp_active) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL + hf * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st_32x32b_x16(tmem_base + lane_addr + O_COL + hf * 32, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
            tmem_st_32x32b_x16(tmem_base + lane_addr + O_COL + hf * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&o[16]));
            tmem_st_wait();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_ready);
        // ---- pass 2: exponentials against the running max, row sum, bf16 P into TMEM
        float l = 0.f;
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
              for (int j = 0; j < 16; ++j) { v[j] = w16[j]; v[16 + j] = 0xff800000u; }
            }
            tmem_ld_wait();
            if (c0 + 32 <= valid) l += exp_group<false>(v, p.sl2, msl, 32, pk);
            else                  l += exp_group<true>(v, p.sl2, msl, valid - c0, pk);
            if (width == 32) {
              tmem_st_32x32b_x16(tmem_base + lane_addr + P_COL + (c0 >> 1), pk);
            } else {
              uint32_t pk8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) pk8[j] = pk[j];
              tmem_st_32x32b_x8(tmem_base + lane_addr + P_COL + (c0 >> 1), pk8);
            }
            tmem_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (c == NCHUNKS - 1 && lane == 0) mbar_arrive(s_empty);
          if (lane == 0) mbar_arrive(&p_full[c]);
        }
        xch_sum[hf * 128 + r] = l;
        named_bar_sync(1 + q, 64);
        l_run = l_run * alpha + (xch_sum[r] + xch_sum[128 + r]);
        m_run = m_new;
      }
      // ---- epilogue: O / l -> bf16 -> HBM (this warp: 32 of the 64 output columns)
      mbar_wait(o_full, (blk_iter - 1) & 1);
      tc_fence_after();
      if (warp_active) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL + hf * 32, o);
        tmem_ld_wait();
        if (tok < p.T) {
          const float inv = 1.0f / l_run;
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
      // the next item's first P.V overwrites O (accumulate flag off), ordered after these reads through o_ready
      tc_fence_before();
    }
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
  static bool attr_done = false;
  if (!attr_done) {
    FP_CUDA(cudaFuncSetAttribute(attention_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done = true;
  }
  const long long nitems = (long long)B * H * p.nq;
  const int grid = nitems < sm_count() ? int(nitems) : sm_count();
  ProfScope prof(PROF_ATTENTION, 4.0 * double(B) * H * double(T) * T * HD, 1, stream);
  attention_long_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, p);
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fp
