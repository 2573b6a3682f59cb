// ViT self-attention on tcgen05: one persistent CTA per SM walks over
// (image, head) pairs; the whole key set of one head (T <= 272 tokens: 261 at 224^2) lives in shared
// memory, so softmax is single-pass (no online rescale):
//
//   warp 0      TMA loader   K,V (double buffered across pairs) and the pair's Q tiles (128 rows each)
//   warp 1      MMA issuer   S = Q K^T  (M128 x N{256,+16} x K64, fp32 in TMEM)
//                            O = P V    (M128 x N64, P from a swizzled smem ring, V as MN-major operand)
//   warp 2      TMEM allocator
//   warps 4..7  softmax      thread = query row: row max, p = exp2((s - max) * scale*log2e) in fp32,
//                            row sum of the unrounded p, P rounded to bf16 into the smem ring in 64-key
//                            chunks (so P.V overlaps the exponentials), finally O / rowsum -> bf16 -> HBM.
//
// Arithmetic contract = flash/xformers attention (oracle/vit.py contract_attention): logits and softmax
// statistics in fp32, un-normalised P rounded to bf16 for the tensor-core P.V, one rounding of O.
#include "common.cuh"
#include "kernels.h"

namespace fp {

namespace {

constexpr int HD = 64;            // head dim
constexpr int QT = 128;           // query rows per tile
constexpr int MAX_TPAD = 272;     // padded key count (multiple of 16)
constexpr int MAX_QTILES = 3;
constexpr int ROW_BYTES = HD * 2;  // 128 B: one swizzle row
constexpr int Q_TILE_BYTES = QT * ROW_BYTES;       // 16 KB
constexpr int KV_BYTES = MAX_TPAD * ROW_BYTES;     // 34 KB
constexpr int P_CHUNK_KEYS = 64;
constexpr int P_CHUNK_BYTES = QT * 128;            // 16 KB
constexpr int P_STAGES = 2;
constexpr int NUM_THREADS = 256;
constexpr int TMEM_COLS = 512;
constexpr int S_COL = 0;
constexpr int O_COL = 384;

constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + MAX_QTILES * Q_TILE_BYTES;  // 2 buffers
constexpr int OFF_V = OFF_K + 2 * KV_BYTES;               // 2 buffers
constexpr int OFF_P = OFF_V + 2 * KV_BYTES;
constexpr int OFF_BAR = OFF_P + P_STAGES * P_CHUNK_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;

struct Params {
  bf16* out;
  int B, T, H;
  int tpad;      // keys padded to a multiple of 16
  int nq;        // query tiles per (image, head)
  int nchunks;   // 64-key P chunks
  float sl2;     // scale * log2(e)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_full = bars;        // [2]
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* q_full = bars + 4;
  uint64_t* q_empty = bars + 5;
  uint64_t* s_full = bars + 6;
  uint64_t* s_empty = bars + 7;
  uint64_t* o_full = bars + 8;
  uint64_t* o_empty = bars + 9;
  uint64_t* p_full = bars + 10;    // [P_STAGES]
  uint64_t* p_empty = bars + 12;   // [P_STAGES]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = p.B * p.H;
  const int half_rows = p.tpad / 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    mbar_init(s_full, 1); mbar_init(s_empty, 4);
    mbar_init(o_full, 1); mbar_init(o_empty, 4);
    for (int i = 0; i < P_STAGES; ++i) { mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ---------------------------------------------------------------------------- TMA loader
    if (lane == 0) {
      int it = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
        const int b = pair / p.H, h = pair - b * p.H;
        const int row0 = b * p.T;
        const int buf = it & 1;
        mbar_wait(&kv_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[buf], 2 * p.tpad * ROW_BYTES);
        uint8_t* sK = smem + OFF_K + buf * KV_BYTES;
        uint8_t* sV = smem + OFF_V + buf * KV_BYTES;
        const int kcol = p.H * HD + h * HD, vcol = 2 * p.H * HD + h * HD;
        tma_load_2d(sK, &tmKV, &kv_full[buf], kcol, row0);
        tma_load_2d(sK + half_rows * ROW_BYTES, &tmKV, &kv_full[buf], kcol, row0 + half_rows);
        tma_load_2d(sV, &tmKV, &kv_full[buf], vcol, row0);
        tma_load_2d(sV + half_rows * ROW_BYTES, &tmKV, &kv_full[buf], vcol, row0 + half_rows);
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, p.nq * Q_TILE_BYTES);
        for (int t = 0; t < p.nq; ++t)
          tma_load_2d(smem + OFF_Q + t * Q_TILE_BYTES, &tmQ, q_full, h * HD, row0 + t * QT);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const int n1 = p.tpad > 256 ? 256 : p.tpad;
      const int n2 = p.tpad - n1;
      const uint32_t idesc_s1 = umma_idesc_bf16(QT, n1, 0, 0);
      const uint32_t idesc_s2 = umma_idesc_bf16(QT, n2 > 0 ? n2 : 16, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(QT, HD, 0, 1);  // B (= V) is MN-major
      uint32_t s_iter = 0, p_iter = 0;
      int it = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t sK = smem_u32(smem + OFF_K + buf * KV_BYTES);
        const uint32_t sV = smem_u32(smem + OFF_V + buf * KV_BYTES);
        mbar_wait(&kv_full[buf], (it >> 1) & 1);
        mbar_wait(q_full, it & 1);
        tc_fence_after();
        for (int t = 0; t < p.nq; ++t, ++s_iter) {
          // ---- S = Q_t K^T
          mbar_wait(s_empty, (s_iter & 1) ^ 1);
          tc_fence_after();
          const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(smem + OFF_Q + t * Q_TILE_BYTES), 16, 1024);
          const uint64_t k_desc1 = umma_smem_desc_sw128(sK, 16, 1024);
          const uint64_t k_desc2 = umma_smem_desc_sw128(sK + 256 * ROW_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            umma_bf16_ss(tmem_base + S_COL, q_desc + uint64_t(2 * k), k_desc1 + uint64_t(2 * k), idesc_s1, k != 0);
            if (n2 > 0)
              umma_bf16_ss(tmem_base + S_COL + 256, q_desc + uint64_t(2 * k), k_desc2 + uint64_t(2 * k), idesc_s2,
                           k != 0);
          }
          umma_commit(s_full);
          if (t == p.nq - 1) umma_commit(q_empty);
          // ---- O = P V, chunk by chunk as the softmax warps produce P
          mbar_wait(o_empty, (s_iter & 1) ^ 1);
          tc_fence_after();
          for (int c = 0; c < p.nchunks; ++c, ++p_iter) {
            const int slot = p_iter % P_STAGES;
            mbar_wait(&p_full[slot], (p_iter / P_STAGES) & 1);
            tc_fence_after();
            const int keys = (p.tpad - c * P_CHUNK_KEYS) < P_CHUNK_KEYS ? (p.tpad - c * P_CHUNK_KEYS) : P_CHUNK_KEYS;
            const uint64_t p_desc = umma_smem_desc_sw128(smem_u32(smem + OFF_P + slot * P_CHUNK_BYTES), 16, 1024);
            for (int k = 0; k < keys / 16; ++k) {
              const uint64_t v_desc =
                  umma_smem_desc_sw128(sV + uint32_t(c * P_CHUNK_KEYS + k * 16) * ROW_BYTES, 1024, 1024);
              umma_bf16_ss(tmem_base + O_COL, p_desc + uint64_t(2 * k), v_desc, idesc_pv, (c | k) != 0);
            }
            umma_commit(&p_empty[slot]);
          }
          umma_commit(o_full);
        }
        umma_commit(&kv_empty[buf]);
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------------------- softmax + epilogue
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row inside the query tile
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    uint32_t s_iter = 0, p_iter = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int b = pair / p.H, h = pair - b * p.H;
      for (int t = 0; t < p.nq; ++t, ++s_iter) {
        const bool warp_active = t * QT + q * 32 < p.T;  // any valid query row in this warp
        const int tok = t * QT + r;
        mbar_wait(s_full, s_iter & 1);
        tc_fence_after();
        // ---- pass 1: row max over the valid keys
        float m = -INFINITY;
        if (warp_active) {
          for (int c0 = 0; c0 < p.tpad; c0 += 32) {
            if (c0 + 32 <= p.tpad) {
              uint32_t v[32];
              tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL + c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + j < p.T) m = fmaxf(m, __uint_as_float(v[j]));
            } else {
              uint32_t v[16];
              tmem_ld_32x32b_x16(tmem_base + lane_addr + S_COL + c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.T) m = fmaxf(m, __uint_as_float(v[j]));
            }
          }
        }
        const float msl = m * p.sl2;
        // ---- pass 2: exponentials, row sum, bf16 P chunks into the smem ring
        float l = 0.f;
        for (int c = 0; c < p.nchunks; ++c, ++p_iter) {
          const int slot = p_iter % P_STAGES;
          mbar_wait(&p_empty[slot], ((p_iter / P_STAGES) & 1) ^ 1);
          if (warp_active) {
            const int c0 = c * P_CHUNK_KEYS;
            const int keys = (p.tpad - c0) < P_CHUNK_KEYS ? (p.tpad - c0) : P_CHUNK_KEYS;
            uint8_t* prow = smem + OFF_P + slot * P_CHUNK_BYTES + r * 128;
            for (int g0 = 0; g0 < keys; g0 += 32) {
              uint32_t v[32];
              if (keys - g0 >= 32) {
                tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL + c0 + g0, v);
              } else {
                uint32_t w[16];
                tmem_ld_32x32b_x16(tmem_base + lane_addr + S_COL + c0 + g0, w);
#pragma unroll
                for (int j = 0; j < 16; ++j) { v[j] = w[j]; v[16 + j] = 0xff800000u; }
              }
              tmem_ld_wait();
              float e[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float x = ex2(fmaf(__uint_as_float(v[j]), p.sl2, -msl));
                e[j] = (c0 + g0 + j < p.T) ? x : 0.f;
                l += e[j];
              }
              const int nvec = (keys - g0 >= 32) ? 4 : 2;
#pragma unroll
              for (int jv = 0; jv < 4; ++jv)
                if (jv < nvec) {
                  const int chunk16 = (g0 >> 3) + jv;  // 16-byte chunk index inside the 128-byte row
                  uint4 o;
                  o.x = pack_bf16x2(e[jv * 8 + 0], e[jv * 8 + 1]);
                  o.y = pack_bf16x2(e[jv * 8 + 2], e[jv * 8 + 3]);
                  o.z = pack_bf16x2(e[jv * 8 + 4], e[jv * 8 + 5]);
                  o.w = pack_bf16x2(e[jv * 8 + 6], e[jv * 8 + 7]);
                  *reinterpret_cast<uint4*>(prow + ((chunk16 ^ (r & 7)) << 4)) = o;
                }
            }
            fence_proxy_async_smem();
          }
          if (c == p.nchunks - 1) {
            // all reads of S are complete -> the next tile's Q K^T may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[slot]);
        }
        // ---- epilogue: O / rowsum -> bf16 -> HBM
        mbar_wait(o_full, s_iter & 1);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        if (warp_active) {
          tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL, o0);
          tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL + 32, o1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
        if (warp_active && tok < p.T) {
          const float inv = 1.0f / l;
          uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t(b) * p.T + tok) * (p.H * HD) + h * HD);
#pragma unroll
          for (int jv = 0; jv < 4; ++jv) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(o0[jv * 8 + 0]) * inv, __uint_as_float(o0[jv * 8 + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(o0[jv * 8 + 2]) * inv, __uint_as_float(o0[jv * 8 + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(o0[jv * 8 + 4]) * inv, __uint_as_float(o0[jv * 8 + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(o0[jv * 8 + 6]) * inv, __uint_as_float(o0[jv * 8 + 7]) * inv);
            dst[jv] = o;
          }
#pragma unroll
          for (int jv = 0; jv < 4; ++jv) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(o1[jv * 8 + 0]) * inv, __uint_as_float(o1[jv * 8 + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(o1[jv * 8 + 2]) * inv, __uint_as_float(o1[jv * 8 + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(o1[jv * 8 + 4]) * inv, __uint_as_float(o1[jv * 8 + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(o1[jv * 8 + 6]) * inv, __uint_as_float(o1[jv * 8 + 7]) * inv);
            dst[4 + jv] = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

int attention_bf16(const bf16* qkv, bf16* out, int B, int T, int H, float scale, cudaStream_t stream) {
  FP_REQUIRE(B > 0 && H > 0 && T > 0, "attention: empty problem");
  const int tpad = (T + 15) / 16 * 16;
  FP_REQUIRE(tpad <= MAX_TPAD, "attention: %d tokens per image exceeds the single-pass limit of %d "
             "(crops above 224x224 need the tiled-key kernel)", T, MAX_TPAD);
  const int nq = (T + QT - 1) / QT;
  FP_REQUIRE(nq <= MAX_QTILES, "attention: too many query tiles");
  const int C = 3 * H * HD;
  CUtensorMap tmQ, tmKV;
  const uint64_t rows = uint64_t(B) * T;
  if (int rc = make_tmap_2d_bf16(&tmQ, qkv, rows, uint64_t(C), uint64_t(C), QT, HD)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmKV, qkv, rows, uint64_t(C), uint64_t(C), uint32_t(tpad / 2), HD)) return rc;
  Params p;
  p.out = out; p.B = B; p.T = T; p.H = H;
  p.tpad = tpad; p.nq = nq;
  p.nchunks = (tpad + P_CHUNK_KEYS - 1) / P_CHUNK_KEYS;
  p.sl2 = scale * 1.4426950408889634f;
  static bool attr_done = false;
  if (!attr_done) {
    FP_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done = true;
  }
  const int npairs = B * H;
  const int grid = npairs < sm_count() ? npairs : sm_count();
  ProfScope prof(PROF_ATTENTION, 4.0 * double(B) * H * double(T) * T * HD, 1, stream);
  attention_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, p);
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fp
