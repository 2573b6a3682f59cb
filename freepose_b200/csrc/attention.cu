// ViT self-attention on tcgen05.  One persistent CTA per SM walks over
// (image, head) pairs; all keys of one head (T <= 272 tokens: 261 at 224^2) sit in shared memory, so softmax is
// single pass (no online rescale).
//
//   warp 0        TMA loader   K,V (double buffered across pairs), Q tiles through a 2-slot ring, tail-query rows
//   warp 1        MMA issuer   S = Q K^T (M128 x N{256,+16} x K64, fp32 accumulators in TMEM)
//                              O = P V   (P read straight from TMEM -- "TS" form --, V as MN-major smem operand)
//   warp 2        TMEM allocator
//   warps 4..11   softmax      two warps per TMEM lane quarter split the key columns of every row: row max,
//                              p = exp2((s - max) * scale*log2e) in fp32, row sum of the unrounded p, P rounded to
//                              bf16 and stored back into TMEM with tcgen05.st (no shared-memory round trip), then
//                              O / rowsum -> bf16 -> HBM.
//   warps 12..15  tail queries token counts such as 261 = 2*128 + 5 leave a query tile with a handful of rows (cls +
//                              registers).  As a tensor-core tile they cost as much as a full one (measured: T=261
//                              took 1.83x the time of T=256), so remainders of <= 8 rows are computed on the CUDA cores
//                              by four otherwise idle warps, straight from the K/V already resident in shared memory
//                              and completely decoupled from the TMEM / mbarrier pipeline of the full tiles.
//
// Arithmetic contract = flash/xformers attention (oracle/vit.py contract_attention): logits and softmax statistics
// in fp32, un-normalised P rounded to bf16 before P.V (fp32 accumulate), one rounding of O.
#include "common.cuh"
#include "kernels.h"

namespace fp {

namespace {

constexpr int HD = 64;             // head dim
constexpr int QT = 128;            // query rows per tile
constexpr int MAX_TPAD = 272;      // padded key count (multiple of 16)
constexpr int ROW_BYTES = HD * 2;  // 128 B: one swizzle row
constexpr int Q_TILE_BYTES = QT * ROW_BYTES;    // 16 KB
constexpr int KV_BYTES = MAX_TPAD * ROW_BYTES;  // 34 KB
constexpr int TAIL_MAX = 8;        // remainder query rows handled by the CUDA-core tail warps
constexpr int TAIL_BOX = 16;       // rows per tail-Q TMA box
constexpr int NUM_SOFTMAX_WARPS = 8;
constexpr int NUM_TAIL_WARPS = 4;
constexpr int NUM_THREADS = 128 + (NUM_SOFTMAX_WARPS + NUM_TAIL_WARPS) * 32;
constexpr int TMEM_COLS = 512;
constexpr int S_COL = 0;      // 272 fp32 columns
constexpr int P_COL = 272;    // 136 columns of packed bf16x2
constexpr int O_COL = 408;    // 64 fp32 columns
constexpr int MAX_CHUNKS = 5; // 64-key chunks of P
constexpr int P_CHUNK_KEYS = 64;

constexpr int OFF_Q = 0;                           // 2 ring slots
constexpr int OFF_K = OFF_Q + 2 * Q_TILE_BYTES;    // 2 buffers
constexpr int OFF_V = OFF_K + 2 * KV_BYTES;        // 2 buffers
constexpr int OFF_XCH = OFF_V + 2 * KV_BYTES;      // float [2][128] max + [2][128] sum
constexpr int OFF_TQ = OFF_XCH + 4 * 128 * 4;      // tail Q rows: 2 slots x 16 rows x 128 B (TMA, 128B swizzle)
constexpr int OFF_TQF = OFF_TQ + 2 * TAIL_BOX * ROW_BYTES;   // float [8][64]: tail queries in fp32
constexpr int OFF_TP = OFF_TQF + TAIL_MAX * HD * 4;          // float [272][8]: bf16-rounded P of the tail rows
constexpr int OFF_TO = OFF_TP + MAX_TPAD * TAIL_MAX * 4;     // float [4 warps][8][64]: partial P.V
constexpr int OFF_TRED = OFF_TO + NUM_TAIL_WARPS * TAIL_MAX * HD * 4;  // float [2][4][8]: max / sum partials
constexpr int OFF_BAR = OFF_TRED + 2 * NUM_TAIL_WARPS * TAIL_MAX * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;

struct Params {
  bf16* out;
  int B, T, H;
  int tpad;       // keys padded to a multiple of 16
  int n_normal;   // query tiles handled row-per-thread
  int n_tail;     // remainder query rows (<= 8) handled by the tail warps, 0 if none
  int nchunks;    // 64-key chunks
  float sl2;      // scale * log2(e)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// exponentials of 32 logits -> 16 packed bf16x2 words; returns the fp32 sum of the unrounded values.
// MASKED: columns >= valid are forced to 0 (only the group that straddles T takes this path).
template <bool MASKED>
__device__ __forceinline__ float exp_group(const uint32_t (&v)[32], float sl2, float msl, int valid,
                                           uint32_t (&packed)[16]) {
  float part[4] = {0.f, 0.f, 0.f, 0.f};  // four independent chains instead of one 32-long dependency
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float e0 = ex2(fmaf(__uint_as_float(v[j]), sl2, -msl));
    float e1 = ex2(fmaf(__uint_as_float(v[j + 1]), sl2, -msl));
    if (MASKED) {
      e0 = j < valid ? e0 : 0.f;
      e1 = j + 1 < valid ? e1 : 0.f;
    }
    part[(j >> 1) & 3] += e0 + e1;
    packed[j >> 1] = pack_bf16x2(e0, e1);
  }
  return (part[0] + part[1]) + (part[2] + part[3]);
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

__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQt,
                 const __grid_constant__ CUtensorMap tmKV, const bf16* __restrict__ qkv_unused, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_full = bars;        // [2]
  uint64_t* kv_empty = bars + 2;   // [2]  MMA commit (+ the four tail warps when there are tail rows)
  uint64_t* q_full = bars + 4;     // [2]
  uint64_t* q_empty = bars + 6;    // [2]
  uint64_t* s_full = bars + 8;
  uint64_t* s_empty = bars + 9;
  uint64_t* o_full = bars + 10;
  uint64_t* o_empty = bars + 11;
  uint64_t* p_full = bars + 12;    // [MAX_CHUNKS]
  uint64_t* tq_full = bars + 17;   // [2]
  uint64_t* tq_empty = bars + 19;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 21);
  float* xch_max = reinterpret_cast<float*>(smem + OFF_XCH);        // [2][128]
  float* xch_sum = xch_max + 2 * 128;                                // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = p.B * p.H;
  const int half_rows = p.tpad / 2;
  const int tiles_per_pair = p.n_normal;
  (void)qkv_unused;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmQt);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], (tiles_per_pair > 0 ? 1 : 0) + (p.n_tail > 0 ? NUM_TAIL_WARPS : 0));
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
      mbar_init(&tq_full[i], 1); mbar_init(&tq_empty[i], NUM_TAIL_WARPS);
    }
    mbar_init(s_full, 1); mbar_init(s_empty, NUM_SOFTMAX_WARPS);
    mbar_init(o_full, 1); mbar_init(o_empty, NUM_SOFTMAX_WARPS);
    for (int i = 0; i < MAX_CHUNKS; ++i) mbar_init(&p_full[i], NUM_SOFTMAX_WARPS);
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
      uint32_t qi = 0;
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
        if (p.n_tail > 0) {
          mbar_wait(&tq_empty[buf], ((it >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&tq_full[buf], TAIL_BOX * ROW_BYTES);
          tma_load_2d(smem + OFF_TQ + buf * TAIL_BOX * ROW_BYTES, &tmQt, &tq_full[buf], h * HD, row0 + p.T - p.n_tail);
        }
        for (int t = 0; t < tiles_per_pair; ++t, ++qi) {
          const int slot = qi & 1;
          mbar_wait(&q_empty[slot], ((qi >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[slot], Q_TILE_BYTES);
          tma_load_2d(smem + OFF_Q + slot * Q_TILE_BYTES, &tmQ, &q_full[slot], h * HD, row0 + t * QT);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------- MMA issuer
    if (lane == 0 && tiles_per_pair > 0) {
      const int n1 = p.tpad > 256 ? 256 : p.tpad;
      const int n2 = p.tpad - n1;
      const uint32_t idesc_s1 = umma_idesc_bf16(QT, n1, 0, 0);
      const uint32_t idesc_s2 = umma_idesc_bf16(QT, n2 > 0 ? n2 : 16, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(QT, HD, 0, 1);  // B (= V) is MN-major
      uint32_t tile_iter = 0;
      int it = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t sK = smem_u32(smem + OFF_K + buf * KV_BYTES);
        const uint32_t sV = smem_u32(smem + OFF_V + buf * KV_BYTES);
        mbar_wait(&kv_full[buf], (it >> 1) & 1);
        for (int t = 0; t < tiles_per_pair; ++t, ++tile_iter) {
          const int slot = tile_iter & 1;
          const uint32_t sQ = smem_u32(smem + OFF_Q + slot * Q_TILE_BYTES);
          mbar_wait(&q_full[slot], (tile_iter >> 1) & 1);
          mbar_wait(s_empty, (tile_iter & 1) ^ 1);
          tc_fence_after();
          // ---- S = Q_t K^T
          const uint64_t q_desc = umma_smem_desc_sw128(sQ, 16, 1024);
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
          umma_commit(&q_empty[slot]);
          // ---- O = P V, P read from TMEM as the softmax warps store it, chunk by chunk
          mbar_wait(o_empty, (tile_iter & 1) ^ 1);
          tc_fence_after();
          for (int c = 0; c < p.nchunks; ++c) {
            mbar_wait(&p_full[c], tile_iter & 1);
            tc_fence_after();
            const int keys = (p.tpad - c * 64) < 64 ? (p.tpad - c * 64) : 64;
            for (int k = 0; k < keys / 16; ++k) {
              const int key0 = c * 64 + k * 16;
              const uint64_t v_desc = umma_smem_desc_sw128(sV + uint32_t(key0) * ROW_BYTES, 1024, 1024);
              umma_bf16_ts(tmem_base + O_COL, tmem_base + P_COL + uint32_t(key0 >> 1), v_desc, idesc_pv, key0 != 0);
            }
          }
          umma_commit(o_full);
        }
        umma_commit(&kv_empty[buf]);
      }
    }
  } else if (warp >= 4 && warp < 4 + NUM_SOFTMAX_WARPS) {
    // ---------------------------------------------------------------------------- softmax + epilogue
    const int q = warp & 3;              // TMEM lane quarter
    const int hf = (warp - 4) >> 2;      // which half of the key columns of a row this warp handles
    const int r = q * 32 + lane;         // row inside the tile
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const int ngroups = (p.tpad + 31) / 32;  // 32-column groups of S (the last one may hold 16 columns)
    uint32_t tile_iter = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int b = pair / p.H, h = pair - b * p.H;
      for (int t = 0; t < tiles_per_pair; ++t, ++tile_iter) {
        const uint32_t par = tile_iter & 1;
        mbar_wait(s_full, par);
        tc_fence_after();
        const bool warp_active = t * QT + q * 32 < p.T - p.n_tail;  // any query row of this warp in the tile
        const int tok = t * QT + r;
        // ---- pass 1: row max (this warp: groups with g % 2 == hf)
        float m = -INFINITY;
        if (warp_active) {
          for (int g = hf; g < ngroups; g += 2) {
            const int c0 = g * 32;
            uint32_t v[32];
            if (p.tpad - c0 >= 32) {
              tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL + c0, v);
            } else {
              uint32_t w16[16];
              tmem_ld_32x32b_x16(tmem_base + lane_addr + S_COL + c0, w16);
#pragma unroll
              for (int j = 0; j < 16; ++j) { v[j] = w16[j]; v[16 + j] = 0xff800000u; }
            }
            tmem_ld_wait();
            if (c0 + 32 <= p.T) m = max_group<false>(v, 32, m);
            else                m = max_group<true>(v, p.T - c0, m);
          }
        }
        xch_max[hf * 128 + r] = m;
        named_bar_sync(1 + q, 64);
        m = fmaxf(xch_max[r], xch_max[128 + r]);
        const float msl = m * p.sl2;
        // ---- pass 2: exponentials, row sum, bf16 P into TMEM (chunk c: this warp owns keys [64c + 32hf, +32))
        float l = 0.f;
        for (int c = 0; c < p.nchunks; ++c) {
          const int c0 = c * 64 + hf * 32;
          if (warp_active && c0 < p.tpad) {
            const int width = p.tpad - c0 >= 32 ? 32 : 16;
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
            if (c0 + 32 <= p.T) l += exp_group<false>(v, p.sl2, msl, 32, pk);
            else                l += exp_group<true>(v, p.sl2, msl, p.T - c0, pk);
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
          if (c == p.nchunks - 1 && lane == 0) mbar_arrive(s_empty);  // all S reads of this warp are done
          if (lane == 0) mbar_arrive(&p_full[c]);
        }
        xch_sum[hf * 128 + r] = l;
        named_bar_sync(1 + q, 64);
        l = xch_sum[r] + xch_sum[128 + r];
        // ---- epilogue: this warp normalises 32 of the 64 output columns
        mbar_wait(o_full, par);
        tc_fence_after();
        uint32_t o[32];
        if (warp_active) {
          tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL + hf * 32, o);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
        if (warp_active && tok < p.T - p.n_tail) {
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
    }
  } else if (warp >= 4 + NUM_SOFTMAX_WARPS && p.n_tail > 0) {
    // ---------------------------------------------------------------------------- tail queries on the CUDA cores
    // 128 threads.  Phase 1: thread <-> key (keys tt, tt+128, tt+256): logits of the <= 8 tail queries against its
    // keys, K rows read from the TMA-swizzled smem tile (row-per-thread reads are conflict free under the 128B
    // swizzle).  Phase 2: softmax statistics across keys (shuffles + smem).  Phase 3: lane <-> two head dims, warp
    // <-> every 4th key: O += P[key] * V[key]; partials combined through smem.
    const int tw = warp - 4 - NUM_SOFTMAX_WARPS;   // 0..3
    const int tt = tw * 32 + lane;                 // 0..127
    const int nt = p.n_tail;
    float* tqf = reinterpret_cast<float*>(smem + OFF_TQF);     // [8][64]
    float* tp = reinterpret_cast<float*>(smem + OFF_TP);       // [272][8]
    float* to = reinterpret_cast<float*>(smem + OFF_TO);       // [4][8][64]
    float* tredm = reinterpret_cast<float*>(smem + OFF_TRED);  // [4][8]
    float* treds = tredm + NUM_TAIL_WARPS * TAIL_MAX;          // [4][8]
    int it = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
      const int b = pair / p.H, h = pair - b * p.H;
      const int buf = it & 1;
      const uint8_t* sK = smem + OFF_K + buf * KV_BYTES;
      const uint8_t* sV = smem + OFF_V + buf * KV_BYTES;
      const uint8_t* sTQ = smem + OFF_TQ + buf * TAIL_BOX * ROW_BYTES;
      mbar_wait(&kv_full[buf], (it >> 1) & 1);
      mbar_wait(&tq_full[buf], (it >> 1) & 1);
      // tail queries -> fp32 (so the inner loop needs no unpacking on the query side)
      for (int i = tt; i < nt * HD / 2; i += NUM_TAIL_WARPS * 32) {
        const int j = i / (HD / 2), dp = i - j * (HD / 2);  // query row, pair of head dims
        const uint32_t u = *reinterpret_cast<const uint32_t*>(sTQ + j * ROW_BYTES + (((dp >> 2) ^ (j & 7)) << 4) + (dp & 3) * 4);
        tqf[j * HD + 2 * dp] = bf16lo(u);
        tqf[j * HD + 2 * dp + 1] = bf16hi(u);
      }
      named_bar_sync(6, NUM_TAIL_WARPS * 32);
      // ---- phase 1: logits
      float s[3][TAIL_MAX];
#pragma unroll
      for (int kk = 0; kk < 3; ++kk)
#pragma unroll
        for (int j = 0; j < TAIL_MAX; ++j) s[kk][j] = 0.f;
      const int nkk = (tt + 256 < p.T) ? 3 : ((tt + 128 < p.T) ? 2 : (tt < p.T ? 1 : 0));
      for (int c = 0; c < 8; ++c) {  // 16-byte chunks of a 128-byte row = 8 head dims
        float kf[3][8];
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const int key = tt + 128 * kk;
          if (kk < nkk) {
            const uint4 u = *reinterpret_cast<const uint4*>(sK + key * ROW_BYTES + ((c ^ (key & 7)) << 4));
            kf[kk][0] = bf16lo(u.x); kf[kk][1] = bf16hi(u.x); kf[kk][2] = bf16lo(u.y); kf[kk][3] = bf16hi(u.y);
            kf[kk][4] = bf16lo(u.z); kf[kk][5] = bf16hi(u.z); kf[kk][6] = bf16lo(u.w); kf[kk][7] = bf16hi(u.w);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) kf[kk][e] = 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < TAIL_MAX; ++j)
          if (j < nt) {
            const float4 qa = *reinterpret_cast<const float4*>(tqf + j * HD + c * 8);
            const float4 qb = *reinterpret_cast<const float4*>(tqf + j * HD + c * 8 + 4);
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
              float a = s[kk][j];
              a = fmaf(qa.x, kf[kk][0], a); a = fmaf(qa.y, kf[kk][1], a);
              a = fmaf(qa.z, kf[kk][2], a); a = fmaf(qa.w, kf[kk][3], a);
              a = fmaf(qb.x, kf[kk][4], a); a = fmaf(qb.y, kf[kk][5], a);
              a = fmaf(qb.z, kf[kk][6], a); a = fmaf(qb.w, kf[kk][7], a);
              s[kk][j] = a;
            }
          }
      }
      // ---- phase 2: softmax statistics over all keys
#pragma unroll
      for (int j = 0; j < TAIL_MAX; ++j)
        if (j < nt) {
          float m = -INFINITY;
#pragma unroll
          for (int kk = 0; kk < 3; ++kk)
            if (kk < nkk) m = fmaxf(m, s[kk][j]);
          m = warp_max(m);
          if (lane == 0) tredm[tw * TAIL_MAX + j] = m;
        }
      named_bar_sync(6, NUM_TAIL_WARPS * 32);
      float lsum[TAIL_MAX];
#pragma unroll
      for (int j = 0; j < TAIL_MAX; ++j) {
        lsum[j] = 0.f;
        if (j < nt) {
          const float m = fmaxf(fmaxf(tredm[j], tredm[TAIL_MAX + j]), fmaxf(tredm[2 * TAIL_MAX + j], tredm[3 * TAIL_MAX + j]));
          const float msl = m * p.sl2;
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const int key = tt + 128 * kk;
            const float e = kk < nkk ? ex2(fmaf(s[kk][j], p.sl2, -msl)) : 0.f;
            lsum[j] += e;
            if (key < p.tpad) tp[key * TAIL_MAX + j] = bf16_round(e);  // P is rounded to bf16 before P.V
          }
          const float ws = warp_sum(lsum[j]);
          if (lane == 0) treds[tw * TAIL_MAX + j] = ws;
        }
      }
      named_bar_sync(6, NUM_TAIL_WARPS * 32);
      // ---- phase 3: O = P V   (lane <-> head dims 2*lane, 2*lane+1; warp <-> keys tw, tw+4, ...)
      float ox[TAIL_MAX], oy[TAIL_MAX];
#pragma unroll
      for (int j = 0; j < TAIL_MAX; ++j) { ox[j] = 0.f; oy[j] = 0.f; }
      for (int key = tw; key < p.T; key += NUM_TAIL_WARPS) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(sV + key * ROW_BYTES + (((lane >> 2) ^ (key & 7)) << 4) + (lane & 3) * 4);
        const float vx = bf16lo(u), vy = bf16hi(u);
        const float4 pa = *reinterpret_cast<const float4*>(tp + key * TAIL_MAX);
        const float4 pb = *reinterpret_cast<const float4*>(tp + key * TAIL_MAX + 4);
        const float pj[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int j = 0; j < TAIL_MAX; ++j)
          if (j < nt) { ox[j] = fmaf(pj[j], vx, ox[j]); oy[j] = fmaf(pj[j], vy, oy[j]); }
      }
      // K, V and the tail Q rows of this pair are no longer needed by these warps
      __syncwarp();
      if (lane == 0) { mbar_arrive(&kv_empty[buf]); mbar_arrive(&tq_empty[buf]); }
#pragma unroll
      for (int j = 0; j < TAIL_MAX; ++j)
        if (j < nt) *reinterpret_cast<float2*>(to + (tw * TAIL_MAX + j) * HD + 2 * lane) = make_float2(ox[j], oy[j]);
      named_bar_sync(6, NUM_TAIL_WARPS * 32);
      for (int i = tt; i < nt * HD / 2; i += NUM_TAIL_WARPS * 32) {
        const int j = i / (HD / 2), dp = i - j * (HD / 2);
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < NUM_TAIL_WARPS; ++w) {
          const float2 x = *reinterpret_cast<const float2*>(to + (w * TAIL_MAX + j) * HD + 2 * dp);
          acc.x += x.x; acc.y += x.y;
        }
        const float l = (treds[j] + treds[TAIL_MAX + j]) + (treds[2 * TAIL_MAX + j] + treds[3 * TAIL_MAX + j]);
        const float inv = 1.0f / l;
        const int tok = p.T - nt + j;
        *reinterpret_cast<uint32_t*>(p.out + (size_t(b) * p.T + tok) * (p.H * HD) + h * HD + 2 * dp) =
            pack_bf16x2(acc.x * inv, acc.y * inv);
      }
      named_bar_sync(6, NUM_TAIL_WARPS * 32);  // smem scratch is reused by the next pair
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
  const int rem = T % QT;
  const int n_tail = (rem > 0 && rem <= TAIL_MAX) ? rem : 0;
  const int n_normal = T / QT + ((rem > TAIL_MAX) ? 1 : 0);
  const int C = 3 * H * HD;
  CUtensorMap tmQ, tmQt, tmKV;
  const uint64_t rows = uint64_t(B) * T;
  if (int rc = make_tmap_2d_bf16(&tmQ, qkv, rows, uint64_t(C), uint64_t(C), QT, HD)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmQt, qkv, rows, uint64_t(C), uint64_t(C), TAIL_BOX, HD)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmKV, qkv, rows, uint64_t(C), uint64_t(C), uint32_t(tpad / 2), HD)) return rc;
  Params p;
  p.out = out; p.B = B; p.T = T; p.H = H;
  p.tpad = tpad; p.n_normal = n_normal; p.n_tail = n_tail;
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
  attention_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmQt, tmKV, qkv, p);
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fp
