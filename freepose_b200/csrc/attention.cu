// ViT self-attention on tcgen05.  One persistent CTA per SM walks over
// (image, head) pairs; all keys of one head (T <= 272 tokens: 261 at 224^2) sit in shared memory, so softmax is
// single pass (no online rescale).
//
//   warp 0        TMA loader   K,V (double buffered across pairs), Q tiles through a 2-slot ring, tail-query rows
//   warp 1        MMA issuer   S = Q K^T (M128 x N{128, tpad-128} x K64, fp32 accumulators in TMEM), issued in two key
//                              halves so that the next tile's logits are produced while the softmax warps are still
//                              exponentiating the current tile (the low half as soon as its columns have been consumed)
//                              O = P V   (P read straight from TMEM -- "TS" form --, V as MN-major smem operand);
//                              O is double buffered so a tile's normalisation/store is deferred into the next tile
//   warps 4..11   softmax      two warps per TMEM lane quarter split the key columns of every row: row max,
//                              p = exp2((s - max) * scale*log2e) in fp32, row sum of the unrounded p, P rounded to
//                              bf16 and stored with tcgen05.st over the logit columns it was computed from (no extra
//                              TMEM, no shared-memory round trip); O / rowsum -> bf16 -> HBM of the previous tile
//                              runs between two exponential chunks of the current one.
//   warps 2..3    tail queries token counts such as 261 = 2*128 + 5 leave a query tile with a handful of rows (cls +
//                              registers).  As a tensor-core tile they cost as much as a full one (measured: T=261
//                              took 1.83x the time of T=256), so remainders of <= 8 rows are computed by two otherwise
//                              idle warps with mma.sync straight from the K/V already resident in shared memory,
//                              completely decoupled from the TMEM / mbarrier pipeline of the full tiles (warp 2 also
//                              allocates the TMEM).
//
// Arithmetic contract = flash/xformers attention (oracle/vit.py contract_attention): logits and softmax statistics
// in fp32, un-normalised P rounded to bf16 before P.V (fp32 accumulate), one rounding of O.
#include <stdio.h>
#include <stdlib.h>

#include "attention_common.cuh"
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
constexpr int NUM_TAIL_WARPS = 2;     // warps 2 and 3
constexpr int NUM_THREADS = 128 + NUM_SOFTMAX_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int S_COL = 0;      // 272 fp32 logit columns; bf16x2 P overwrites the first half of every consumed 32-column group
constexpr int O_COL = 272;    // 2 x 64 fp32 columns (double buffered across tiles)
constexpr int S_PART = 128;   // S is issued as keys [0,128), [128,256), [256,tpad): a part of the next tile follows the last P.V that reads it
constexpr int MAX_CHUNKS = 5; // 64-key chunks of P
constexpr int P_CHUNK_KEYS = 64;

constexpr int OFF_Q = 0;                           // 2 ring slots
constexpr int OFF_K = OFF_Q + 2 * Q_TILE_BYTES;    // 2 buffers
constexpr int OFF_V = OFF_K + 2 * KV_BYTES;        // 2 buffers
constexpr int OFF_XCH = OFF_V + 2 * KV_BYTES;      // float [2 tile parities][2 halves][128] max + the same for sum
constexpr int OFF_TQ = OFF_XCH + 8 * 128 * 4;      // tail Q rows: 2 slots x 16 rows x 128 B (TMA, 128B swizzle)
constexpr int OFF_OST = OFF_TQ + 2 * TAIL_BOX * ROW_BYTES;   // output staging: 8 softmax warps x (32 rows x 64 B), 64B swizzle
constexpr int OFF_TO = OFF_OST + NUM_SOFTMAX_WARPS * 2048;   // float [4 warps][8][64]: partial P.V of the tail rows
static_assert(OFF_OST % 1024 == 0, "TMA store staging must keep the swizzle alignment");
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
  long long* dbg;  // perf experiments: per-phase cycle counters of softmax warp 4 of CTA 0 (nullptr = off)
};

using namespace attn;

// T_CONST: token count known at compile time (261 = the 224^2 crops of the headline path: every chunk width, mask and
// trip count folds to a constant and the straddling-group code is emitted once); 0 = any T <= 272 at run time.
template <bool TIMING, int T_CONST, unsigned POLY>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQt,
                 const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmOut, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the __shared__ array keeps the shared address space: LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_full = bars;        // [2]
  uint64_t* kv_empty = bars + 2;   // [2]  MMA commit (+ the four tail warps when there are tail rows)
  uint64_t* q_full = bars + 4;     // [2]
  uint64_t* q_empty = bars + 6;    // [2]
  uint64_t* s_full = bars + 8;      // [3]  key columns [0,128) / [128,256) / [256,tpad) of S
  uint64_t* o_full = bars + 11;    // [2]
  uint64_t* o_empty = bars + 13;   // [2]
  uint64_t* p_full = bars + 15;    // [MAX_CHUNKS]; all eight arrivals on chunk c also mean "S columns of chunk c are consumed"
  uint64_t* tq_full = bars + 20;   // [2]
  uint64_t* tq_empty = bars + 22;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 24);
  float* xch_max = reinterpret_cast<float*>(smem + OFF_XCH);        // [2 parities][2 halves][128]
  float* xch_sum = xch_max + 4 * 128;                                // [2 parities][2 halves][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = p.B * p.H;
  constexpr int kRem = T_CONST % QT;
  const int T = T_CONST ? T_CONST : p.T;
  const int tpad = T_CONST ? (T_CONST + 15) / 16 * 16 : p.tpad;
  const int n_tail = T_CONST ? ((kRem > 0 && kRem <= TAIL_MAX) ? kRem : 0) : p.n_tail;
  const int tiles_per_pair = T_CONST ? (T_CONST / QT + (kRem > TAIL_MAX ? 1 : 0)) : p.n_normal;
  const int nchunks = T_CONST ? ((T_CONST + 15) / 16 * 16 + P_CHUNK_KEYS - 1) / P_CHUNK_KEYS : p.nchunks;
  constexpr bool kAllRowsLive = T_CONST != 0 && (kRem == 0 || kRem <= TAIL_MAX);  // no partially filled full tile
  const int half_rows = tpad / 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmQt);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], (tiles_per_pair > 0 ? 1 : 0) + (n_tail > 0 ? NUM_TAIL_WARPS : 0));
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
      mbar_init(&tq_full[i], 1); mbar_init(&tq_empty[i], NUM_TAIL_WARPS);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&s_full[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], NUM_SOFTMAX_WARPS); }
    for (int i = 0; i < MAX_CHUNKS; ++i) mbar_init(&p_full[i], NUM_SOFTMAX_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // The specialised instance keeps three of a row share's four 32-logit groups in registers between the max and the
  // exponential pass (TMEM reads run at 64 B/clk per scheduler: reading S twice was the floor of this kernel).  With
  // 12 warps every thread may use 168 registers, which is enough without warpgroup register reallocation.
  constexpr bool kSingleRead = T_CONST == 261;

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
        mbar_arrive_expect_tx(&kv_full[buf], 2 * tpad * ROW_BYTES);
        uint8_t* sK = smem + OFF_K + buf * KV_BYTES;
        uint8_t* sV = smem + OFF_V + buf * KV_BYTES;
        const int kcol = p.H * HD + h * HD, vcol = 2 * p.H * HD + h * HD;
        tma_load_2d(sK, &tmKV, &kv_full[buf], kcol, row0);
        tma_load_2d(sK + half_rows * ROW_BYTES, &tmKV, &kv_full[buf], kcol, row0 + half_rows);
        tma_load_2d(sV, &tmKV, &kv_full[buf], vcol, row0);
        tma_load_2d(sV + half_rows * ROW_BYTES, &tmKV, &kv_full[buf], vcol, row0 + half_rows);
        if (n_tail > 0) {
          mbar_wait(&tq_empty[buf], ((it >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&tq_full[buf], TAIL_BOX * ROW_BYTES);
          tma_load_2d(smem + OFF_TQ + buf * TAIL_BOX * ROW_BYTES, &tmQt, &tq_full[buf], h * HD, row0 + T - n_tail);
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
    // One thread feeds the tensor pipe, so its scalar instruction stream is on the critical path (measured: with
    // descriptors rebuilt and run-time loop bounds per MMA the issue rate was ~150 cycles per tcgen05.mma and the pipe
    // starved).  Everything below is therefore base + compile-time offset: descriptors are built once, chunk / k-step
    // loops are fully unrolled, tile coordinates advance incrementally (no divisions).
    if (tiles_per_pair > 0 && elect_one()) {
      const uint32_t idesc_pv = umma_idesc_bf16(QT, HD, 0, 1);  // B (= V) is MN-major
      int part_n[3], part_last_chunk[3];   // key count of each part, last 64-key P chunk that lives in its columns
      uint32_t idesc_s[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int end = (i + 1) * S_PART < tpad ? (i + 1) * S_PART : tpad;
        part_n[i] = end - i * S_PART > 0 ? end - i * S_PART : 0;
        part_last_chunk[i] = (end + P_CHUNK_KEYS - 1) / P_CHUNK_KEYS - 1;
        idesc_s[i] = umma_idesc_bf16(QT, part_n[i] > 0 ? part_n[i] : 16, 0, 0);
      }
      const int my_pairs = (npairs - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
      const uint32_t ntiles = uint32_t(my_pairs) * uint32_t(tiles_per_pair);
      // descriptor of (slot / buffer 0, offset 0); a byte offset adds (offset >> 4) to the start-address field
      const uint64_t q_desc0 = umma_smem_desc_sw128(smem_u32(smem + OFF_Q), 16, 1024);
      const uint64_t k_desc0 = umma_smem_desc_sw128(smem_u32(smem + OFF_K), 16, 1024);
      const uint64_t v_desc0 = umma_smem_desc_sw128(smem_u32(smem + OFF_V), 1024, 1024);
      // S = Q_g K^T for one key part of tile g (pair iteration `it`, tile `t`) of this CTA.  Tensor-core work executes
      // in issue order, so a part may be issued as soon as the P.V steps reading the P in its columns have been issued.
      auto issue_s = [&](uint32_t g, int it, int t, int part) {
        const int buf = it & 1, slot = g & 1;
        if (part == 0) {
          if (t == 0) mbar_wait(&kv_full[buf], (it >> 1) & 1);
          mbar_wait(&q_full[slot], (g >> 1) & 1);
          tc_fence_after();
        }
        const uint64_t q_desc = q_desc0 + uint64_t(slot * (Q_TILE_BYTES >> 4));
        const uint64_t k_desc = k_desc0 + uint64_t(buf * (KV_BYTES >> 4) + part * ((S_PART * ROW_BYTES) >> 4));
        const uint32_t d = tmem_base + S_COL + part * S_PART;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(d, q_desc + uint64_t(2 * k), k_desc + uint64_t(2 * k), idesc_s[part], k != 0);
        umma_commit(&s_full[part]);
        // empty parts (short token counts) complete together with the last real one; so does the Q slot
#pragma unroll
        for (int nx = part + 1; nx <= 3; ++nx) {
          if (nx < 3 && part_n[nx] > 0) break;
          if (nx < 3) umma_commit(&s_full[nx]);
          else        umma_commit(&q_empty[slot]);
        }
      };
      if (ntiles > 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (part_n[i] > 0) issue_s(0, 0, 0, i);
      }
      int it = 0, t = 0;
      for (uint32_t g = 0; g < ntiles; ++g) {
        int nit = it, nt = t + 1;            // coordinates of tile g + 1
        if (nt == tiles_per_pair) { nt = 0; ++nit; }
        const bool has_next = g + 1 < ntiles;
        const int buf = it & 1;
        const uint32_t ob = g & 1;
        const uint64_t v_desc = v_desc0 + uint64_t(buf * (KV_BYTES >> 4));
        const uint32_t d_o = tmem_base + O_COL + ob * HD;
        // ---- O[ob] = P V, P read from TMEM as the softmax warps store it, chunk by chunk
        mbar_wait(&o_empty[ob], ((g >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < MAX_CHUNKS; ++c) {
          if (c < nchunks) {
            mbar_wait(&p_full[c], g & 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int key0 = c * 64 + k * 16;
              if (key0 < tpad) {
                // P of keys [key0, key0+16): 8 packed columns at the start of the 32-column logit group they came from
                const uint32_t pcol = uint32_t((key0 & ~31) + ((key0 & 16) >> 1));
                umma_bf16_ts(d_o, tmem_base + S_COL + pcol, v_desc + uint64_t(key0 * (ROW_BYTES >> 4)), idesc_pv, key0 != 0);
              }
            }
            if (c == nchunks - 1) {
              umma_commit(&o_full[ob]);
              if (t == tiles_per_pair - 1) umma_commit(&kv_empty[buf]);
            }
            if (has_next) {   // logits of the next tile for the key parts whose columns are now free
#pragma unroll
              for (int i = 0; i < 3; ++i)
                if (part_n[i] > 0 && c == part_last_chunk[i]) issue_s(g + 1, nit, nt, i);
            }
          }
        }
        it = nit; t = nt;
      }
    }
  } else if (warp >= 4 && warp < 4 + NUM_SOFTMAX_WARPS) {
    // ---------------------------------------------------------------------------- softmax + epilogue
    const int q = warp & 3;              // TMEM lane quarter
    const int hf = (warp - 4) >> 2;      // which half of the key columns of a row this warp handles
    const int r = q * 32 + lane;         // row inside the tile
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    uint8_t* out_stage = smem + OFF_OST + (warp - 4) * 2048;  // 32 rows x 64 B, source of this warp's bulk stores
    const bool timing = TIMING && p.dbg != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0;
    long long tepi = 0;

    // Normalise and store tile gp (image b, head h, query tile t) from O[gp & 1].  Runs one tile late, between two
    // exponential chunks of tile gp + 1 (or after the loop for the last tile): by then P.V of tile gp has long finished,
    // and its latencies (TMEM load, staging, bulk store) hide under the other warp's exponentials.
    auto epilogue = [&](uint32_t gp, int b, int h, int t) {
      long long te0 = 0;
      if (timing) te0 = clock64();
      const uint32_t ob = gp & 1;
      const float l = xch_sum[ob * 256 + r] + xch_sum[ob * 256 + 128 + r];
      const bool warp_active = kAllRowsLive || t * QT + q * 32 < T - n_tail;        // any query row of this warp in the tile
      const bool full_rows = kAllRowsLive || t * QT + q * 32 + 32 <= T - n_tail;    // all 32 rows of this warp are queries
      const int tok = t * QT + r;
      if (full_rows) {
        tma_store_wait_read();   // the previous tile's bulk store has finished reading the staging tile
        __syncwarp();
      }
      mbar_wait(&o_full[ob], (gp >> 1) & 1);
      tc_fence_after();
      uint32_t o[32];
      if (warp_active) {
        tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL + ob * HD + hf * 32, o);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[ob]);
      if (warp_active && (full_rows || tok < T - n_tail)) {
        const float inv = 1.0f / l;
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
            *reinterpret_cast<uint4*>(out_stage + lane * 64 + ((jv ^ ((lane >> 1) & 3)) << 4)) = w[jv];
        } else {
          uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t(b) * T + tok) * (p.H * HD) + h * HD + hf * 32);
#pragma unroll
          for (int jv = 0; jv < 4; ++jv) dst[jv] = w[jv];
        }
      }
      if (full_rows) {
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_2d(&tmOut, out_stage, h * HD + hf * 32, b * T + t * QT + q * 32);  // 32 rows x 32 columns
          tma_store_commit();
        }
      }
      if (timing) tepi += clock64() - te0;
    };

    uint32_t g = 0;                      // tile counter of this CTA
    int pb = 0, ph = 0, pt = 0;          // coordinates of tile g - 1 (its epilogue is still owed)
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int b = pair / p.H, h = pair - b * p.H;
      for (int t = 0; t < tiles_per_pair; ++t, ++g) {
        const uint32_t par = g & 1;
        long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0, tk4 = 0;
        float l = 0.f;
        if constexpr (kSingleRead) {
          // ---- 261 tokens: this warp's logits are columns [64c + 32hf, +32) for c = 0..3 plus, for hf = 0, columns
          // 256..263 (keys 256..260 are real).  One TMEM read into registers, max, exponentials from registers.
          if (timing) { tk0 = clock64(); tepi = 0; }
          // Registers hold three of the four 32-column groups between the passes; the fourth (and the 8-column tail) is
          // reduced first and read again while the others are exponentiated (176 instead of 272 columns read per row).
          uint32_t s0[32], s1[32], s2[32], s3[32], s4[8];
          const uint32_t sbase = tmem_base + lane_addr + S_COL + hf * 32;
          mbar_wait(&s_full[0], par);
          tc_fence_after();
          if (timing) tk1 = clock64();
          tmem_ld_32x32b_x32(sbase, s0);
          tmem_ld_32x32b_x32(sbase + 64, s1);
          mbar_wait(&s_full[1], par);
          tc_fence_after();
          tmem_ld_32x32b_x32(sbase + 128, s2);
          tmem_ld_32x32b_x32(sbase + 192, s3);
          mbar_wait(&s_full[2], par);
          tc_fence_after();
          if (hf == 0) tmem_ld_32x32b_x8(tmem_base + lane_addr + S_COL + 256, s4);
          tmem_ld_wait();
          float m = max_group<false>(s3, 32, -INFINITY);     // group 3 and the tail die here ...
          if (hf == 0) {
#pragma unroll
            for (int j = 0; j < T_CONST - 256; ++j) m = fmaxf(m, __uint_as_float(s4[j]));
          }
          m = max_group<false>(s0, 32, m);
          m = max_group<false>(s1, 32, m);
          m = max_group<false>(s2, 32, m);
          if (timing) tk2 = clock64();
          xch_max[par * 256 + hf * 128 + r] = m;
          named_bar_sync(1 + q, 64);
          m = fmaxf(xch_max[par * 256 + r], xch_max[par * 256 + 128 + r]);
          const float msl = m * p.sl2;
          if (timing) tk3 = clock64();
          // P of a chunk is published one chunk late: its tcgen05.st completes under the next chunk's exponentials
          auto publish = [&](int c) {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[c]);
          };
          uint32_t pk[16];
          l += exp_group<false, POLY>(s0, p.sl2, msl, 32, pk);
          tmem_st_32x32b_x16(sbase, pk);
          tmem_ld_32x32b_x32(sbase + 192, s3);                                      // ... and are read a second time here
          if (hf == 0) tmem_ld_32x32b_x8(tmem_base + lane_addr + S_COL + 256, s4);
          l += exp_group<false, POLY>(s1, p.sl2, msl, 32, pk);
          publish(0);
          tmem_st_32x32b_x16(sbase + 64, pk);
          if (g > 0 && hf == 0) epilogue(g - 1, pb, ph, pt);
          l += exp_group<false, POLY>(s2, p.sl2, msl, 32, pk);
          publish(1);
          tmem_st_32x32b_x16(sbase + 128, pk);
          tmem_ld_wait();
          l += exp_group<false, POLY>(s3, p.sl2, msl, 32, pk);
          publish(2);
          tmem_st_32x32b_x16(sbase + 192, pk);
          if (g > 0 && hf == 1) epilogue(g - 1, pb, ph, pt);
          if (hf == 0) {
            uint32_t p4[8];
            float e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              e[j] = j < T_CONST - 256 ? ex2(fmaf(__uint_as_float(s4[j]), p.sl2, -msl)) : 0.f;
              l += e[j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { p4[j] = pack_bf16x2(e[2 * j], e[2 * j + 1]); p4[4 + j] = 0u; }   // keys 264..271: P = 0
            publish(3);
            tmem_st_32x32b_x8(tmem_base + lane_addr + S_COL + 256, p4);
            publish(4);
          } else {
            publish(3);
            if (lane == 0) mbar_arrive(&p_full[4]);
          }
        } else {
          if (timing) { tk0 = clock64(); tepi = 0; }
          mbar_wait(&s_full[0], par);
          tc_fence_after();
          if (timing) tk1 = clock64();
          int parts_ready = 1;             // the key parts of S arrive on their own barriers, in order
          auto need_part = [&](int part) {
            while (parts_ready <= part) {
              mbar_wait(&s_full[parts_ready], par);
              tc_fence_after();
              ++parts_ready;
            }
          };
          const bool warp_active = kAllRowsLive || t * QT + q * 32 < T - n_tail;  // any query row of this warp in the tile
          // Key columns are dealt in 64-column chunks: chunk c of this warp = columns [64c + 32hf, +32) (the last one
          // may hold 16).  In the specialised instance both passes are software pipelined: the TMEM load of chunk c+1
          // is in flight while chunk c is reduced / exponentiated (two register buffers, chunk loop unrolled by hand).
          // The run-time-T instance keeps load -> wait -> compute: there the second buffer only costs registers
          // (measured: 0.61 ms with the pipeline vs 0.52 ms without, B=521, T=261 forced through it).
          constexpr bool kPrefetch = T_CONST != 0;
          auto chunk_exists = [&](int c) { return warp_active && c < nchunks && c * 64 + hf * 32 < tpad; };
          auto issue_ld = [&](int c, uint32_t (&v)[32]) {
            const int c0 = c * 64 + hf * 32;
            if (tpad - c0 >= 32) {
              tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL + c0, v);
            } else {
              tmem_ld_32x32b_x16(tmem_base + lane_addr + S_COL + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
  #pragma unroll
              for (int j = 16; j < 32; ++j) v[j] = 0xff800000u;  // -inf
            }
          };
          uint32_t va[32], vb[32];
          // ---- pass 1: row max
          float m = -INFINITY;
          auto max_step = [&](int c, uint32_t (&cur)[32], uint32_t (&nxt)[32]) {
            if (!chunk_exists(c)) return;
            need_part(((kPrefetch ? c + 1 : c) * 64) / S_PART < 2 ? ((kPrefetch ? c + 1 : c) * 64) / S_PART : 2);
            if (!kPrefetch) { issue_ld(c, cur); tmem_ld_wait(); }
            const bool more = kPrefetch && chunk_exists(c + 1);
            if (more) issue_ld(c + 1, nxt);
            const int c0 = c * 64 + hf * 32;
            if (c0 + 32 <= T) m = max_group<false>(cur, 32, m);
            else                m = max_group<true>(cur, T - c0, m);
            if (more) tmem_ld_wait();
          };
          if (kPrefetch && chunk_exists(0)) { issue_ld(0, va); tmem_ld_wait(); }
          max_step(0, va, vb); max_step(1, vb, va); max_step(2, va, vb); max_step(3, vb, va); max_step(4, va, vb);
          need_part(2);   // (every tile consumes one phase of all three barriers, whatever its chunk count)
          if (timing) tk2 = clock64();
          if (kPrefetch && chunk_exists(0)) issue_ld(0, va);  // pass 2's first chunk travels during the max exchange
          // exchange buffers alternate with the tile parity: the partner warp may already be a phase ahead, and the
          // row sums written at the end of this tile are read one tile later (deferred epilogue)
          xch_max[par * 256 + hf * 128 + r] = m;
          named_bar_sync(1 + q, 64);
          m = fmaxf(xch_max[par * 256 + r], xch_max[par * 256 + 128 + r]);
          const float msl = m * p.sl2;
          if (kPrefetch && chunk_exists(0)) tmem_ld_wait();
          if (timing) tk3 = clock64();
          // ---- pass 2: exponentials, row sum, bf16 P into TMEM over the logits just consumed
          auto exp_step = [&](int c, uint32_t (&cur)[32], uint32_t (&nxt)[32]) {
            if (c >= nchunks) return;
            const bool have = chunk_exists(c), more = kPrefetch && chunk_exists(c + 1);
            if (!kPrefetch && have) { issue_ld(c, cur); tmem_ld_wait(); }
            if (more) issue_ld(c + 1, nxt);
            if (have) {
              const int c0 = c * 64 + hf * 32;
              uint32_t pk[16];
              if (c0 + 32 <= T) l += exp_group<false, POLY>(cur, p.sl2, msl, 32, pk);
              else                l += exp_group<true, POLY>(cur, p.sl2, msl, T - c0, pk);
              if (tpad - c0 >= 32) {
                tmem_st_32x32b_x16(tmem_base + lane_addr + S_COL + c0, pk);
              } else {
                tmem_st_32x32b_x8(tmem_base + lane_addr + S_COL + c0, *reinterpret_cast<uint32_t(*)[8]>(&pk[0]));
              }
            }
            if (more) tmem_ld_wait();
            if (have) tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[c]);
          };
          // the two warps of a scheduler take the owed epilogue at different chunks: while one sits in its latencies the
          // other keeps the MUFU busy
          exp_step(0, va, vb);
          if (g > 0 && hf == 0) epilogue(g - 1, pb, ph, pt);
          exp_step(1, vb, va); exp_step(2, va, vb);
          if (g > 0 && hf == 1) epilogue(g - 1, pb, ph, pt);
          exp_step(3, vb, va); exp_step(4, va, vb);
        }
        xch_sum[par * 256 + hf * 128 + r] = l;
        pb = b; ph = h; pt = t;
        if (timing) {
          tk4 = clock64();
          p.dbg[0] += tk1 - tk0; p.dbg[1] += tk2 - tk1; p.dbg[2] += tk3 - tk2; p.dbg[3] += tk4 - tk3 - tepi;
          p.dbg[6] += tepi; p.dbg[7] += 1;
        }
      }
    }
    if (g > 0) {
      named_bar_sync(1 + q, 64);           // the partner's row sums of the last tile
      epilogue(g - 1, pb, ph, pt);
    }
    tma_store_wait_all();  // the staging tile must outlive the bulk stores reading it
  } else if ((warp == 2 || warp == 3) && n_tail > 0) {
    // ---------------------------------------------------------------------------- tail queries (<= 8 rows)
    // Warp-level mma.sync.m16n8k16 on the K/V tiles already in shared memory (ldmatrix understands the TMA 128B
    // swizzle: every 8x8 sub-matrix row is one 16-byte chunk).  Keys are dealt to the two warps in blocks of 16.  The
    // logits are computed twice (once for the row max, once for the exponentials) instead of being kept: 17 key blocks
    // of accumulators would not fit the register budget this warpgroup is left with.  The logit accumulators of two
    // 8-key tiles are exactly the A fragment of the following P.V step, so P never leaves registers.  Only rows 0..7
    // of the 16-row fragments carry queries (rows 8..15 are ignored).
    const int tw = warp - 2;                       // 0..1
    const int tt = tw * 32 + lane;                 // 0..63
    const int nt = n_tail;
    const int g = lane >> 2, tq = lane & 3;        // fragment row group / thread-in-group
    float* to = reinterpret_cast<float*>(smem + OFF_TO);       // [2 warps][8 rows][64]
    float* tredm = reinterpret_cast<float*>(smem + OFF_TRED);  // [2][8]
    float* treds = tredm + NUM_TAIL_WARPS * TAIL_MAX;          // [2][8]
    const int nblk16 = tpad / 16;                // 16-key blocks, dealt round-robin to the warps
    int it = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
      const int b = pair / p.H, h = pair - b * p.H;
      const int buf = it & 1;
      const uint32_t sK = smem_u32(smem + OFF_K + buf * KV_BYTES);
      const uint32_t sV = smem_u32(smem + OFF_V + buf * KV_BYTES);
      const uint32_t sTQ = smem_u32(smem + OFF_TQ + buf * TAIL_BOX * ROW_BYTES);
      mbar_wait(&kv_full[buf], (it >> 1) & 1);
      mbar_wait(&tq_full[buf], (it >> 1) & 1);
      // ---- A fragments of the tail queries: 16 rows x 64 dims = 4 k-steps x {a0..a3}
      uint32_t qa[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        // ldmatrix.x4: matrices (rows 0-7, dims 16ks..+7), (rows 8-15, same dims), (rows 0-7, dims +8..), (rows 8-15, +8..)
        const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = 2 * ks + (lane >> 4);
        ldmatrix_x4(sTQ + row * ROW_BYTES + ((chunk ^ (row & 7)) << 4), qa[ks]);
      }
      // logits of key block blk: sacc[tile][4], tile = 8-key half
      auto block_logits = [&](int blk, float (&sacc)[2][4]) {
#pragma unroll
        for (int tile = 0; tile < 2; ++tile)
#pragma unroll
          for (int e = 0; e < 4; ++e) sacc[tile][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          // B fragments (K^T): ldmatrix.x4 -> tile0 {b0,b1}, tile1 {b0,b1}; key row = blk*16 + tile*8 + (lane&7)
          uint32_t kb[4];
          const int key = blk * 16 + ((lane >> 4) << 3) + (lane & 7);
          const int chunk = 2 * ks + ((lane >> 3) & 1);
          ldmatrix_x4(sK + key * ROW_BYTES + ((chunk ^ (key & 7)) << 4), kb);
          mma_bf16_16816(sacc[0], qa[ks], kb[0], kb[1]);
          mma_bf16_16816(sacc[1], qa[ks], kb[2], kb[3]);
        }
      };
      // ---- pass 1: row max over this warp's key blocks
      float mrow = -INFINITY;  // max over this thread's logits of row g (rows >= nt are ignored later)
#pragma unroll 1
      for (int blk = tw; blk < nblk16; blk += NUM_TAIL_WARPS) {
        float sacc[2][4];
        block_logits(blk, sacc);
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const int k0 = blk * 16 + tile * 8 + 2 * tq;  // keys of c0, c1
          if (k0 < T) mrow = fmaxf(mrow, sacc[tile][0]);
          if (k0 + 1 < T) mrow = fmaxf(mrow, sacc[tile][1]);
        }
      }
      mrow = fmaxf(mrow, __shfl_xor_sync(0xffffffffu, mrow, 1));
      mrow = fmaxf(mrow, __shfl_xor_sync(0xffffffffu, mrow, 2));
      if (tq == 0) tredm[tw * TAIL_MAX + g] = mrow;
      named_bar_sync(6, NUM_TAIL_WARPS * 32);
      const float m = fmaxf(tredm[g], tredm[TAIL_MAX + g]);
      const float msl = m * p.sl2;
      // ---- pass 2: P = exp2(...), row sums, O partial = P V with P straight from the accumulator registers
      float oacc[8][4];
#pragma unroll
      for (int nd = 0; nd < 8; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) oacc[nd][e] = 0.f;
      float lsum = 0.f;
#pragma unroll 1
      for (int blk = tw; blk < nblk16; blk += NUM_TAIL_WARPS) {
        float sacc[2][4];
        block_logits(blk, sacc);
        uint32_t pa[4];
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const int k0 = blk * 16 + tile * 8 + 2 * tq;
          const float e0 = k0 < T ? ex2(fmaf(sacc[tile][0], p.sl2, -msl)) : 0.f;
          const float e1 = k0 + 1 < T ? ex2(fmaf(sacc[tile][1], p.sl2, -msl)) : 0.f;
          lsum += e0 + e1;
          pa[2 * tile] = pack_bf16x2(e0, e1);   // rows g     (a0 / a2)
          pa[2 * tile + 1] = 0u;                // rows g + 8 (a1 / a3): unused query rows
        }
#pragma unroll
        for (int nd2 = 0; nd2 < 4; ++nd2) {
          // V as B operand (k = keys, n = dims): ldmatrix.x4.trans -> dims tile 2*nd2 {b0,b1}, tile 2*nd2+1 {b0,b1}
          uint32_t vb[4];
          const int key = blk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int chunk = 2 * nd2 + (lane >> 4);
          ldmatrix_x4_trans(sV + key * ROW_BYTES + ((chunk ^ (key & 7)) << 4), vb);
          mma_bf16_16816(oacc[2 * nd2], pa, vb[0], vb[1]);
          mma_bf16_16816(oacc[2 * nd2 + 1], pa, vb[2], vb[3]);
        }
      }
      // K, V and the tail Q rows of this pair are no longer needed by these warps
      __syncwarp();
      if (lane == 0) { mbar_arrive(&kv_empty[buf]); mbar_arrive(&tq_empty[buf]); }
      lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
      lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
      if (tq == 0) treds[tw * TAIL_MAX + g] = lsum;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd)  // rows g (c0, c1) only; dims nd*8 + 2*tq
        *reinterpret_cast<float2*>(to + (tw * TAIL_MAX + g) * HD + nd * 8 + 2 * tq) = make_float2(oacc[nd][0], oacc[nd][1]);
      named_bar_sync(6, NUM_TAIL_WARPS * 32);
      for (int i = tt; i < nt * HD / 2; i += NUM_TAIL_WARPS * 32) {
        const int j = i / (HD / 2), dp = i - j * (HD / 2);
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < NUM_TAIL_WARPS; ++w) {
          const float2 x = *reinterpret_cast<const float2*>(to + (w * TAIL_MAX + j) * HD + 2 * dp);
          acc.x += x.x; acc.y += x.y;
        }
        const float l = treds[j] + treds[TAIL_MAX + j];
        const float inv = 1.0f / l;
        const int tok = T - nt + j;
        *reinterpret_cast<uint32_t*>(p.out + (size_t(b) * T + tok) * (p.H * HD) + h * HD + 2 * dp) =
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
  if (tpad > MAX_TPAD) {   // crops above 224^2
    static const int v1 = [] { const char* e = getenv("FP_ATTN_LONG_V1"); return e ? atoi(e) : 0; }();
    return v1 ? attention_long_bf16(qkv, out, B, T, H, scale, stream) : attention_pair_bf16(qkv, out, B, T, H, scale, stream);
  }
  {
    // 224^2 crops: the two-stream kernel.  FP_ATTN_SPLIT=0 selects the single-stream kernel below (A/B measurements);
    // FP_ATTN_POLY=0x1111 / 0x5555 moves 25 / 50 % of the exponentials to the FMA pipe.
    static const int use_split = [] { const char* e = getenv("FP_ATTN_SPLIT"); return e ? atoi(e) : 1; }();
    static const unsigned poly = [] { const char* e = getenv("FP_ATTN_POLY"); return e ? unsigned(strtoul(e, nullptr, 0)) : 0u; }();
    if (T == 261 && use_split) return attention_split_bf16(qkv, out, B, T, H, scale, poly, stream);
  }
  const int rem = T % QT;
  const int n_tail = (rem > 0 && rem <= TAIL_MAX) ? rem : 0;
  const int n_normal = T / QT + ((rem > TAIL_MAX) ? 1 : 0);
  const int C = 3 * H * HD;
  CUtensorMap tmQ, tmQt, tmKV, tmOut;
  const uint64_t rows = uint64_t(B) * T;
  if (int rc = make_tmap_2d_bf16(&tmQ, qkv, rows, uint64_t(C), uint64_t(C), QT, HD)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmQt, qkv, rows, uint64_t(C), uint64_t(C), TAIL_BOX, HD)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmKV, qkv, rows, uint64_t(C), uint64_t(C), uint32_t(tpad / 2), HD)) return rc;
  if (int rc = make_tmap_2d_bf16_sw64(&tmOut, out, rows, uint64_t(H) * HD, uint64_t(H) * HD, 32)) return rc;
  Params p;
  p.out = out; p.B = B; p.T = T; p.H = H;
  p.tpad = tpad; p.n_normal = n_normal; p.n_tail = n_tail;
  // perf experiment: FP_ATTN_DBG = device address of 16 int64 phase counters (softmax warp 4 / MMA warp of CTA 0)
  p.dbg = getenv("FP_ATTN_DBG") ? reinterpret_cast<long long*>(strtoull(getenv("FP_ATTN_DBG"), nullptr, 0)) : nullptr;
  p.nchunks = (tpad + P_CHUNK_KEYS - 1) / P_CHUNK_KEYS;
  p.sl2 = scale * 1.4426950408889634f;
  const int npairs = B * H;
  const int grid = npairs < sm_count() ? npairs : sm_count();
  ProfScope prof(PROF_ATTENTION, 4.0 * double(B) * H * double(T) * T * HD, 1, stream);
  const bool special = T == 261 && !getenv("FP_ATTN_GENERIC");  // 224^2 crops: (224/14)^2 + 5 tokens
  // POLY = 0: every exponential on the MUFU.  Moving 25-50 % of them to the FMA pipe (ex2_poly_x2) was measured twice
  // and is slower even now that the exponential pass sits at the MUFU floor (B=521, T=261: 0.283 ms without, 0.318 /
  // 0.333 / 0.324 ms at 25 / 37.5 / 50 %): with two softmax warps per scheduler the longer dependent FMA chains cost
  // more latency than the MUFU cycles they free.
#define FP_LAUNCH_ATTN_POLY(TIMING_, TC_)                                                                            \
  do {                                                                                                               \
    auto kern = attention_kernel<TIMING_, TC_, 0u>;                                                                  \
    FP_ENSURE_DYN_SMEM(kern, SMEM_BYTES);                                                                            \
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmQt, tmKV, tmOut, p);                                      \
  } while (0)
  if (p.dbg && special) FP_LAUNCH_ATTN_POLY(true, 261);
  else if (p.dbg)       FP_LAUNCH_ATTN_POLY(true, 0);
  else if (special)     FP_LAUNCH_ATTN_POLY(false, 261);
  else                  FP_LAUNCH_ATTN_POLY(false, 0);
#undef FP_LAUNCH_ATTN_POLY
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fp
