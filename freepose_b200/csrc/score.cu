// Per-hypothesis score + top-k.
//
// Reference arithmetic (pose_estimator.py:85-90, online_pose_estimator.py:68-79), model dtype bf16:
//     tn = F.normalize(feats_t, dim=-1)   -> norm = bf16(sqrt(sum t^2)); tn = bf16(t / max(norm, eps))
//     qn = F.normalize(query,  dim=-1)
//     s  = einsum('b n d, b n d -> b n')  -> bf16(sum_d tn*qn)        (fp32 accumulate)
//     score = s.mean(-1)                  -> bf16(sum_n s / P)        (fp32 accumulate)
//     fine / mask_scores:  score = sum_n(s*w) / sum_n(w) in fp32
// The kernel reproduces those rounding points and FIXES the fp32 summation order so that the CPU
// oracle (oracle/score.py, "engine order") can restate it bit-exactly:
//   * a D-long reduction: lane l owns elements c*256 + l*8 + j and keeps two running sums, over its even and over its
//     odd j (c outer), acc = acc + a*b (a*b is exact in fp32 for bf16 inputs); even + odd, then the xor-butterfly
//     16,8,4,2,1 (rowops.cuh);
//   * a P-long reduction: lane l accumulates n = l, l+32, ... ascending, then the same butterfly.
// HBM-bound: every template token row (2 KB) is read exactly once (score_rows_kernel), the per-patch cosines take a
// second, tiny pass (score_reduce_kernel).
#include "common.cuh"
#include "kernels.h"
#include "rowops.cuh"

namespace fp {

namespace {

constexpr int SC_WARPS = 16;        // prep_query: one warp per query row
constexpr int ROW_WARPS = 8;        // rows kernel: 256 threads, two CTAs per SM
constexpr int ROWS_PER_ITEM = 4;    // hypotheses per work item: 4 x 2 KB template rows in flight per warp
constexpr int RED_WARPS = 8;        // reduce kernel: one warp per hypothesis
using namespace rowops;

// qn = normalised (or verbatim) query tokens, one warp per row; also zeroes the rows kernel's work counter
__global__ void __launch_bounds__(SC_WARPS * 32)
prep_query_kernel(const bf16* __restrict__ q, bf16* __restrict__ qn, int P, int D, int normalise,
                  unsigned* __restrict__ counter) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0u;
  const int n = blockIdx.x * SC_WARPS + warp;
  if (n >= P) return;
  const int chunks = D / 256;
  uint4 u[MAX_CHUNKS];
  load_row(q + size_t(n) * D, lane, chunks, u);
  uint4* op = reinterpret_cast<uint4*>(qn + size_t(n) * D);
  if (!normalise) {
#pragma unroll
    for (int c = 0; c < MAX_CHUNKS; ++c)
      if (c < chunks) op[c * 32 + lane] = u[c];
    return;
  }
  const float nrm = row_norm(u, chunks);
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < chunks) {
      float f[8];
      unpack8(u[c], f);
      uint4 o;
      o.x = pack_bf16x2(__fdiv_rn(f[0], nrm), __fdiv_rn(f[1], nrm));
      o.y = pack_bf16x2(__fdiv_rn(f[2], nrm), __fdiv_rn(f[3], nrm));
      o.z = pack_bf16x2(__fdiv_rn(f[4], nrm), __fdiv_rn(f[5], nrm));
      o.w = pack_bf16x2(__fdiv_rn(f[6], nrm), __fdiv_rn(f[7], nrm));
      op[c * 32 + lane] = o;
    }
}

// Peer-memory exchange (fp_score_publish): the score stage itself is the "send" side of the
// all-gather.  Every rank owns an exchange buffer [2 parities][world * per_rank scores | world flags] that its peers map
// through CUDA IPC; a score is stored into the slot (rank * per_rank + b) of EVERY rank's buffer over NVLink as it is
// produced, and the last CTA to finish publishes "rank r is complete for this epoch" in every buffer.  The receiving
// side is the top-k kernel, which waits for the world flags of its own buffer.  No collective call, no extra kernel.
struct PeerExchange {
  float* const* peers;      // device array [world]: the ranks' exchange buffers (this rank's own included), or nullptr
  int world, rank, per_rank;
  unsigned epoch;           // increases by one per exchange; parity selects the half of the buffer
  unsigned* done;           // device counter of finished CTAs (this rank's, zero between launches)
};
__host__ __device__ inline size_t exchange_half_floats(int world, int per_rank) {
  return (size_t(world) * per_rank + size_t(world) + 63) / 64 * 64;     // scores, then one flag per rank; 256-byte multiple
}

// Stage 1, HBM-bound: per-patch cosines s[b][n] = bf16(sum_d bf16(t/|t|) * qn).  Work item = (patch n, 4 consecutive
// hypotheses): the warp issues the four 2 KB template rows at once (8 KB in flight per warp, 16 warps per SM), reads
// the query row n once for the four, and takes its next item from a global counter -- no tail, whatever B is.
//
// tn = bf16(t / nrm) is evaluated as bf16(t * fl(1/nrm)) -- the same value, always: t and nrm are bf16 (8-bit
// significands T, N), a bf16 rounding boundary is m * 2^k with m an odd 9-bit integer, and T * 2^c = N * m has no
// solution (m is odd and m > T), so |t/nrm - boundary| >= boundary / (N * m) > 2^-17 relative, while the reciprocal and
// the product together err by < 2^-22 (tests/test_oracle_golden.py::test_reciprocal_normalisation_is_exact runs all
// significand pairs).
template <int CHUNKS>   // D / 256 at compile time: no per-chunk predicates in the unrolled row loops
__global__ void __launch_bounds__(ROW_WARPS * 32, 2)
score_rows_kernel(const bf16* __restrict__ feats_t, const bf16* __restrict__ qn, int B, int P,
                  float* __restrict__ patch, unsigned* __restrict__ counter) {
  const int lane = threadIdx.x & 31;
  constexpr int chunks = CHUNKS;
  constexpr int D = CHUNKS * 256;
  const unsigned total = unsigned((B + ROWS_PER_ITEM - 1) / ROWS_PER_ITEM) * unsigned(P);
  for (;;) {
    unsigned item = 0;
    if (lane == 0) item = atomicAdd(counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    const int g = int(item / unsigned(P)), n = int(item - unsigned(g) * unsigned(P));
    const int b0 = g * ROWS_PER_ITEM;
    uint4 t[ROWS_PER_ITEM][MAX_CHUNKS];
#pragma unroll
    for (int r = 0; r < ROWS_PER_ITEM; ++r)
      if (b0 + r < B) load_row(feats_t + (size_t(b0 + r) * P + n) * D, lane, chunks, t[r]);
    f32x2 qp[MAX_CHUNKS][4];   // the query row as fp32 pairs (even, odd element of every bf16x2 word)
    {
      uint4 qv[MAX_CHUNKS];
      load_row(qn + size_t(n) * D, lane, chunks, qv);
#pragma unroll
      for (int c = 0; c < MAX_CHUNKS; ++c)
        if (c < chunks) { qp[c][0] = word2(qv[c].x); qp[c][1] = word2(qv[c].y); qp[c][2] = word2(qv[c].z); qp[c][3] = word2(qv[c].w); }
    }
#pragma unroll
    for (int r = 0; r < ROWS_PER_ITEM; ++r) {
      if (b0 + r >= B) break;
      const float rnrm = __frcp_rn(row_norm(t[r], chunks));
      const f32x2 rn2 = pack2(rnrm, rnrm);
      f32x2 acc2 = pack2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < MAX_CHUNKS; ++c)
        if (c < chunks) {
          const uint32_t w[4] = {t[r][c].x, t[r][c].y, t[r][c].z, t[r][c].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float p0, p1;
            unpack2(mul2(word2(w[i]), rn2), p0, p1);          // t * (1/nrm), both elements of the word in one FMUL2
            acc2 = fma2(word2(pack_bf16x2(p0, p1)), qp[c][i], acc2);   // bf16(t/nrm) * q: exact products, two running sums
          }
        }
      float acc = hsum2(acc2);
      acc = warp_sum(acc);
      if (lane == 0) patch[size_t(b0 + r) * P + n] = bf16_round(acc);
    }
  }
}

// Stage 2: one warp per hypothesis reduces its P per-patch cosines (fixed order: lane l sums n = l, l+32, ... ascending,
// then the butterfly) to the score -- mean in bf16, or the mask-weighted mean in fp32 -- and, on the exchange path,
// stores it into every rank's buffer; the last CTA raises this rank's flags.
__global__ void __launch_bounds__(RED_WARPS * 32)
score_reduce_kernel(const float* __restrict__ patch, const float* __restrict__ weights, int B, int P,
                    float* __restrict__ scores, const PeerExchange px) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * RED_WARPS + warp;
  if (b < B) {
    const float* s = patch + size_t(b) * P;
    float v;
    if (weights == nullptr) {
      float acc = 0.f;
      for (int n = lane; n < P; n += 32) acc = __fadd_rn(acc, s[n]);
      acc = warp_sum(acc);
      v = bf16_round(__fdiv_rn(acc, float(P)));
    } else {
      const float* w = weights + size_t(b) * P;
      float num = 0.f, den = 0.f;
      for (int n = lane; n < P; n += 32) {
        const float wn = w[n];
        num = __fadd_rn(num, __fmul_rn(s[n], wn));
        den = __fadd_rn(den, wn);
      }
      num = warp_sum(num);
      den = warp_sum(den);
      v = __fdiv_rn(num, den);
    }
    if (lane == 0) scores[b] = v;
    if (px.peers != nullptr && lane < px.world) {
      // the same value into this rank's slot of every rank's buffer (lane r -> rank r; one NVLink store each)
      float* dst = px.peers[lane] + (px.epoch & 1u) * exchange_half_floats(px.world, px.per_rank);
      dst[size_t(px.rank) * px.per_rank + b] = v;
    }
  }
  if (px.peers != nullptr) {
    // publish: stores above -> system-scope fence -> count this CTA; the last CTA raises this rank's flag everywhere
    __threadfence_system();
    __syncthreads();
    if (warp == 0) {
      unsigned last = 0;
      if (lane == 0) {
        __threadfence_system();
        last = atomicAdd(px.done, 1u) == gridDim.x - 1;
        __threadfence_system();
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        if (lane == 0) *px.done = 0;
        if (lane < px.world) {
          float* half = px.peers[lane] + (px.epoch & 1u) * exchange_half_floats(px.world, px.per_rank);
          unsigned* flags = reinterpret_cast<unsigned*>(half + size_t(px.world) * px.per_rank);
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags + px.rank), "r"(px.epoch) : "memory");
        }
      }
    }
  }
}

// Deterministic top-k: descending value, ties -> lowest index (NaN sorts first, like torch.topk).
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn != bn) return vn;
  if (!vn && v != bv) return v > bv;
  return i < bi;
}

__global__ void __launch_bounds__(1024)
topk_kernel(const float* __restrict__ scores, int B, int k, int* __restrict__ idx_out, float* __restrict__ val_out,
            uint8_t* __restrict__ taken, const unsigned* __restrict__ wait_flags, int wait_world, unsigned wait_epoch) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (wait_flags != nullptr) {
    // receiving side of the peer-memory exchange: every rank's flag must have reached this epoch (bounded spin: a
    // protocol bug or a dead peer traps instead of hanging the device)
    if (threadIdx.x < wait_world) {
      unsigned v, spins = 0;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(wait_flags + threadIdx.x) : "memory");
        if (int(v - wait_epoch) >= 0) break;
        if (++spins > (1u << 25)) __trap();      // ~10 s of polling
      } while (true);
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < B; i += blockDim.x) taken[i] = 0;
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    float bv = 0.f;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < B; i += blockDim.x)
      if (!taken[i]) {
        const float v = scores[i];
        if (bi == 0x7fffffff || better(v, i, bv, bi)) { bv = v; bi = i; }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || better(ov, oi, bv, bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = sv[lane];
      bi = si[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || better(ov, oi, bv, bi))) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        if (bi != 0x7fffffff) {
          idx_out[r] = bi;
          val_out[r] = bv;
          taken[bi] = 1;
        } else {
          idx_out[r] = -1;
          val_out[r] = 0.f;
        }
      }
    }
    __syncthreads();
  }
}

// FFA pooling: mask (res x res, u8) -> 14x14 max-pool -> mean of the selected patch tokens.
// Reference (extract_retrieval_features.py:51-57): cv2.resize(mask, (g, g), INTER_AREA) > 0 is true
// exactly when any pixel of the 14x14 cell is set; feat[mask].mean(0) on bf16 rounds the fp32-accumulated
// mean to bf16, then .float().  Order: patches ascending, acc = acc + x.
__global__ void __launch_bounds__(256)
ffa_kernel(const bf16* __restrict__ feats, const uint8_t* __restrict__ masks, int res, int g, int D,
           float* __restrict__ out, int* __restrict__ valid) {
  extern __shared__ uint8_t cell[];  // [g*g]
  const int v = blockIdx.x;
  const uint8_t* m = masks + size_t(v) * res * res;
  const int P = g * g;
  for (int c = threadIdx.x; c < P; c += blockDim.x) {
    const int cy = c / g, cx = c - cy * g;
    int any = 0;
    for (int y = 0; y < 14; ++y) {
      const uint8_t* rowp = m + size_t(cy * 14 + y) * res + cx * 14;
#pragma unroll
      for (int x = 0; x < 14; ++x) any |= rowp[x];
    }
    cell[c] = any ? 1 : 0;
  }
  __syncthreads();
  int cnt = 0;
  for (int c = 0; c < P; ++c) cnt += cell[c];
  const bf16* f = feats + size_t(v) * P * D;
  for (int d = threadIdx.x * 2; d < D; d += blockDim.x * 2) {
    float a0 = 0.f, a1 = 0.f;
    for (int c = 0; c < P; ++c)
      if (cell[c]) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(f + size_t(c) * D + d);
        a0 = __fadd_rn(a0, bf16lo(u));
        a1 = __fadd_rn(a1, bf16hi(u));
      }
    // empty mask -> 0/0 = NaN, which the reference detects and skips (extract_retrieval_features.py:59-65)
    out[size_t(v) * D + d] = bf16_round(__fdiv_rn(a0, float(cnt)));
    out[size_t(v) * D + d + 1] = bf16_round(__fdiv_rn(a1, float(cnt)));
  }
  if (threadIdx.x == 0 && valid != nullptr) valid[v] = cnt;
}

}  // namespace

// workspace: [query copy P*D bf16][top-k scratch B bytes][work counter][per-patch cosines B*P fp32]
static size_t ws_taken_off(int P, int D) { return size_t(P) * D * sizeof(bf16); }
static size_t ws_counter_off(int B, int P, int D) { return (ws_taken_off(P, D) + size_t(B) + 15) / 16 * 16; }
static size_t ws_patch_off(int B, int P, int D) { return ws_counter_off(B, P, D) + 16; }
size_t score_workspace_bytes(int B, int P, int D) { return ws_patch_off(B, P, D) + size_t(B) * P * sizeof(float) + 256; }

// prep_query -> rows -> reduce on `stream`; scores_out[b] (and every peer's slot when px.peers is set)
static int score_stages(const bf16* feats_t, const bf16* feat_q, const float* weights, int B, int P, int D,
                        int normalise_query, float* scores_out, float* patch_scores_out, const PeerExchange& px,
                        void* workspace, cudaStream_t stream) {
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  bf16* qn = reinterpret_cast<bf16*>(ws);
  unsigned* counter = reinterpret_cast<unsigned*>(ws + ws_counter_off(B, P, D));
  float* patch = patch_scores_out ? patch_scores_out : reinterpret_cast<float*>(ws + ws_patch_off(B, P, D));
  prep_query_kernel<<<(P + SC_WARPS - 1) / SC_WARPS, SC_WARPS * 32, 0, stream>>>(feat_q, qn, P, D, normalise_query, counter);
  FP_CUDA(cudaGetLastError());
  const long long items = (long long)((B + ROWS_PER_ITEM - 1) / ROWS_PER_ITEM) * P;
  const long long want = (items + ROW_WARPS - 1) / ROW_WARPS;
  const int grid = int(want < 2LL * sm_count() ? want : 2LL * sm_count());
  switch (D / 256) {
    case 1: score_rows_kernel<1><<<grid, ROW_WARPS * 32, 0, stream>>>(feats_t, qn, B, P, patch, counter); break;
    case 2: score_rows_kernel<2><<<grid, ROW_WARPS * 32, 0, stream>>>(feats_t, qn, B, P, patch, counter); break;
    case 3: score_rows_kernel<3><<<grid, ROW_WARPS * 32, 0, stream>>>(feats_t, qn, B, P, patch, counter); break;
    default: score_rows_kernel<4><<<grid, ROW_WARPS * 32, 0, stream>>>(feats_t, qn, B, P, patch, counter); break;
  }
  FP_CUDA(cudaGetLastError());
  score_reduce_kernel<<<(B + RED_WARPS - 1) / RED_WARPS, RED_WARPS * 32, 0, stream>>>(patch, weights, B, P, scores_out, px);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int score_topk(const bf16* feats_t, const bf16* feat_q, const float* weights, int B, int P, int D,
               int normalise_query, float* scores_out, float* patch_scores_out, int k, int* topk_idx,
               float* topk_val, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  FP_REQUIRE(D % 256 == 0 && D <= 256 * MAX_CHUNKS, "score: D=%d must be a multiple of 256 and <= 1024", D);
  FP_REQUIRE(B >= 0 && P > 0, "score: bad shape B=%d P=%d", B, P);
  FP_REQUIRE(k >= 0 && k <= B, "score: k=%d out of range for B=%d hypotheses", k, B);
  FP_REQUIRE(workspace_bytes >= score_workspace_bytes(B, P, D), "score: workspace too small");
  FP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "score: workspace must be 16-byte aligned");
  if (B == 0) return 0;
  ProfScope prof(PROF_SCORE, (double(B) + 1) * P * D * 2, k > 0 ? 4 : 3, stream);
  PeerExchange none{};
  if (int rc = score_stages(feats_t, feat_q, weights, B, P, D, normalise_query, scores_out, patch_scores_out, none,
                            workspace, stream)) return rc;
  if (k > 0) {
    uint8_t* taken = reinterpret_cast<uint8_t*>(workspace) + ws_taken_off(P, D);
    topk_kernel<<<1, 1024, 0, stream>>>(scores_out, B, k, topk_idx, topk_val, taken, nullptr, 0, 0u);
    FP_CUDA(cudaGetLastError());
  }
  return 0;
}

size_t exchange_bytes(int world, int per_rank) { return 2 * exchange_half_floats(world, per_rank) * sizeof(float) + 256; }

// scores of this rank's B (<= per_rank) hypotheses -> slot `rank` of every rank's exchange buffer + completion flags.
// `peers` = device array of the world buffer pointers; own buffer = peers[rank] as seen by this process.
int score_publish(const bf16* feats_t, const bf16* feat_q, const float* weights, int B, int P, int D, int normalise_query,
                  float* const* peers_dev, float* own_buffer, int rank, int world, int per_rank, unsigned epoch,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  FP_REQUIRE(D % 256 == 0 && D <= 256 * MAX_CHUNKS, "score: D=%d must be a multiple of 256 and <= 1024", D);
  FP_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world, "publish: rank %d / world %d (at most 32 ranks)", rank, world);
  FP_REQUIRE(B > 0 && B <= per_rank && P > 0, "publish: B=%d must be in [1, %d] (every rank needs at least one hypothesis)", B, per_rank);
  FP_REQUIRE(workspace_bytes >= score_workspace_bytes(B, P, D) + 256, "publish: workspace too small");
  FP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "score: workspace must be 16-byte aligned");
  FP_REQUIRE(peers_dev != nullptr && own_buffer != nullptr, "publish: null exchange buffers");
  // the CTA counter of the publish lives behind everything else in the workspace: the caller keeps this workspace
  // across calls and zeroes it once (the last CTA of every launch resets the counter)
  unsigned* done = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(workspace) + (score_workspace_bytes(B, P, D) + 15) / 16 * 16);
  PeerExchange px{peers_dev, world, rank, per_rank, epoch, done};
  float* local = own_buffer + (epoch & 1u) * exchange_half_floats(world, per_rank) + size_t(rank) * per_rank;
  ProfScope prof(PROF_SCORE, (double(B) + 1) * P * D * 2, 3, stream);
  return score_stages(feats_t, feat_q, weights, B, P, D, normalise_query, local, nullptr, px, workspace, stream);
}

// receiving side: waits for all ranks' flags of `epoch` in this rank's own buffer, then the deterministic top-k over
// the first n_total slots
int topk_after_exchange(float* own_buffer, int world, int per_rank, int n_total, unsigned epoch, int k, int* topk_idx,
                        float* topk_val, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  FP_REQUIRE(n_total > 0 && n_total <= world * per_rank && k > 0 && k <= n_total, "topk_after_exchange: bad sizes");
  FP_REQUIRE(workspace_bytes >= size_t(n_total), "topk: workspace too small");
  float* half = own_buffer + (epoch & 1u) * exchange_half_floats(world, per_rank);
  const unsigned* flags = reinterpret_cast<const unsigned*>(half + size_t(world) * per_rank);
  ProfScope prof(PROF_SCORE, double(n_total) * 4, 1, stream);
  topk_kernel<<<1, 1024, 0, stream>>>(half, n_total, k, topk_idx, topk_val, reinterpret_cast<uint8_t*>(workspace), flags, world,
                                      epoch);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int topk_only(const float* scores, int B, int k, int* topk_idx, float* topk_val, void* workspace,
              size_t workspace_bytes, cudaStream_t stream) {
  FP_REQUIRE(k >= 0 && k <= B, "topk: k=%d out of range for B=%d", k, B);
  FP_REQUIRE(workspace_bytes >= size_t(B), "topk: workspace too small");
  if (k == 0) return 0;
  ProfScope prof(PROF_SCORE, double(B) * 4, 1, stream);
  topk_kernel<<<1, 1024, 0, stream>>>(scores, B, k, topk_idx, topk_val, reinterpret_cast<uint8_t*>(workspace), nullptr, 0, 0u);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int ffa_pool(const bf16* feats, const uint8_t* masks, int V, int res, int g, int D, float* out, int* valid,
             cudaStream_t stream) {
  FP_REQUIRE(res == g * 14, "ffa: mask resolution %d != 14 * grid %d", res, g);
  FP_REQUIRE(D % 2 == 0, "ffa: D must be even");
  if (V <= 0) return 0;
  ProfScope prof(PROF_SCORE, double(V) * g * g * D * 2, 1, stream);
  ffa_kernel<<<V, 256, size_t(g) * g, stream>>>(feats, masks, res, g, D, out, valid);
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fp
