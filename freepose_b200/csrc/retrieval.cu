// Mesh retrieval: coarse scan of the FFA/cls feature database + top-100, fine per-view
// re-rank of the candidates, and the soft-vote accumulation of the video script.
//
// Reference arithmetic (scripts/extract_proposals_ground.py:39-41,136-160; ..._video.py:148-190), model dtype bf16:
//     db   = F.normalize(db.to(bf16), dim=-1)                      -> rowops::normalise_row rounding points
//     s    = (db @ feature).float()                                 -> bf16(sum_d db*q), fp32 accumulate
//     s, I = torch.topk(s, 100)                                     -> descending, ties -> lowest index (torch: unspecified)
//     fine = F.normalize(views.to(bf16)); p = (fine @ feature).float(); torch.topk(p, k).values.cpu().numpy().mean()
//     best = max(scores, key=scores.get)                            -> first maximum in coarse-rank order
//     video: s_frame[I] = scores (dense zeros elsewhere); mean over frames; topk(1)
// HBM-bound: the coarse scan reads every database row (2 KB) exactly once for ALL queries of a call (the reference
// re-reads the 94 MB database per proposal); the fine stage reads the candidates' view rows from a device-resident
// store (46 037 meshes x 600 views x 1024 bf16 = 56.6 GB fits the 180 GB of HBM3e; the reference does 100 np.load +
// H2D copies per proposal).  The fp32 summation order is the one of score.cu so oracle/retrieval.py restates it
// bit-exactly.
#include "common.cuh"
#include "kernels.h"
#include "rowops.cuh"

namespace fp {

namespace {

using namespace rowops;

constexpr int RT_WARPS = 8;
constexpr int MAX_QUERIES = 32;   // queries staged in shared memory per scan launch
constexpr int TOPK_MAX = 1024;
constexpr int FINE_MAX_VIEWS = 4096;
constexpr int FINE_MAX_K = 128;

// ------------------------------------------------------------------------------------------ F.normalize rows
template <bool SRC_F32>
__global__ void __launch_bounds__(RT_WARPS * 32)
normalize_rows_kernel(const void* __restrict__ src, bf16* __restrict__ dst, long long M, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = D / 256;
  for (long long m = (long long)blockIdx.x * RT_WARPS + warp; m < M; m += (long long)gridDim.x * RT_WARPS) {
    uint4 u[MAX_CHUNKS];
    if (SRC_F32) {
      const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + m * D);
#pragma unroll
      for (int c = 0; c < MAX_CHUNKS; ++c)
        if (c < chunks) {
          const float4 a = p[(c * 32 + lane) * 2], b = p[(c * 32 + lane) * 2 + 1];  // .to(bfloat16): round to nearest even
          u[c] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
        }
    } else {
      load_row(reinterpret_cast<const bf16*>(src) + m * D, lane, chunks, u);
    }
    normalise_row(u, chunks);
    uint4* op = reinterpret_cast<uint4*>(dst + m * D);
#pragma unroll
    for (int c = 0; c < MAX_CHUNKS; ++c)
      if (c < chunks) op[c * 32 + lane] = u[c];
  }
}

// ------------------------------------------------------------------------------------------ coarse scan
// scores[q, m] = bf16(db[m] . query[q]); one warp per database row, queries in shared memory.
__global__ void __launch_bounds__(RT_WARPS * 32)
scan_kernel(const bf16* __restrict__ db, const bf16* __restrict__ queries, long long M, int D, int Q,
            float* __restrict__ scores) {
  extern __shared__ uint4 s_q[];  // [Q][D/8]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = D / 256;
  const int row16 = D / 8;
  for (int i = threadIdx.x; i < Q * row16; i += blockDim.x) s_q[i] = reinterpret_cast<const uint4*>(queries)[i];
  __syncthreads();
  // the NEXT row of this warp is in flight while the current one is reduced against the Q queries
  const long long stride = (long long)gridDim.x * RT_WARPS;
  long long m = (long long)blockIdx.x * RT_WARPS + warp;
  uint4 nxt[MAX_CHUNKS];
  if (m < M) load_row(db + m * D, lane, chunks, nxt);
  for (; m < M; m += stride) {
    uint4 u[MAX_CHUNKS];
#pragma unroll
    for (int c = 0; c < MAX_CHUNKS; ++c) u[c] = nxt[c];
    if (m + stride < M) load_row(db + (m + stride) * D, lane, chunks, nxt);
    for (int q = 0; q < Q; ++q) {
      uint4 qv[MAX_CHUNKS];
#pragma unroll
      for (int c = 0; c < MAX_CHUNKS; ++c)
        if (c < chunks) qv[c] = s_q[q * row16 + c * 32 + lane];
      const float d = row_dot(u, qv, chunks);
      if (lane == 0) scores[(long long)q * M + m] = bf16_round(d);
    }
  }
}

// ------------------------------------------------------------------------------------------ large top-k
// Order-preserving key: larger key = earlier in torch.topk order (NaN first, then descending value; -0 == +0).
__device__ __forceinline__ uint32_t topk_key(float x) {
  if (x != x) return 0xffffffffu;
  if (x == 0.f) return 0x80000000u;
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// exclusive prefix count of `flag` over the 1024 threads of the block (index order); total returned to everyone
__device__ __forceinline__ int block_rank(bool flag, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t bal = __ballot_sync(0xffffffffu, flag);
  const int within = __popc(bal & ((1u << lane) - 1));
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  int before = 0, tot = 0;
  for (int w = 0; w < 32; ++w) {
    const int c = s_warp[w];
    before += w < warp ? c : 0;
    tot += c;
  }
  __syncthreads();
  total = tot;
  return before + within;
}

__global__ void __launch_bounds__(1024)
topk_large_kernel(const float* __restrict__ scores, long long M, int k, int* __restrict__ idx_out,
                  float* __restrict__ val_out) {
  __shared__ int hist[256];
  __shared__ int s_warp[32];
  __shared__ uint32_t s_prefix, s_mask;
  __shared__ int s_krem;
  __shared__ unsigned long long sel[TOPK_MAX];
  const float* x = scores + (long long)blockIdx.x * M;
  int* io = idx_out + (long long)blockIdx.x * k;
  float* vo = val_out + (long long)blockIdx.x * k;
  const int tid = threadIdx.x;
  if (tid == 0) { s_prefix = 0; s_mask = 0; s_krem = k; }
  // ---- radix select (4 passes, most significant byte first): the key of the k-th element in topk order
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix, mask = s_mask;
    for (long long i0 = 0; i0 < M; i0 += 1024) {
      const long long i = i0 + tid;
      uint32_t key = 0;
      const bool in = i < M && ((key = topk_key(x[i])) & mask) == prefix;
      const int bin = in ? int((key >> shift) & 0xff) : -1;
      // warp-aggregated shared atomics: cosine scores share their leading bytes, one hot bin would serialise 46k atomics
      const uint32_t peers = __match_any_sync(0xffffffffu, bin);
      if (in && (tid & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
    }
    __syncthreads();
    if (tid == 0) {
      int krem = s_krem, d = 255;
      for (; d > 0; --d) {
        if (hist[d] >= krem) break;
        krem -= hist[d];
      }
      s_krem = krem;
      s_prefix = prefix | (uint32_t(d) << shift);
      s_mask = mask | (0xffu << shift);
    }
    __syncthreads();
  }
  const uint32_t thr = s_prefix;
  const int need_eq = s_krem;           // elements equal to the threshold key to take (lowest indices first)
  const int n_gt = k - need_eq;         // elements strictly ahead of the threshold
  // ---- ordered collection
  for (int i = tid; i < TOPK_MAX; i += 1024) sel[i] = 0ull;
  __syncthreads();
  int base_gt = 0, base_eq = 0;
  for (long long i0 = 0; i0 < M; i0 += 1024) {
    const long long i = i0 + tid;
    const uint32_t key = i < M ? topk_key(x[i]) : 0u;
    const bool gt = i < M && key > thr, eq = i < M && key == thr;
    int tot_gt, tot_eq;
    const int r_gt = block_rank(gt, s_warp, tot_gt);
    const int r_eq = block_rank(eq, s_warp, tot_eq);
    const unsigned long long item = (static_cast<unsigned long long>(key) << 32) | (0xffffffffu - uint32_t(i));
    if (gt) sel[base_gt + r_gt] = item;
    if (eq && base_eq + r_eq < need_eq) sel[n_gt + base_eq + r_eq] = item;
    base_gt += tot_gt;
    base_eq += tot_eq;
  }
  __syncthreads();
  // ---- bitonic sort, descending on (key, -index)
  int n = 1;
  while (n < k) n <<= 1;
  for (int size = 2; size <= n; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < n / 2; t += 1024) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = sel[lo], b = sel[hi];
        if ((a < b) == desc) { sel[lo] = b; sel[hi] = a; }
      }
      __syncthreads();
    }
  for (int r = tid; r < k; r += 1024) {
    const int i = int(0xffffffffu - uint32_t(sel[r] & 0xffffffffull));
    io[r] = i;
    vo[r] = x[i];
  }
}

// ------------------------------------------------------------------------------------------ fine re-rank
// out[q, c] = mean(topk_k(views(cand[q, c]) . query[q]))  with numpy's float32 pairwise summation for n <= 128.
__global__ void __launch_bounds__(RT_WARPS * 32)
fine_kernel(const bf16* __restrict__ views, const long long* __restrict__ view_start,
            const int* __restrict__ view_count, const int* __restrict__ cand, const bf16* __restrict__ queries, int C,
            int D, int k, float* __restrict__ out) {
  __shared__ float s_sc[FINE_MAX_VIEWS];
  __shared__ float s_top[FINE_MAX_K];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x, q = blockIdx.y;
  const int mesh = cand[q * C + c];
  const int chunks = D / 256;
  if (mesh < 0) {
    if (threadIdx.x == 0) out[q * C + c] = -INFINITY;
    return;
  }
  const long long v0 = view_start[mesh];
  const int V = view_count[mesh];
  uint4 qv[MAX_CHUNKS];
  load_row(queries + size_t(q) * D, lane, chunks, qv);
  for (int v = warp; v < V; v += RT_WARPS) {
    uint4 u[MAX_CHUNKS];
    load_row(views + (v0 + v) * D, lane, chunks, u);
    const float d = row_dot(u, qv, chunks);
    if (lane == 0) s_sc[v] = bf16_round(d);
  }
  __syncthreads();
  if (warp != 0) return;
  // k rounds of (max value, lowest index); selected entries are knocked out with a sentinel flag in the sign of the index
  for (int r = 0; r < k; ++r) {
    float bv = 0.f;
    int bi = 0x7fffffff;
    for (int v = lane; v < V; v += 32) {
      const float x = s_sc[v];
      const uint32_t key = topk_key(x);
      if (__float_as_uint(x) == 0xffffffffu) continue;  // taken (this NaN payload never comes out of bf16_round)
      if (bi == 0x7fffffff || key > topk_key(bv) || (key == topk_key(bv) && v < bi)) { bv = x; bi = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff &&
          (bi == 0x7fffffff || topk_key(ov) > topk_key(bv) || (topk_key(ov) == topk_key(bv) && oi < bi))) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      s_top[r] = bi != 0x7fffffff ? bv : __uint_as_float(0x7fc00000u);  // k > V: the host rejects it (torch raises)
      if (bi != 0x7fffffff) s_sc[bi] = __uint_as_float(0xffffffffu);
    }
    __syncwarp();
  }
  if (lane == 0) {
    // numpy add.reduce on a contiguous float32 vector (n <= 128): < 8 sequential; otherwise eight strided partial sums,
    // combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the remainder sequentially; mean = sum / n in float32
    float sum;
    if (k < 8) {
      sum = s_top[0];
      for (int i = 1; i < k; ++i) sum = __fadd_rn(sum, s_top[i]);
    } else {
      float r8[8];
      for (int j = 0; j < 8; ++j) r8[j] = s_top[j];
      int i = 8;
      for (; i < k - (k % 8); i += 8)
        for (int j = 0; j < 8; ++j) r8[j] = __fadd_rn(r8[j], s_top[i + j]);
      sum = __fadd_rn(__fadd_rn(__fadd_rn(r8[0], r8[1]), __fadd_rn(r8[2], r8[3])),
                      __fadd_rn(__fadd_rn(r8[4], r8[5]), __fadd_rn(r8[6], r8[7])));
      for (; i < k; ++i) sum = __fadd_rn(sum, s_top[i]);
    }
    out[q * C + c] = __fdiv_rn(sum, float(k));
  }
}

// ------------------------------------------------------------------------------------------ video soft vote
__global__ void softvote_add_kernel(float* __restrict__ acc, const int* __restrict__ idx, const float* __restrict__ val,
                                    int P, int C, long long M) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * C) return;
  const int p = i / C, m = idx[i];
  if (m >= 0) acc[(long long)p * M + m] = __fadd_rn(acc[(long long)p * M + m], val[i]);  // indices are unique per row
}

__global__ void softvote_mean_kernel(const float* __restrict__ acc, float* __restrict__ out, long long n, float frames) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __fdiv_rn(acc[i], frames);
}

bool row_dim_ok(int D) { return D % 256 == 0 && D > 0 && D <= 256 * MAX_CHUNKS; }

}  // namespace

int normalize_rows(const void* src, int src_is_f32, long long M, int D, bf16* dst, cudaStream_t stream) {
  FP_REQUIRE(row_dim_ok(D), "normalize_rows: D=%d must be a multiple of 256 and <= 1024", D);
  FP_REQUIRE(M >= 0, "normalize_rows: negative row count");
  FP_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
             "normalize_rows: pointers must be 16-byte aligned");
  if (M == 0) return 0;
  ProfScope prof(PROF_SCORE, double(M) * D * (src_is_f32 ? 6 : 4), 1, stream);
  const long long want = (M + RT_WARPS - 1) / RT_WARPS;
  const int grid = int(want < (long long)sm_count() * 16 ? want : (long long)sm_count() * 16);
  if (src_is_f32) normalize_rows_kernel<true><<<grid, RT_WARPS * 32, 0, stream>>>(src, dst, M, D);
  else            normalize_rows_kernel<false><<<grid, RT_WARPS * 32, 0, stream>>>(src, dst, M, D);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int retrieval_scan(const bf16* db, const bf16* queries, long long M, int D, int Q, float* scores, cudaStream_t stream) {
  FP_REQUIRE(row_dim_ok(D), "retrieval_scan: D=%d must be a multiple of 256 and <= 1024", D);
  FP_REQUIRE(M >= 0 && Q >= 0, "retrieval_scan: bad shape M=%lld Q=%d", M, Q);
  FP_REQUIRE((reinterpret_cast<uintptr_t>(db) & 15) == 0 && (reinterpret_cast<uintptr_t>(queries) & 15) == 0,
             "retrieval_scan: pointers must be 16-byte aligned");
  if (M == 0 || Q == 0) return 0;
  FP_ENSURE_DYN_SMEM(scan_kernel, MAX_QUERIES * 1024 * 2);
  const long long want = (M + RT_WARPS - 1) / RT_WARPS;
  const int grid = int(want < (long long)sm_count() * 8 ? want : (long long)sm_count() * 8);
  for (int q0 = 0; q0 < Q; q0 += MAX_QUERIES) {  // the database is re-read once per 32 queries
    const int nq = Q - q0 < MAX_QUERIES ? Q - q0 : MAX_QUERIES;
    ProfScope prof(PROF_SCORE, double(M) * D * 2 + double(nq) * M * 4, 1, stream);
    scan_kernel<<<grid, RT_WARPS * 32, size_t(nq) * D * 2, stream>>>(db, queries + size_t(q0) * D, M, D, nq,
                                                                      scores + (long long)q0 * M);
    FP_CUDA(cudaGetLastError());
  }
  return 0;
}

int topk_rows(const float* scores, int Q, long long M, int k, int* idx, float* val, cudaStream_t stream) {
  FP_REQUIRE(k >= 0 && k <= TOPK_MAX, "topk_rows: k=%d exceeds the limit of %d", k, TOPK_MAX);
  FP_REQUIRE(k <= M, "topk_rows: selected index k out of range (k=%d, row length %lld)", k, M);
  FP_REQUIRE(M < (1ll << 31), "topk_rows: row length %lld exceeds 2^31", M);
  if (Q <= 0 || k == 0) return 0;
  ProfScope prof(PROF_SCORE, double(Q) * M * 4 * 5, 1, stream);
  topk_large_kernel<<<Q, 1024, 0, stream>>>(scores, M, k, idx, val);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int retrieval_fine(const bf16* views, const long long* view_start, const int* view_count, int max_views,
                   const int* cand, const bf16* queries, int Q, int C, int D, int k, float* out, cudaStream_t stream) {
  FP_REQUIRE(row_dim_ok(D), "retrieval_fine: D=%d must be a multiple of 256 and <= 1024", D);
  FP_REQUIRE(k >= 1 && k <= FINE_MAX_K, "retrieval_fine: k=%d must be in [1, %d]", k, FINE_MAX_K);
  FP_REQUIRE(max_views <= FINE_MAX_VIEWS, "retrieval_fine: %d views per mesh exceeds the limit of %d", max_views,
             FINE_MAX_VIEWS);
  if (Q <= 0 || C <= 0) return 0;
  ProfScope prof(PROF_SCORE, double(Q) * C * max_views * D * 2, 1, stream);
  fine_kernel<<<dim3(C, Q), RT_WARPS * 32, 0, stream>>>(views, view_start, view_count, cand, queries, C, D, k, out);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int softvote_add(float* acc, const int* idx, const float* val, int P, int C, long long M, cudaStream_t stream) {
  if (P <= 0 || C <= 0) return 0;
  ProfScope prof(PROF_SCORE, double(P) * C * 12, 1, stream);
  softvote_add_kernel<<<(P * C + 255) / 256, 256, 0, stream>>>(acc, idx, val, P, C, M);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int softvote_mean(const float* acc, float* out, long long n, int frames, cudaStream_t stream) {
  FP_REQUIRE(frames > 0, "softvote_mean: no frames");
  if (n <= 0) return 0;
  ProfScope prof(PROF_SCORE, double(n) * 8, 1, stream);
  const long long want = (n + 255) / 256;
  softvote_mean_kernel<<<int(want < 4096 ? want : 4096), 256, 0, stream>>>(acc, out, n, float(frames));
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fp
