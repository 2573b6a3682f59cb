// Internal C++ interface between the C ABI (capi.cu), the ViT engine (vit.cu) and the kernels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fp {

typedef __nv_bfloat16 bf16;

const char* last_error();
void prof_enable(int on);
void prof_reset();
const char* prof_name(int kind);
long long launch_count();
int prof_collect(int kind, double* total_ms, double* total_work, long long* launches);

// ---------------------------------------------------------------------------------------- GEMM
enum EpilogueMode { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_LS_RES = 2, EPI_PATCH_EMBED = 3 };

struct GemmArgs {
  const bf16* A;   // [M, lda], K contiguous
  int lda;
  const bf16* W;   // [N, K], K contiguous (nn.Linear weight layout)
  bf16* out;       // [*, ldo]
  int ldo;
  int M, N, K;
  int mode;
  const bf16* bias;   // [N]
  const bf16* gamma;  // [N]             (EPI_BIAS_LS_RES)
  const bf16* res;    // [M, ldo] residual (EPI_BIAS_LS_RES, may alias out) | pos-embed [1+P, N] (EPI_PATCH_EMBED)
  int patches_per_img, tokens_per_img, token_offset;  // EPI_PATCH_EMBED row remap
  // EPI_BIAS_LS_RES only, optional (ln_out != nullptr): LayerNorm of the finished rows of `out` by extra warps of the
  // same kernel -> ln_out [M, N].  ln_counters: ceil(M / 128) words, zeroed once by the caller and shared by consecutive
  // fused launches; ln_epoch = 1, 2, ... counts those launches (the kernel expects counter >= epoch * arrivals).
  const bf16* ln_w = nullptr;
  const bf16* ln_b = nullptr;
  bf16* ln_out = nullptr;
  float ln_eps = 0.f;
  unsigned* ln_counters = nullptr;
  unsigned ln_epoch = 0;
};
int gemm_bf16(const GemmArgs& a, cudaStream_t stream);
bool gemm_fuses_layernorm(int M, int N);   // whether gemm_bf16 accepts ln_out for this problem (large M, N = 1024 | 768)

// ---------------------------------------------------------------------------------------- attention
// qkv: [B*T, 3*H*64] bf16 (q | k | v, head-major inside each), out: [B*T, H*64] bf16.
// T <= 272 (crops up to 224^2): all keys of a head resident in TMEM, one softmax pass.  Above that the call is routed
// to attention_long_bf16: key blocks of 256 with an online softmax (oracle: contract_attention(key_block=256)).
int attention_bf16(const bf16* qkv, bf16* out, int B, int T, int H, float scale, cudaStream_t stream);
int attention_long_bf16(const bf16* qkv, bf16* out, int B, int T, int H, float scale, cudaStream_t stream);
// T > 272: two query tiles in flight, key blocks of 96, double-buffered logits (attention_pair.cu; oracle:
// contract_attention(key_block=96)).  attention_long_bf16 is its predecessor (key blocks of 256), kept for A/B runs.
int attention_pair_bf16(const bf16* qkv, bf16* out, int B, int T, int H, float scale, cudaStream_t stream);
// T == 261 (224^2 crops): two independent key streams per query tile (attention_split.cu)
int attention_split_bf16(const bf16* qkv, bf16* out, int B, int T, int H, float scale, unsigned poly_mask,
                         cudaStream_t stream);

// ---------------------------------------------------------------------------------------- elementwise
// y = bf16(((x - mean) * rstd) * w + b), fp32 two-pass statistics, rows of D = 1024.
// Row r of the output is written at out + (r / rows_per_group * out_group_stride + r % rows_per_group) * D
// after skipping `in_skip` leading rows of every `in_group_stride`-row input group (used by the
// final norm to drop cls + register tokens while normalising).
int layernorm_bf16(const bf16* x, const bf16* w, const bf16* b, bf16* out, int rows, int D, float eps,
                   int in_group_stride, int in_skip, int rows_per_group, cudaStream_t stream);

// Image (B, 3, res, res) -> patch matrix [B*g*g, Kpad] (col = c*196 + ky*14 + kx, zero padded).
// src_is_f32 = 1: fp32 [0,1] image, the reference's bf16 Normalize is applied on the fly;
// src_is_f32 = 0: already-normalised bf16 image.
int im2col_patches(const void* img, int src_is_f32, bf16* out, int B, int res, int Kpad, cudaStream_t stream);

// Float image in [0,1] (B,3,res,res) fp32 -> reference preprocessing in bf16:
// x_bf16 = bf16(x); y = bf16(bf16(x_bf16 - bf16(mean_c)) / bf16(std_c))  (reference dino.py:12,16)
int normalize_image(const float* img, bf16* out, int B, int res, cudaStream_t stream);

// Rows 0..R of every image's token block: special[(1+R), D] (cls+pos[0], registers) broadcast.
int write_special_tokens(const bf16* special, bf16* tokens, int B, int T, int n_special, int D, cudaStream_t stream);

// ---------------------------------------------------------------------------------------- score
// Position-aligned per-patch cosine (reference pose_estimator.py:85-90).  See score.cu.
int score_topk(const bf16* feats_t, const bf16* feat_q, const float* weights, int B, int P, int D,
               int normalise_query, float* scores_out, float* patch_scores_out, int k, int* topk_idx,
               float* topk_val, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t score_workspace_bytes(int B, int P, int D);
int topk_only(const float* scores, int B, int k, int* topk_idx, float* topk_val, void* workspace,
              size_t workspace_bytes, cudaStream_t stream);
// peer-memory exchange of the scores (score.cu: PeerExchange)
size_t exchange_bytes(int world, int per_rank);
int score_publish(const bf16* feats_t, const bf16* feat_q, const float* weights, int B, int P, int D, int normalise_query,
                  float* const* peers_dev, float* own_buffer, int rank, int world, int per_rank, unsigned epoch,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream);
int topk_after_exchange(float* own_buffer, int world, int per_rank, int n_total, unsigned epoch, int k, int* topk_idx,
                        float* topk_val, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int p2p_alloc(size_t bytes, void** ptr, void* handle64);
int p2p_open(const void* handle64, void** ptr);
int p2p_close(void* ptr);
int p2p_free(void* ptr);

// FFA pooling (reference extract_retrieval_features.py:49-57)
int ffa_pool(const bf16* feats, const uint8_t* masks, int V, int res, int g, int D, float* out,
             int* valid, cudaStream_t stream);

// ---------------------------------------------------------------------------------------- retrieval (retrieval.cu)
// F.normalize(x.to(bf16), dim=-1) row-wise; src fp32 or bf16 [M, D] -> dst bf16 [M, D].
int normalize_rows(const void* src, int src_is_f32, long long M, int D, bf16* dst, cudaStream_t stream);
// scores[q, m] = bf16(db[m] . queries[q]) as fp32; db [M, D], queries [Q, D], both bf16 and already normalised.
int retrieval_scan(const bf16* db, const bf16* queries, long long M, int D, int Q, float* scores, cudaStream_t stream);
// per row of scores [Q, M]: top-k (k <= 1024) descending, NaN first, ties -> lowest index; idx/val [Q, k].
int topk_rows(const float* scores, int Q, long long M, int k, int* idx, float* val, cudaStream_t stream);
// out[q, c] = float32 mean of the top-k per-view scores of candidate mesh cand[q, c] (views of mesh m are rows
// view_start[m] .. +view_count[m] of `views`, normalised bf16); cand < 0 -> -inf.
int retrieval_fine(const bf16* views, const long long* view_start, const int* view_count, int max_views,
                   const int* cand, const bf16* queries, int Q, int C, int D, int k, float* out, cudaStream_t stream);
int softvote_add(float* acc, const int* idx, const float* val, int P, int C, long long M, cudaStream_t stream);
int softvote_mean(const float* acc, float* out, long long n, int frames, cudaStream_t stream);

// ---------------------------------------------------------------------------------------- raster
struct RasterArgs {
  const float* verts;     // [V, 3] fp32 object-space
  const int32_t* faces;   // [F, 3]
  const uint8_t* colors;  // [V, 3] u8 vertex colours (may be null when a texture is given)
  int V, F;
  const float* poses;     // [B, 12] row-major 3x4 (R | t), object -> OpenCV camera
  int B;
  float fx, fy, cx, cy;
  int res;
  int msaa;               // 1 or 4
  int cull_backfaces;
  const uint8_t* gamma_lut;  // [65536] u8
  uint8_t* rgb;           // [B, res, res, 3] u8
  float* depth;           // [B, res, res] fp32
  int primitive;          // 0 = triangles, 1 = points (faces ignored)
  const float* uv;        // [V, 2] fp32 or null
  const uint8_t* texture; // RGBA8 mip chain or null
  int tex_w, tex_h, tex_levels;
  const float* srgb_lut;  // [65536] fp32
  float ambient;          // 0 = 2.0
  float znear, zfar;      // 0 = 0.05 / 100
  const float* view_k;    // [B, 4] per-view fx, fy, cx, cy or null
};
int raster_workspace_bytes(int B, int V, int F, int res, int msaa, size_t* bytes);
int rasterize(const RasterArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t stream);

// ---------------------------------------------------------------------------------------- geometry
int mask_bbox(const float* depth, int B, int res, int fallback_lo, int fallback_hi, int min_count,
              int32_t* bbox_out, int32_t* count_out, uint8_t* mask_out, cudaStream_t stream);
int crop_resize_pad(const void* src, int src_is_u8_hwc, const int32_t* boxes, const bf16* norm_lut, void* dst,
                    int dst_is_patches, int B, int src_h, int src_w, int T, int Kpad, int32_t* status,
                    cudaStream_t stream);
int depth_extents(const float* depth, const int32_t* view_idx, int n_out, int res, const double* kinv_dev,
                  double* out, cudaStream_t stream);

// ---------------------------------------------------------------------------------------- refiner confidence pass
int roi_align(const float* image, int C, int H, int W, const float* boxes, int n, int out_h, int out_w,
              int sampling_ratio, float* out, cudaStream_t stream);
int depth_mask_cubic(const float* depth, int B, int res, int stride, int g, uint8_t* mask, cudaStream_t stream);
int patch_cosine(const bf16* a, const bf16* b, const uint8_t* mask, int rows, int D, float* out, cudaStream_t stream);

}  // namespace fp

// ---------------------------------------------------------------------------------------- ViT
struct fp_vit_weights;
struct fp_comm_id;
namespace fp {
int comm_unique_id(fp_comm_id* out);
int comm_create(const fp_comm_id* id, int rank, int world, void** comm_out);
int comm_allgather_scores(void* comm, float* scores, int per_rank, cudaStream_t stream);
int comm_destroy(void* comm);
}  // namespace fp
namespace fp {
size_t vit_workspace_bytes(int dim, int mlp_dim, int B, int res);
int vit_forward(const fp_vit_weights* w, const void* input, int input_kind, int B, int res, int layer,
                int feature_type, void* out, void* workspace, size_t workspace_bytes, cudaStream_t stream);
}  // namespace fp
