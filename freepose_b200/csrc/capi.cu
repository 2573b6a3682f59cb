// extern "C" surface of libfreepose_b200.so (include/freepose_b200.h).  Thin: validates, forwards to the
// kernels' launchers, never throws, never allocates.
#include "freepose_b200.h"

#include "common.cuh"
#include "kernels.h"

using fp::bf16;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline const bf16* B16(const void* p) { return reinterpret_cast<const bf16*>(p); }
static inline bf16* B16(void* p) { return reinterpret_cast<bf16*>(p); }

extern "C" {

FP_API int fp_abi_version(void) { return FP_ABI_VERSION; }
FP_API const char* fp_last_error(void) { return fp::last_error(); }
FP_API int fp_device_sm_count(void) { return fp::sm_count(); }
FP_API long long fp_launch_count(void) { return fp::launch_count(); }
FP_API void fp_profile_enable(int on) { fp::prof_enable(on); }
FP_API void fp_profile_reset(void) { fp::prof_reset(); }
FP_API int fp_profile_num_kinds(void) { return fp::PROF_NUM_KINDS; }
FP_API const char* fp_profile_kind_name(int kind) { return fp::prof_name(kind); }
FP_API int fp_profile_collect(int kind, double* total_ms, double* total_work, long long* launches) {
  return fp::prof_collect(kind, total_ms, total_work, launches);
}

FP_API size_t fp_vit_workspace_bytes(int dim, int mlp_dim, int batch, int res) {
  return fp::vit_workspace_bytes(dim, mlp_dim, batch, res);
}

FP_API int fp_vit_forward(const fp_vit_weights* weights, const void* input, int input_kind, int batch, int res,
                          int layer, int feature_type, void* out_tokens_bf16, void* workspace,
                          size_t workspace_bytes, void* stream) {
  return fp::vit_forward(weights, input, input_kind, batch, res, layer, feature_type, out_tokens_bf16, workspace,
                         workspace_bytes, S(stream));
}

FP_API int fp_gemm_bf16(const void* A, int lda, const void* W, void* out, int ldo, int M, int N, int K, int mode,
                        const void* bias, const void* gamma, const void* residual_or_pos, int patches_per_img,
                        int tokens_per_img, int token_offset, void* stream) {
  fp::GemmArgs a{};
  a.A = B16(A); a.lda = lda; a.W = B16(W); a.out = B16(out); a.ldo = ldo;
  a.M = M; a.N = N; a.K = K; a.mode = mode;
  a.bias = B16(bias); a.gamma = B16(gamma); a.res = B16(residual_or_pos);
  a.patches_per_img = patches_per_img; a.tokens_per_img = tokens_per_img; a.token_offset = token_offset;
  return fp::gemm_bf16(a, S(stream));
}

FP_API int fp_layernorm_bf16(const void* x, const void* w, const void* b, void* out, int rows, int dim, float eps,
                             int in_group_stride, int in_skip, int rows_per_group, void* stream) {
  return fp::layernorm_bf16(B16(x), B16(w), B16(b), B16(out), rows, dim, eps, in_group_stride, in_skip,
                            rows_per_group, S(stream));
}

FP_API int fp_attention_bf16(const void* qkv, void* out, int batch, int tokens, int heads, float scale,
                             void* stream) {
  return fp::attention_bf16(B16(qkv), B16(out), batch, tokens, heads, scale, S(stream));
}

FP_API int fp_im2col_patches(const void* image, int src_is_f32, void* patches_bf16, int batch, int res, int kpad,
                             void* stream) {
  return fp::im2col_patches(image, src_is_f32, B16(patches_bf16), batch, res, kpad, S(stream));
}

FP_API int fp_normalize_image(const float* image, void* out_bf16, int batch, int res, void* stream) {
  return fp::normalize_image(image, B16(out_bf16), batch, res, S(stream));
}

FP_API size_t fp_score_workspace_bytes(int B, int P, int D) { return fp::score_workspace_bytes(B, P, D); }

FP_API int fp_score_topk(const void* feats_t, const void* feat_q, const float* weights, int B, int P, int D,
                         int normalise_query, float* scores_out, float* patch_scores_out, int k,
                         int32_t* topk_idx, float* topk_val, void* workspace, size_t workspace_bytes,
                         void* stream) {
  return fp::score_topk(B16(feats_t), B16(feat_q), weights, B, P, D, normalise_query, scores_out,
                        patch_scores_out, k, topk_idx, topk_val, workspace, workspace_bytes, S(stream));
}

FP_API int fp_topk(const float* scores, int B, int k, int32_t* topk_idx, float* topk_val, void* workspace,
                   size_t workspace_bytes, void* stream) {
  return fp::topk_only(scores, B, k, topk_idx, topk_val, workspace, workspace_bytes, S(stream));
}

FP_API int fp_ffa_pool(const void* feats, const uint8_t* masks, int V, int res, int D, float* out, int32_t* valid,
                       void* stream) {
  if (res <= 0 || res % 14 != 0) {
    fp::set_error("ffa: mask resolution %d is not a multiple of 14", res);
    return -1;
  }
  return fp::ffa_pool(B16(feats), masks, V, res, res / 14, D, out, valid, S(stream));
}

FP_API int fp_normalize_rows(const void* src, int src_is_f32, int64_t rows, int D, void* dst_bf16, void* stream) {
  return fp::normalize_rows(src, src_is_f32, rows, D, B16(dst_bf16), S(stream));
}

FP_API int fp_retrieval_scan(const void* db, const void* queries, int64_t M, int D, int Q, float* scores,
                             void* stream) {
  return fp::retrieval_scan(B16(db), B16(queries), M, D, Q, scores, S(stream));
}

FP_API int fp_topk_rows(const float* scores, int Q, int64_t M, int k, int32_t* idx, float* val, void* stream) {
  return fp::topk_rows(scores, Q, M, k, idx, val, S(stream));
}

FP_API int fp_retrieval_fine(const void* views, const int64_t* view_start, const int32_t* view_count, int max_views,
                             const int32_t* cand, const void* queries, int Q, int C, int D, int k, float* out,
                             void* stream) {
  return fp::retrieval_fine(B16(views), reinterpret_cast<const long long*>(view_start), view_count, max_views, cand,
                            B16(queries), Q, C, D, k, out, S(stream));
}

FP_API int fp_softvote_add(float* acc, const int32_t* idx, const float* val, int P, int C, int64_t M, void* stream) {
  return fp::softvote_add(acc, idx, val, P, C, M, S(stream));
}

FP_API int fp_softvote_mean(const float* acc, float* out, int64_t n, int frames, void* stream) {
  return fp::softvote_mean(acc, out, n, frames, S(stream));
}

FP_API int fp_raster_workspace_bytes(int B, int V, int F, int res, int msaa, size_t* bytes) {
  return fp::raster_workspace_bytes(B, V, F, res, msaa, bytes);
}

FP_API int fp_rasterize(const fp_raster_args* g, void* workspace, size_t workspace_bytes, void* stream) {
  if (g == nullptr) {
    fp::set_error("raster: null args");
    return -1;
  }
  fp::RasterArgs a;
  a.verts = g->verts; a.faces = g->faces; a.colors = g->colors; a.V = g->V; a.F = g->F;
  a.poses = g->poses; a.B = g->B; a.fx = g->fx; a.fy = g->fy; a.cx = g->cx; a.cy = g->cy;
  a.res = g->res; a.msaa = g->msaa; a.cull_backfaces = g->cull_backfaces; a.gamma_lut = g->gamma_lut;
  a.rgb = g->rgb; a.depth = g->depth;
  a.primitive = g->primitive; a.uv = g->uv; a.texture = g->texture;
  a.tex_w = g->tex_w; a.tex_h = g->tex_h; a.tex_levels = g->tex_levels; a.srgb_lut = g->srgb_lut;
  a.ambient = g->ambient; a.znear = g->znear; a.zfar = g->zfar; a.view_k = g->view_k;
  return fp::rasterize(a, workspace, workspace_bytes, S(stream));
}

FP_API int fp_mask_bbox(const float* depth, int B, int res, int fallback_lo, int fallback_hi, int min_count,
                        int32_t* bbox_out, int32_t* count_out, uint8_t* mask_out, void* stream) {
  return fp::mask_bbox(depth, B, res, fallback_lo, fallback_hi, min_count, bbox_out, count_out, mask_out, S(stream));
}

FP_API int fp_crop_resize_pad(const void* src, int src_is_u8_hwc, const int32_t* boxes, const void* norm_lut,
                              void* dst, int dst_is_patches, int B, int src_h, int src_w, int T, int kpad,
                              int32_t* status, void* stream) {
  return fp::crop_resize_pad(src, src_is_u8_hwc, boxes, B16(norm_lut), dst, dst_is_patches, B, src_h, src_w, T, kpad,
                             status, S(stream));
}

FP_API int fp_depth_extents(const float* depth, const int32_t* view_idx, int n, int res, const double* kinv,
                            double* out, void* stream) {
  return fp::depth_extents(depth, view_idx, n, res, kinv, out, S(stream));
}

FP_API int fp_roi_align(const float* image, int channels, int height, int width, const float* boxes, int n, int out_h,
                        int out_w, int sampling_ratio, float* out, void* stream) {
  return fp::roi_align(image, channels, height, width, boxes, n, out_h, out_w, sampling_ratio, out, S(stream));
}

FP_API int fp_depth_mask_cubic(const float* depth, int B, int res, int src_stride, int g, uint8_t* mask_out,
                               void* stream) {
  return fp::depth_mask_cubic(depth, B, res, src_stride, g, mask_out, S(stream));
}

FP_API int fp_patch_cosine(const void* feats_a, const void* feats_b, const uint8_t* mask, int rows, int dim, float* out,
                           void* stream) {
  return fp::patch_cosine(B16(feats_a), B16(feats_b), mask, rows, dim, out, S(stream));
}

FP_API int fp_comm_unique_id(fp_comm_id* id) { return fp::comm_unique_id(id); }
FP_API int fp_comm_create(const fp_comm_id* id, int rank, int world, void** comm) {
  return fp::comm_create(id, rank, world, comm);
}
FP_API int fp_allgather_scores(void* comm, float* scores, int per_rank, void* stream) {
  return fp::comm_allgather_scores(comm, scores, per_rank, S(stream));
}
FP_API int fp_comm_destroy(void* comm) { return fp::comm_destroy(comm); }

FP_API size_t fp_exchange_bytes(int world, int per_rank) { return fp::exchange_bytes(world, per_rank); }
FP_API int fp_p2p_alloc(size_t bytes, void** ptr, fp_p2p_handle* handle) { return fp::p2p_alloc(bytes, ptr, handle); }
FP_API int fp_p2p_open(const fp_p2p_handle* handle, void** ptr) { return fp::p2p_open(handle, ptr); }
FP_API int fp_p2p_close(void* ptr) { return fp::p2p_close(ptr); }
FP_API int fp_p2p_free(void* ptr) { return fp::p2p_free(ptr); }
FP_API int fp_score_publish(const void* feats_t, const void* feat_q, const float* weights, int B, int P, int D,
                            int normalise_query, void* const* peers, void* own_buffer, int rank, int world, int per_rank,
                            unsigned epoch, void* publish_workspace, size_t workspace_bytes, void* stream) {
  return fp::score_publish(B16(feats_t), B16(feat_q), weights, B, P, D, normalise_query,
                           reinterpret_cast<float* const*>(peers), reinterpret_cast<float*>(own_buffer), rank, world, per_rank,
                           epoch, publish_workspace, workspace_bytes, S(stream));
}
FP_API int fp_topk_after_exchange(void* own_buffer, int world, int per_rank, int n_total, unsigned epoch, int k,
                                  int32_t* topk_idx, float* topk_val, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  return fp::topk_after_exchange(reinterpret_cast<float*>(own_buffer), world, per_rank, n_total, epoch, k, topk_idx, topk_val,
                                 workspace, workspace_bytes, S(stream));
}

}  // extern "C"
