/* freepose_b200 -- C ABI of the B200 (sm_100a) render-and-compare pose engine.
 *
 * Drop-in boundary for FreePose's per-proposal hot path.  The reference has no FFI of its own for this
 * path: its boundary is the Python surface of src/pipeline.  Every entry point below
 * therefore names the reference Python lines whose arithmetic it replaces; the ctypes binding a maintainer
 * adds is shown in INTEGRATION.md and implemented in freepose_b200/_lib.py.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless marked "host";
 *   - buffers are borrowed for the duration of the call, never retained; no hidden allocations: callers
 *     pass workspaces sized by the matching *_workspace_bytes();
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return 0 on success, negative on error (-1 argument/contract violation, -2 CUDA error);
 *     fp_last_error() returns the message for the calling thread;
 *   - bf16 tensors are raw uint16 bit patterns (torch.bfloat16 storage).
 */
#ifndef FREEPOSE_B200_H_
#define FREEPOSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FP_API __attribute__((visibility("default")))
#else
#define FP_API
#endif

#define FP_ABI_VERSION 3

FP_API int fp_abi_version(void);
FP_API const char* fp_last_error(void);
FP_API int fp_device_sm_count(void);

/* Launch accounting and per-kernel-family CUDA-event timing (used by bench.py for gpu_launches and the live
 * roofline numbers; no reference counterpart -- the reference has no timers). */
FP_API long long fp_launch_count(void);            /* kernels launched by this library so far               */
FP_API void fp_profile_enable(int on);             /* bracket every launch with events on its stream        */
FP_API void fp_profile_reset(void);
FP_API int fp_profile_num_kinds(void);
FP_API const char* fp_profile_kind_name(int kind);
/* host out: summed event time (ms), summed algorithmic work (FLOPs for the gemm and attention kinds, bytes otherwise),
 * launches.  Synchronises on the recorded events. */
FP_API int fp_profile_collect(int kind, double* total_ms, double* total_work, long long* launches);

/* ------------------------------------------------------------------------------------------------
 * DINOv2 ViT-{L,B}/14-reg feature extractor (the shapes in the comments are ViT-L's: dim 1024, 16 heads, MLP 4096;
 * ViT-B = 768 / 12 / 3072 is the model of tracking_refiner.py:20-23)
 * replaces: src/pipeline/retrieval/dino.py:14-32  (DINOv2FeatureExtractor.forward:
 *           Normalize -> prepare_tokens_with_masks -> blocks[:layer] -> norm -> token slice)
 *           and hub forward_features()["x_norm_patchtokens"] (tracking_refiner.py:79,84) with layer = depth
 * ------------------------------------------------------------------------------------------------ */
typedef struct fp_vit_layer {          /* all bf16; nn.Linear weights are [out, in] row-major        */
  const void *ln1_w, *ln1_b;           /* [1024]                                                     */
  const void *qkv_w, *qkv_b;           /* [3072, 1024], [3072]                                       */
  const void *proj_w, *proj_b;         /* [1024, 1024], [1024]                                       */
  const void *ls1;                     /* [1024] LayerScale gamma                                    */
  const void *ln2_w, *ln2_b;
  const void *fc1_w, *fc1_b;           /* [4096, 1024], [4096]                                       */
  const void *fc2_w, *fc2_b;           /* [1024, 4096], [1024]                                       */
  const void *ls2;
} fp_vit_layer;

typedef struct fp_vit_weights {
  int depth;                           /* number of entries in `layers` (24 for the full checkpoint) */
  const fp_vit_layer* layers;          /* HOST array of `depth` structs holding device pointers      */
  const void* patch_w;                 /* [1024, 640] bf16: conv weight (1024,3,14,14) flattened to
                                          588 columns (c*196 + ky*14 + kx) and zero padded to 640    */
  const void* patch_b;                 /* [1024]                                                     */
  const void* norm_w;                  /* [1024] final LayerNorm                                     */
  const void* norm_b;
  int pos_res;                         /* crop resolution the two tensors below were prepared for    */
  const void* pos_embed;               /* [1 + g*g, 1024] bf16: bicubic(antialias) resampled
                                          pos_embed for g = pos_res/14 (row 0 = cls position)        */
  const void* special_tokens;          /* [5, 1024] bf16: row 0 = bf16(cls_token + pos_embed[0]),
                                          rows 1..4 = register tokens                                */
  int dim, heads, mlp_dim;             /* ABI 2: 1024/16/4096 (ViT-L) or 768/12/3072 (ViT-B); head dim is 64;
                                          all three 0 = ViT-L                                        */
} fp_vit_weights;

enum { FP_INPUT_IMAGE_F32 = 0,   /* (B,3,res,res) fp32 in [0,1]; bf16 Normalize applied (dino.py:12,16) */
       FP_INPUT_IMAGE_BF16 = 1,  /* (B,3,res,res) bf16, already normalised                              */
       FP_INPUT_PATCHES = 2 };   /* [B*g*g, 640] bf16 normalised patch matrix (fp_crop_resize_pad)      */

enum { FP_FEATURE_ALL = 0,       /* (B, 1+4+g*g, 1024)                                                  */
       FP_FEATURE_CLS = 1,       /* (B, 1024)        dino.py:25-26                                      */
       FP_FEATURE_REG = 2,       /* (B, 4, 1024)     dino.py:27-28                                      */
       FP_FEATURE_PATCH = 3 };   /* (B, g*g, 1024)   dino.py:29-30                                      */

FP_API size_t fp_vit_workspace_bytes(int dim, int mlp_dim, int batch, int res);
FP_API int fp_vit_forward(const fp_vit_weights* weights /* host struct */, const void* input, int input_kind,
                          int batch, int res, int layer, int feature_type, void* out_tokens_bf16,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Stage-level entry points (same kernels; used by the per-stage parity tests and by fp_vit_forward). */
enum { FP_EPI_BIAS = 0, FP_EPI_BIAS_GELU = 1, FP_EPI_BIAS_LS_RES = 2, FP_EPI_PATCH_EMBED = 3 };
/* out[M,N] = epilogue(A[M,K] @ W[N,K]^T): tcgen05 GEMM, fp32 accumulate.  replaces nn.Linear (+GELU /
 * +LayerScale+residual) inside the hub Block and the patch-embed conv (dino.py:16-21). */
FP_API int fp_gemm_bf16(const void* A, int lda, const void* W, void* out, int ldo, int M, int N, int K, int mode,
                        const void* bias, const void* gamma, const void* residual_or_pos, int patches_per_img,
                        int tokens_per_img, int token_offset, void* stream);
/* rows of dim (1024 or 768): y = bf16(((x-mean)*rstd)*w + b); replaces nn.LayerNorm(eps=1e-6) in the hub Block / norm */
FP_API int fp_layernorm_bf16(const void* x, const void* w, const void* b, void* out, int rows, int dim, float eps,
                             int in_group_stride, int in_skip, int rows_per_group, void* stream);
/* qkv [B*T, 3072] -> out [B*T, 1024]; replaces the hub (Mem-Eff)Attention core, 16 heads x 64 */
FP_API int fp_attention_bf16(const void* qkv, void* out, int batch, int tokens, int heads, float scale,
                             void* stream);
FP_API int fp_im2col_patches(const void* image, int src_is_f32, void* patches_bf16, int batch, int res, int kpad,
                             void* stream);
/* replaces torchvision T.Normalize on the bf16 image (dino.py:12,16) */
FP_API int fp_normalize_image(const float* image, void* out_bf16, int batch, int res, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Score + top-k
 * replaces: src/pipeline/estimators/pose_estimator.py:85-92   (normalize, einsum, mean, topk(3))
 *           src/pipeline/estimators/online_pose_estimator.py:68-79 (incl. mask_scores weighting, max/argmax)
 * feats_t (B,P,D) bf16 template tokens, feat_q (P,D) bf16 query tokens, weights (B,P) fp32 or NULL.
 * scores_out (B) fp32: bf16-rounded mean as float (weights NULL) or the fp32 weighted mean.
 * patch_scores_out (B,P) fp32 or NULL.  top-k: descending, ties -> lowest index.
 * Three or four launches on `stream`: query normalisation, per-patch cosines (every template token row is read once),
 * per-hypothesis reduction, top-k.  The workspace holds the normalised query, the top-k scratch and the (B,P) fp32
 * per-patch cosines (when patch_scores_out is NULL); feats_t must be 16-byte aligned, B * P < 2^31.
 * ------------------------------------------------------------------------------------------------ */
FP_API size_t fp_score_workspace_bytes(int B, int P, int D);
FP_API int fp_score_topk(const void* feats_t, const void* feat_q, const float* weights, int B, int P, int D,
                         int normalise_query, float* scores_out, float* patch_scores_out, int k,
                         int32_t* topk_idx, float* topk_val, void* workspace, size_t workspace_bytes,
                         void* stream);
/* top-k over an fp32 score vector already on the device (e.g. after the multi-GPU all-gather) */
FP_API int fp_topk(const float* scores, int B, int k, int32_t* topk_idx, float* topk_val, void* workspace,
                   size_t workspace_bytes, void* stream);
/* replaces: scripts/extract_retrieval_features.py:49-57 (FFA: masked mean of patch tokens).
 * masks (V,res,res) u8, feats (V,g*g,D) bf16 -> out (V,D) fp32, valid (V) int32 = selected patch count */
FP_API int fp_ffa_pool(const void* feats, const uint8_t* masks, int V, int res, int D, float* out, int32_t* valid,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Mesh retrieval
 * replaces: scripts/extract_proposals_ground.py:39-41 (database normalisation), :136-160 (coarse scan, topk(100),
 *           per-view fine re-rank) and scripts/extract_proposals_ground_video.py:148-190 (same + soft vote over frames)
 * All feature rows have D = 256, 512, 768 or 1024 elements.  Top-k order: NaN first, descending, ties -> lowest index.
 * ------------------------------------------------------------------------------------------------ */
/* F.normalize(x.to(bfloat16), dim=-1): src (rows,D) fp32 or bf16 -> dst (rows,D) bf16 */
FP_API int fp_normalize_rows(const void* src, int src_is_f32, int64_t rows, int D, void* dst_bf16, void* stream);
/* scores[q,m] = float(bf16(db[m] . queries[q])): db (M,D) and queries (Q,D) normalised bf16 -> scores (Q,M) fp32.
 * The database is read once for up to 32 queries. */
FP_API int fp_retrieval_scan(const void* db, const void* queries, int64_t M, int D, int Q, float* scores,
                             void* stream);
/* torch.topk(scores[q], k) for every row: scores (Q,M) fp32 -> idx (Q,k) int32, val (Q,k) fp32; k <= 1024 */
FP_API int fp_topk_rows(const float* scores, int Q, int64_t M, int k, int32_t* idx, float* val, void* stream);
/* fine re-rank: views = device-resident per-view features of the meshes, normalised bf16, mesh m owning rows
 * view_start[m] .. view_start[m]+view_count[m]; cand (Q,C) int32 mesh indices (<0: skipped, -inf).  out[q,c] = float32
 * numpy-order mean of topk(views(cand[q,c]) . queries[q], k).  k <= 128, views per mesh <= 4096. */
FP_API int fp_retrieval_fine(const void* views, const int64_t* view_start, const int32_t* view_count, int max_views,
                             const int32_t* cand, const void* queries, int Q, int C, int D, int k, float* out,
                             void* stream);
/* video soft vote: acc (P,M) fp32 += scatter of one frame's (idx, val) (P,C); then mean over `frames` */
FP_API int fp_softvote_add(float* acc, const int32_t* idx, const float* val, int P, int C, int64_t M, void* stream);
FP_API int fp_softvote_mean(const float* acc, float* out, int64_t n, int frames, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Rasteriser
 * replaces: src/pipeline/retrieval/renderer.py:43-95 (MeshRenderer.render / render_from_poses: one pyrender
 *           GL draw + glReadPixels per pose)
 * ------------------------------------------------------------------------------------------------ */
typedef struct fp_raster_args {
  const float* verts;       /* [V,3] fp32 mesh vertices (object frame, already at rendering scale)      */
  const int32_t* faces;     /* [F,3]                                                                     */
  const uint8_t* colors;    /* [V,3] u8 vertex colours                                                   */
  int V, F;
  const float* poses;       /* [B,12] row-major 3x4 object->camera (OpenCV frame), rows of the 4x4 pose  */
  int B;
  float fx, fy, cx, cy;     /* pyrender.IntrinsicsCamera(fx, fy, cx, cy), renderer.py:37                 */
  int res;
  int msaa;                 /* 4 = pyrender's multisampled offscreen target, 1 = centre sampling         */
  int cull_backfaces;       /* 0 = RenderFlags.SKIP_CULL_FACES (reference default, renderer.py:63-66)    */
  const uint8_t* gamma_lut; /* [65536] u8: round(255 * (i/65535)^(1/2.2))                                */
  uint8_t* rgb;             /* out [B,res,res,3] u8                                                      */
  float* depth;             /* out [B,res,res] fp32 metres, 0 = background                               */
  /* ABI 2: the other two inputs renderer.py:43-51,70-78 accepts */
  int primitive;            /* 0 = triangles; 1 = points (trimesh.PointCloud -> pyrender.Mesh.from_points: one
                               1-pixel sprite per vertex, flat vertex colour; faces/F ignored)                */
  const float* uv;          /* [V,2] fp32 texture coordinates, v up (trimesh TextureVisuals.uv), or NULL    */
  const uint8_t* texture;   /* RGBA8 mip chain (level 0 first; level l is max(1,w>>l) x max(1,h>>l), rows
                               top-down), or NULL = vertex colours only; colors may be NULL with a texture   */
  int tex_w, tex_h, tex_levels;
  const float* srgb_lut;    /* [65536] fp32: (i/65535)^2.2, pyrender's srgb_to_linear applied after filtering */
  /* the refiner's render set-up (tracking_refiner.py:31-45): ambient 5, znear 1e-4 / zfar 9999, a camera per frame */
  float ambient;            /* scene ambient light; 0 = 2.0 (renderer.py:53-55)                             */
  float znear, zfar;        /* 0 = pyrender's IntrinsicsCamera defaults 0.05 / 100                           */
  const float* view_k;      /* [B,4] fp32 per-view fx,fy,cx,cy (overrides the four scalars) or NULL          */
} fp_raster_args;
FP_API int fp_raster_workspace_bytes(int B, int V, int F, int res, int msaa, size_t* bytes /* host out */);
FP_API int fp_rasterize(const fp_raster_args* args /* host struct */, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Geometry around the renders
 * ------------------------------------------------------------------------------------------------ */
/* replaces: renderer.py:98-117 (mask = depth > 0; < min_count px -> mask[lo:hi, lo:hi] = True; mask_to_bbox).
 * bbox_out (B,4) int32 xmin,ymin,xmax,ymax; count_out (B) or NULL; mask_out (B,res,res) u8 or NULL */
FP_API int fp_mask_bbox(const float* depth, int B, int res, int fallback_lo, int fallback_hi, int min_count,
                        int32_t* bbox_out, int32_t* count_out, uint8_t* mask_out, void* stream);
/* replaces: src/utils/bbox_utils.py:20-56 (CropResizePad.__call__) fused with renderer.py:119-129 and, for the
 * patch-matrix output, with dino.py:12,16 (Normalize) and the patch-embed im2col.
 * boxes (B,4) int32 x1,y1,x2,y2 already extended/clamped (slice semantics: x2,y2 exclusive).
 * src: u8 HWC (B,src_h,src_w,3) or fp32 CHW (B,3,src_h,src_w).  dst: fp32 CHW (B,3,T,T) or the bf16 patch
 * matrix [B*(T/14)^2, kpad] (needs norm_lut [3*256] bf16 = Normalize(bf16(v/255)) per channel).
 * status: int32 device word, set to 1+index of a box the reference would have failed on (else untouched). */
FP_API int fp_crop_resize_pad(const void* src, int src_is_u8_hwc, const int32_t* boxes, const void* norm_lut,
                              void* dst, int dst_is_patches, int B, int src_h, int src_w, int T, int kpad,
                              int32_t* status, void* stream);
/* replaces: src/pipeline/utils.py:122-145 (depthmap_to_pointcloud) reduced to what utils.py:148-170
 * (get_z_from_pointcloud) and pose_estimator.py:103-111 consume.  view_idx (n) int32 or NULL selects views;
 * kinv (9) fp64 device = inv(K) row-major; out (n,8) fp64: xmin,xmax,ymin,ymax,sum_x,sum_y,sum_z,count */
FP_API int fp_depth_extents(const float* depth, const int32_t* view_idx, int n, int res, const double* kinv,
                            double* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Refiner pose-confidence pass: tracking_refiner.py:70-100 = roi_align crop of the photo
 * at 518^2, one render at the cropped K (fp_rasterize with ambient 5, view_k), ViT-B/14-reg on both
 * (fp_vit_forward, layer = depth), masked per-patch cosine.
 * ------------------------------------------------------------------------------------------------ */
/* replaces: torchvision.ops.roi_align(image[None], boxes, (out_h,out_w), sampling_ratio=2) in
 * src/pipeline/refiner_utils.py:128-133.  image (C,H,W) fp32, boxes (n,4) fp32 x1,y1,x2,y2 on the device,
 * out (n,C,out_h,out_w) fp32; spatial_scale 1, aligned=False. */
FP_API int fp_roi_align(const float* image, int channels, int height, int width, const float* boxes, int n, int out_h,
                        int out_w, int sampling_ratio, float* out, void* stream);
/* replaces: cv2.resize((depth > 0).astype(float32), (g,g), INTER_CUBIC) > 0.5, tracking_refiner.py:75.
 * depth (B,src_stride,src_stride) fp32 of which the top-left res x res is the image (the rasteriser renders 518 px
 * views into 520 px targets) -> mask_out (B,g,g) u8 */
FP_API int fp_depth_mask_cubic(const float* depth, int B, int res, int src_stride, int g, uint8_t* mask_out,
                               void* stream);
/* replaces: tracking_refiner.py:80-88.  feats_a, feats_b (rows, dim) bf16 token rows, mask (rows) u8 or NULL ->
 * out (rows) fp32 = mask * cos(a, b) */
FP_API int fp_patch_cosine(const void* feats_a, const void* feats_b, const uint8_t* mask, int rows, int dim, float* out,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU; the hypotheses of a proposal are split contiguously over the
 * ranks and the ONLY exchange is an all-gather of the per-hypothesis fp32 scores, after which every rank runs the same
 * deterministic fp_topk.  No reference counterpart: the reference scales with SLURM array jobs + CSV concatenation
 * (scripts/dino_inference.py:37-40, scripts/merge_results.py:14-29).
 * NCCL is resolved with dlopen at the first call (the libnccl.so.2 PyTorch ships); the id is 128 opaque bytes made
 * by rank 0 and handed to the other ranks by the host (torch.distributed / MPI / a file).
 * ------------------------------------------------------------------------------------------------ */
typedef struct fp_comm_id { char bytes[128]; } fp_comm_id;
FP_API int fp_comm_unique_id(fp_comm_id* id /* host out */);
FP_API int fp_comm_create(const fp_comm_id* id /* host */, int rank, int world, void** comm /* host out */);
/* scores: (world * per_rank) fp32 on the device; this rank's slice [rank*per_rank, +per_rank) was written by
 * fp_score_topk (scores_out pointing into the buffer).  In-place all-gather enqueued on `stream`. */
FP_API int fp_allgather_scores(void* comm, float* scores, int per_rank, void* stream);
FP_API int fp_comm_destroy(void* comm);

/* The same exchange WITHOUT a collective call: peer memory over NVLink (CUDA IPC; one process per GPU on one node).
 * Every rank owns an exchange buffer of fp_exchange_bytes(world, per_rank) that it allocates with fp_p2p_alloc and whose
 * 64-byte handle the host hands to the other ranks, which map it with fp_p2p_open.  fp_score_publish is fp_score_topk's
 * score stage with the all-gather fused in: each score is stored into slot (rank, b) of EVERY rank's buffer as it is
 * produced, and the last CTA raises this rank's flag in every buffer.  fp_topk_after_exchange waits (on the device) for
 * the world flags of the own buffer and runs the deterministic top-k.  `epoch` must increase by one per exchange on all
 * ranks (its parity selects one of two halves of the buffer, so a rank may run one exchange ahead of a peer).
 * peers: DEVICE array of `world` pointers (own buffer included, as mapped in this process); publish_workspace:
 * fp_score_workspace_bytes(B,P,D) + 256 bytes, kept by the caller across calls and zeroed once. */
typedef struct fp_p2p_handle { char bytes[64]; } fp_p2p_handle;
FP_API size_t fp_exchange_bytes(int world, int per_rank);
FP_API int fp_p2p_alloc(size_t bytes, void** ptr /* host out */, fp_p2p_handle* handle /* host out */);
FP_API int fp_p2p_open(const fp_p2p_handle* handle /* host */, void** ptr /* host out */);
FP_API int fp_p2p_close(void* ptr);
FP_API int fp_p2p_free(void* ptr);
FP_API int fp_score_publish(const void* feats_t, const void* feat_q, const float* weights, int B, int P, int D,
                            int normalise_query, void* const* peers, void* own_buffer, int rank, int world, int per_rank,
                            unsigned epoch, void* publish_workspace, size_t workspace_bytes, void* stream);
FP_API int fp_topk_after_exchange(void* own_buffer, int world, int per_rank, int n_total, unsigned epoch, int k,
                                  int32_t* topk_idx, float* topk_val, void* workspace, size_t workspace_bytes,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FREEPOSE_B200_H_ */
