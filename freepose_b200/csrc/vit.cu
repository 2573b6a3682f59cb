// DINOv2 ViT-L/14-reg (or ViT-B/14-reg: dim 768, 12 heads, MLP 3072) forward to layer L + final norm (reference src/pipeline/retrieval/dino.py:14-32):
// sequencing of the kernels over caller-provided workspace.  Stateless: weights and buffers are borrowed
// device pointers (include/freepose_b200.h: fp_vit_weights / fp_vit_forward).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "freepose_b200.h"

namespace fp {

namespace {
constexpr int KPAD = 640, NREG = 4;
size_t align256(size_t x) { return (x + 255) / 256 * 256; }
}  // namespace

size_t vit_workspace_bytes(int dim, int mlp_dim, int B, int res) {
  const size_t D = dim > 0 ? dim : 1024, MLP = mlp_dim > 0 ? mlp_dim : 4096;
  const int g = res / 14, P = g * g, T = P + 1 + NREG;
  const size_t M = size_t(B) * T;
  return align256(M * D * 2) * 3        // residual stream x, scratch h (LN1 out / attention out), h2 (LN2 out)
         + align256((M / 128 + 2) * 4)  // fused-LayerNorm row-block counters
         + align256(M * 3 * D * 2)      // qkv
         + align256(M * MLP * 2)        // mlp hidden
         + align256(size_t(B) * P * KPAD * 2)  // patch matrix
         + 256;
}

int vit_forward(const fp_vit_weights* w, const void* input, int input_kind, int B, int res, int layer,
                int feature_type, void* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  FP_REQUIRE(w != nullptr && w->layers != nullptr, "vit: null weights");
  const int D = w->dim > 0 ? w->dim : 1024, HEADS = w->heads > 0 ? w->heads : 16, MLP = w->mlp_dim > 0 ? w->mlp_dim : 4096;
  FP_REQUIRE((D == 1024 || D == 768) && HEADS * 64 == D && MLP % 256 == 0,
             "vit: unsupported model %d / %d heads / MLP %d (ViT-L/14 and ViT-B/14, head dim 64)", D, HEADS, MLP);
  FP_REQUIRE(res > 0 && res % 14 == 0, "vit: crop resolution %d is not a multiple of the 14-pixel patch", res);
  FP_REQUIRE(layer >= 0 && layer <= w->depth, "vit: layer %d outside [0, %d]", layer, w->depth);
  FP_REQUIRE(w->pos_res == res, "vit: position embedding was prepared for %d px crops, got %d", w->pos_res, res);
  FP_REQUIRE(input_kind >= 0 && input_kind <= 2, "vit: unknown input kind %d", input_kind);
  FP_REQUIRE(feature_type >= 0 && feature_type <= 3, "vit: unknown feature type %d", feature_type);
  if (B <= 0) return 0;
  const int g = res / 14, P = g * g, T = P + 1 + NREG;
  const size_t M = size_t(B) * T;
  FP_REQUIRE(M < (size_t(1) << 31) / 4, "vit: batch too large for one call");
  FP_REQUIRE(workspace_bytes >= vit_workspace_bytes(D, MLP, B, res), "vit: workspace too small (%zu < %zu)",
             workspace_bytes, vit_workspace_bytes(D, MLP, B, res));
  FP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "vit: workspace must be 256-byte aligned");

  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  bf16* x = reinterpret_cast<bf16*>(ws); ws += align256(M * size_t(D) * 2);
  bf16* h = reinterpret_cast<bf16*>(ws); ws += align256(M * D * 2);
  bf16* h2 = reinterpret_cast<bf16*>(ws); ws += align256(M * D * 2);
  unsigned* ln_counters = reinterpret_cast<unsigned*>(ws); ws += align256((M / 128 + 2) * 4);
  bf16* qkv = reinterpret_cast<bf16*>(ws); ws += align256(M * 3 * D * 2);
  bf16* mlp = reinterpret_cast<bf16*>(ws); ws += align256(M * MLP * 2);
  bf16* patches = reinterpret_cast<bf16*>(ws);

  // ---- tokens: patch-embed GEMM (+bias, +pos-embed, scattered behind the cls/register rows)
  const bf16* pm = patches;
  if (input_kind == FP_INPUT_PATCHES) {
    pm = reinterpret_cast<const bf16*>(input);
  } else {
    if (int rc = im2col_patches(input, input_kind == FP_INPUT_IMAGE_F32, patches, B, res, KPAD, stream)) return rc;
  }
  if (int rc = write_special_tokens(reinterpret_cast<const bf16*>(w->special_tokens), x, B, T, 1 + NREG, D, stream))
    return rc;
  {
    GemmArgs a{};
    a.A = pm; a.lda = KPAD; a.W = reinterpret_cast<const bf16*>(w->patch_w);
    a.out = x; a.ldo = D; a.M = B * P; a.N = D; a.K = KPAD; a.mode = EPI_PATCH_EMBED;
    a.bias = reinterpret_cast<const bf16*>(w->patch_b);
    a.res = reinterpret_cast<const bf16*>(w->pos_embed);
    a.patches_per_img = P; a.tokens_per_img = T; a.token_offset = 1 + NREG;
    if (int rc = gemm_bf16(a, stream)) return rc;
  }

  // ---- transformer blocks
  // norm2 and the NEXT block's norm1 CAN run inside the residual GEMMs that produce their input (gemm.cu: layernorm_warps)
  // when the batch is large enough for the 2-CTA kernels.  FP_FUSE_LN: bit 0 = norm1 of the next block inside fc2, bit 1 =
  // norm2 inside proj.  Bit-identical tokens either way; OFF by default: measured neutral (DESIGN.md section 6 -- under
  // the power cap the LayerNorm work costs the GEMM as much time inside the kernel as its own launch costs outside).
  const char* fuse_env = getenv("FP_FUSE_LN");
  const int fuse_mask = gemm_fuses_layernorm(int(M), D) ? (fuse_env ? atoi(fuse_env) : 0) : 0;
  const bool fuse_fc2 = (fuse_mask & 1) != 0, fuse_proj = (fuse_mask & 2) != 0;
  unsigned ln_epoch = 0;
  if (fuse_mask && layer > 0) FP_CUDA(cudaMemsetAsync(ln_counters, 0, (M / 128 + 2) * 4, stream));
  const float scale = 0.125f;  // head_dim^-0.5
  for (int l = 0; l < layer; ++l) {
    const fp_vit_layer& L = w->layers[l];
    auto P16 = [](const void* p) { return reinterpret_cast<const bf16*>(p); };
    if (!fuse_fc2 || l == 0)   // (with fusion the previous block's fc2 has already written norm1(x) into h)
      if (int rc = layernorm_bf16(x, P16(L.ln1_w), P16(L.ln1_b), h, int(M), D, 1e-6f, int(M), 0, int(M), stream)) return rc;
    GemmArgs a{};
    a.A = h; a.lda = D; a.W = P16(L.qkv_w); a.out = qkv; a.ldo = 3 * D; a.M = int(M); a.N = 3 * D; a.K = D;
    a.mode = EPI_BIAS; a.bias = P16(L.qkv_b);
    if (int rc = gemm_bf16(a, stream)) return rc;
    if (int rc = attention_bf16(qkv, h, B, T, HEADS, scale, stream)) return rc;
    a = GemmArgs{};
    a.A = h; a.lda = D; a.W = P16(L.proj_w); a.out = x; a.ldo = D; a.M = int(M); a.N = D; a.K = D;
    a.mode = EPI_BIAS_LS_RES; a.bias = P16(L.proj_b); a.gamma = P16(L.ls1); a.res = x;
    if (fuse_proj) {   // norm2 -> h2 (h holds the attention output this GEMM is reading)
      a.ln_w = P16(L.ln2_w); a.ln_b = P16(L.ln2_b); a.ln_out = h2; a.ln_eps = 1e-6f;
      a.ln_counters = ln_counters; a.ln_epoch = ++ln_epoch;
    }
    if (int rc = gemm_bf16(a, stream)) return rc;
    if (!fuse_proj)
      if (int rc = layernorm_bf16(x, P16(L.ln2_w), P16(L.ln2_b), h2, int(M), D, 1e-6f, int(M), 0, int(M), stream)) return rc;
    a = GemmArgs{};
    a.A = h2; a.lda = D; a.W = P16(L.fc1_w); a.out = mlp; a.ldo = MLP; a.M = int(M); a.N = MLP; a.K = D;
    a.mode = EPI_BIAS_GELU; a.bias = P16(L.fc1_b);
    if (int rc = gemm_bf16(a, stream)) return rc;
    a = GemmArgs{};
    a.A = mlp; a.lda = MLP; a.W = P16(L.fc2_w); a.out = x; a.ldo = D; a.M = int(M); a.N = D; a.K = MLP;
    a.mode = EPI_BIAS_LS_RES; a.bias = P16(L.fc2_b); a.gamma = P16(L.ls2); a.res = x;
    if (fuse_fc2 && l + 1 < layer) {   // the next block's norm1 -> h
      const fp_vit_layer& Ln = w->layers[l + 1];
      a.ln_w = P16(Ln.ln1_w); a.ln_b = P16(Ln.ln1_b); a.ln_out = h; a.ln_eps = 1e-6f;
      a.ln_counters = ln_counters; a.ln_epoch = ++ln_epoch;
    }
    if (int rc = gemm_bf16(a, stream)) return rc;
  }

  // ---- final norm, fused with the token slice of dino.py:25-30
  int skip = 0, per = T;
  switch (feature_type) {
    case FP_FEATURE_ALL: skip = 0; per = T; break;
    case FP_FEATURE_CLS: skip = 0; per = 1; break;
    case FP_FEATURE_REG: skip = 1; per = NREG; break;
    case FP_FEATURE_PATCH: skip = 1 + NREG; per = P; break;
  }
  return layernorm_bf16(x, reinterpret_cast<const bf16*>(w->norm_w), reinterpret_cast<const bf16*>(w->norm_b),
                        reinterpret_cast<bf16*>(out), B * per, D, 1e-6f, T, skip, per, stream);
}

}  // namespace fp
