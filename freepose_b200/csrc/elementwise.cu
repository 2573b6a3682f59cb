// HBM-bound helper kernels of the ViT path: LayerNorm, image normalisation + patch gather (im2col of a
// stride-14 conv is a pure permutation), special-token broadcast.  All use 16-byte vector accesses.
#include "common.cuh"
#include "kernels.h"
#include "rowops.cuh"
#include "ln_row.cuh"

namespace fp {

namespace {

// ------------------------------------------------------------------------------------------------
// LayerNorm over rows of D = 256 * NCH (ViT-L: 1024, ViT-B: 768): one warp per row, NCH x 16-byte loads per lane,
// fp32 two-pass stats.  Contract (oracle/vit.py contract_layernorm): y = bf16(((x - mean) * rstd) * w + b).
//
// HBM-bound on paper (4 KB per row), but inside a step the SM clock sits at the GEMMs' power-capped ~1.3 GHz and the
// first version's ~380 instructions per row (scalar adds, gamma / beta unpacked again for every row) made it issue-bound
// there: 4.7-5.0 TB/s in-step against 6.3 TB/s standalone.  Now: persistent warps (two CTAs per SM, rows dealt
// round-robin) keep gamma and beta unpacked in registers as fp32 pairs, every element-wise step is a packed fp32x2
// instruction (FADD2 / FFMA2 / FMUL2: the same IEEE operation per lane), the sums run in two independent pair
// accumulators, and the next row is in flight while the current one is normalised: ~170 instructions per row.
// ------------------------------------------------------------------------------------------------
constexpr int LN_WARPS = 8;
using rowops::f32x2;

template <int NCH>
__global__ void __launch_bounds__(LN_WARPS * 32, 2)
layernorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                 bf16* __restrict__ out, int rows, float eps, int in_group_stride, int in_skip,
                 int rows_per_group) {
  using namespace rowops;
  constexpr int LN_D = NCH * 256;
  constexpr int NP = NCH * 4;   // fp32 pairs per lane
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = gridDim.x * LN_WARPS;
  int r = blockIdx.x * LN_WARPS + warp;
  if (r >= rows) return;
  f32x2 w2[NP], b2[NP];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const uint4 wu = __ldg(reinterpret_cast<const uint4*>(w) + c * 32 + lane);
    const uint4 bu = __ldg(reinterpret_cast<const uint4*>(b) + c * 32 + lane);
    w2[c * 4 + 0] = word2(wu.x); w2[c * 4 + 1] = word2(wu.y); w2[c * 4 + 2] = word2(wu.z); w2[c * 4 + 3] = word2(wu.w);
    b2[c * 4 + 0] = word2(bu.x); b2[c * 4 + 1] = word2(bu.y); b2[c * 4 + 2] = word2(bu.z); b2[c * 4 + 3] = word2(bu.w);
  }
  auto row_ptr = [&](int row) {
    const int grp = row / rows_per_group;
    const int idx = row - grp * rows_per_group;
    return reinterpret_cast<const uint4*>(x + (size_t(grp) * in_group_stride + in_skip + idx) * LN_D);
  };
  uint4 nxt[NCH];
  {
    const uint4* xp = row_ptr(r);
#pragma unroll
    for (int c = 0; c < NCH; ++c) nxt[c] = xp[c * 32 + lane];
  }
  for (; r < rows; r += nw) {
    f32x2 v[NP];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      v[c * 4 + 0] = word2(nxt[c].x); v[c * 4 + 1] = word2(nxt[c].y); v[c * 4 + 2] = word2(nxt[c].z); v[c * 4 + 3] = word2(nxt[c].w);
    }
    if (r + nw < rows) {
      const uint4* xp = row_ptr(r + nw);
#pragma unroll
      for (int c = 0; c < NCH; ++c) nxt[c] = xp[c * 32 + lane];
    }
    lnrow::normalise_pairs<NP>(v, 1.0f / LN_D, eps);   // (x - mean) * rstd
    uint4* op = reinterpret_cast<uint4*>(out + size_t(r) * LN_D);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = lnrow::affine_word(v[c * 4 + j], w2[c * 4 + j], b2[c * 4 + j]);
      op[c * 32 + lane] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Image -> patch matrix.  Output element (row = b*g*g + py*g + px, col = c*196 + ky*14 + kx) is pixel
// (b, c, py*14 + ky, px*14 + kx); columns [588, Kpad) are zero (the GEMM's K is padded to a 64 multiple).
// SRC_F32: input is the fp32 [0,1] image and the reference's bf16 Normalize arithmetic is applied
// (dino.py:12,16 with the model in bf16: tensor.sub_(mean).div_(std) on bf16 tensors).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float normalize_px(float x, int c) {
  // bf16(mean), bf16(std) -- as_tensor(mean, dtype=bf16) in torchvision's F.normalize
  const float mean = c == 0 ? 0.484375f : (c == 1 ? 0.455078125f : 0.40625f);
  const float stdv = c == 0 ? 0.228515625f : (c == 1 ? 0.2236328125f : 0.224609375f);
  const float xb = bf16_round(x);
  const float d = bf16_round(__fsub_rn(xb, mean));
  return bf16_round(__fdiv_rn(d, stdv));
}

template <bool SRC_F32>
__global__ void __launch_bounds__(256)
im2col_kernel(const void* __restrict__ img, bf16* __restrict__ out, int B, int res, int g, int Kpad) {
  const size_t total = size_t(B) * g * g * Kpad;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int col = int(i % Kpad);
    const size_t row = i / Kpad;
    float v = 0.f;
    if (col < 588) {
      const int c = col / 196, rem = col - c * 196;
      const int ky = rem / 14, kx = rem - ky * 14;
      const int px = int(row % g);
      const size_t t = row / g;
      const int py = int(t % g);
      const size_t b = t / g;
      const size_t src = ((b * 3 + c) * res + (py * 14 + ky)) * res + (px * 14 + kx);
      if (SRC_F32)
        v = normalize_px(reinterpret_cast<const float*>(img)[src], c);
      else
        v = __bfloat162float(reinterpret_cast<const bf16*>(img)[src]);
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

__global__ void __launch_bounds__(256)
normalize_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int res) {
  const size_t plane = size_t(res) * res;
  const size_t total = size_t(B) * 3 * plane;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int c = int((i / plane) % 3);
    out[i] = __float2bfloat16_rn(normalize_px(img[i], c));
  }
}

__global__ void __launch_bounds__(128)
special_tokens_kernel(const bf16* __restrict__ special, bf16* __restrict__ tokens, int B, int T, int n_special, int D) {
  // grid = (B * n_special), 128 threads x 16 bytes = 1024 bf16
  const int b = blockIdx.x / n_special, j = blockIdx.x - b * n_special;
  const uint4* src = reinterpret_cast<const uint4*>(special + size_t(j) * D);
  uint4* dst = reinterpret_cast<uint4*>(tokens + (size_t(b) * T + j) * D);
  for (int i = threadIdx.x; i < D / 8; i += blockDim.x) dst[i] = __ldg(src + i);
}

}  // namespace

int layernorm_bf16(const bf16* x, const bf16* w, const bf16* b, bf16* out, int rows, int D, float eps,
                   int in_group_stride, int in_skip, int rows_per_group, cudaStream_t stream) {
  FP_REQUIRE(D == 1024 || D == 768, "layernorm: D=%d unsupported (ViT-L: 1024, ViT-B: 768)", D);
  if (rows <= 0) return 0;
  FP_REQUIRE(rows_per_group > 0, "layernorm: rows_per_group must be positive");
  const int want = (rows + LN_WARPS - 1) / LN_WARPS;
  const int blocks = want < 2 * sm_count() ? want : 2 * sm_count();
  ProfScope prof(PROF_LAYERNORM, 2.0 * double(rows) * D * 2, 1, stream);
  if (D == 1024)
    layernorm_kernel<4><<<blocks, LN_WARPS * 32, 0, stream>>>(x, w, b, out, rows, eps, in_group_stride, in_skip,
                                                              rows_per_group);
  else
    layernorm_kernel<3><<<blocks, LN_WARPS * 32, 0, stream>>>(x, w, b, out, rows, eps, in_group_stride, in_skip,
                                                              rows_per_group);
  FP_CUDA(cudaGetLastError());
  return 0;
}

static int grid_for(size_t total, int threads) {
  size_t blocks = (total + threads - 1) / threads;
  const size_t cap = size_t(sm_count()) * 16;
  return int(blocks < cap ? (blocks ? blocks : 1) : cap);
}

int im2col_patches(const void* img, int src_is_f32, bf16* out, int B, int res, int Kpad, cudaStream_t stream) {
  FP_REQUIRE(res % 14 == 0 && res > 0, "im2col: resolution %d is not a multiple of the 14-pixel patch", res);
  FP_REQUIRE(Kpad >= 588 && Kpad % 64 == 0, "im2col: Kpad=%d must be a multiple of 64 and >= 588", Kpad);
  if (B <= 0) return 0;
  const int g = res / 14;
  const size_t total = size_t(B) * g * g * Kpad;
  ProfScope prof(PROF_TOKEN_PREP, double(total) * 2 + double(B) * 3 * res * res * (src_is_f32 ? 4 : 2), 1, stream);
  if (src_is_f32)
    im2col_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(img, out, B, res, g, Kpad);
  else
    im2col_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(img, out, B, res, g, Kpad);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int normalize_image(const float* img, bf16* out, int B, int res, cudaStream_t stream) {
  if (B <= 0) return 0;
  const size_t total = size_t(B) * 3 * res * res;
  ProfScope prof(PROF_TOKEN_PREP, double(total) * 6, 1, stream);
  normalize_kernel<<<grid_for(total, 256), 256, 0, stream>>>(img, out, B, res);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int write_special_tokens(const bf16* special, bf16* tokens, int B, int T, int n_special, int D, cudaStream_t stream) {
  FP_REQUIRE(D % 8 == 0, "special tokens: D must be a multiple of 8");
  if (B <= 0 || n_special <= 0) return 0;
  ProfScope prof(PROF_TOKEN_PREP, double(B) * n_special * D * 2, 1, stream);
  special_tokens_kernel<<<B * n_special, 128, 0, stream>>>(special, tokens, B, T, n_special, D);
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fp
