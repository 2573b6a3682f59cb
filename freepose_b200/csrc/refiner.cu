// Kernels of the refiner's pose-confidence pass (reference
// src/pipeline/estimators/tracking_refiner.py:45-100 and src/pipeline/refiner_utils.py:92-137): the photo crop
// (torchvision.ops.roi_align), the 37x37 validity mask (cv2.resize INTER_CUBIC of depth > 0) and the masked per-patch
// cosine between the photo's and the render's ViT-B tokens.  All three are small HBM-bound gathers / reductions.
#include "common.cuh"
#include "kernels.h"

namespace fp {

namespace {

// torchvision's bilinear_interpolate (roi_align_common.h / roi_align_kernel.cu), evaluated without FMA contraction in
// the CPU kernel's expression order (the reference crops on CPU tensors).
__device__ __forceinline__ float roi_bilinear(const float* __restrict__ img, int H, int W, float y, float x) {
  if (y < -1.0f || y > float(H) || x < -1.0f || x > float(W)) return 0.f;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int y_low = int(y), x_low = int(x), y_high, x_high;
  if (y_low >= H - 1) { y_high = y_low = H - 1; y = float(y_low); } else { y_high = y_low + 1; }
  if (x_low >= W - 1) { x_high = x_low = W - 1; x = float(x_low); } else { x_high = x_low + 1; }
  const float ly = __fsub_rn(y, float(y_low)), lx = __fsub_rn(x, float(x_low));
  const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
  const float v1 = img[size_t(y_low) * W + x_low], v2 = img[size_t(y_low) * W + x_high];
  const float v3 = img[size_t(y_high) * W + x_low], v4 = img[size_t(y_high) * W + x_high];
  const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx), w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)), __fmul_rn(w3, v3)), __fmul_rn(w4, v4));
}

// roi_align(image[None], boxes, (out_h, out_w), spatial_scale=1, sampling_ratio=S, aligned=False)
__global__ void __launch_bounds__(256)
roi_align_kernel(const float* __restrict__ image, int C, int H, int W, const float* __restrict__ boxes, int n,
                 int out_h, int out_w, int S, float* __restrict__ out) {
  const size_t total = size_t(n) * C * out_h * out_w;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int pw = int(i % out_w);
    const int ph = int((i / out_w) % out_h);
    const int c = int((i / (size_t(out_w) * out_h)) % C);
    const int r = int(i / (size_t(out_w) * out_h * C));
    const float x1 = boxes[4 * r], y1 = boxes[4 * r + 1], x2 = boxes[4 * r + 2], y2 = boxes[4 * r + 3];
    const float roi_w = fmaxf(__fsub_rn(x2, x1), 1.0f), roi_h = fmaxf(__fsub_rn(y2, y1), 1.0f);
    const float bin_h = __fdiv_rn(roi_h, float(out_h)), bin_w = __fdiv_rn(roi_w, float(out_w));
    const int gh = S > 0 ? S : int(ceilf(__fdiv_rn(roi_h, float(out_h))));
    const int gw = S > 0 ? S : int(ceilf(__fdiv_rn(roi_w, float(out_w))));
    const float count = float(max(gh * gw, 1));
    const float* plane = image + size_t(c) * H * W;
    float acc = 0.f;
    for (int iy = 0; iy < gh; ++iy) {
      const float y = __fadd_rn(__fadd_rn(y1, __fmul_rn(float(ph), bin_h)),
                                __fdiv_rn(__fmul_rn(__fadd_rn(float(iy), 0.5f), bin_h), float(gh)));
      for (int ix = 0; ix < gw; ++ix) {
        const float x = __fadd_rn(__fadd_rn(x1, __fmul_rn(float(pw), bin_w)),
                                  __fdiv_rn(__fmul_rn(__fadd_rn(float(ix), 0.5f), bin_w), float(gw)));
        acc = __fadd_rn(acc, roi_bilinear(plane, H, W, y, x));
      }
    }
    out[i] = __fdiv_rn(acc, count);
  }
}

// cv2.resize(mask.astype(float32), (g, g), interpolation=INTER_CUBIC) > 0.5 with mask = depth > 0: separable 4-tap
// cubic (A = -0.75) around source coordinate (d + 0.5) * res / g - 0.5, replicated borders, no antialiasing.  For the
// refiner's 518 -> 37 the scale is exactly 14, the fractional offset 0.5 and every weight a multiple of 1/32, so the
// fp32 result is exact whatever the summation order.
__device__ __forceinline__ void cubic_coeffs(float x, float w[4]) {
  const float A = -0.75f;
  w[0] = ((A * (x + 1.f) - 5.f * A) * (x + 1.f) + 8.f * A) * (x + 1.f) - 4.f * A;
  w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  w[2] = ((A + 2.f) * (1.f - x) - (A + 3.f)) * (1.f - x) * (1.f - x) + 1.f;
  w[3] = 1.f - w[0] - w[1] - w[2];
}

__global__ void __launch_bounds__(256)
depth_mask_cubic_kernel(const float* __restrict__ depth, int B, int res, int stride, int g, uint8_t* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * g * g) return;
  const int dx = i % g, dy = (i / g) % g, b = i / (g * g);
  const double scale = double(res) / double(g);
  const float fx = float((dx + 0.5) * scale - 0.5), fy = float((dy + 0.5) * scale - 0.5);
  const int sx = int(floorf(fx)), sy = int(floorf(fy));
  float wx[4], wy[4];
  cubic_coeffs(fx - float(sx), wx);
  cubic_coeffs(fy - float(sy), wy);
  const float* d = depth + size_t(b) * stride * stride;   // the top-left res x res of a stride x stride image
  float acc = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int yy = min(max(sy - 1 + r, 0), res - 1);
    float row = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int xx = min(max(sx - 1 + c, 0), res - 1);
      row += (d[size_t(yy) * stride + xx] > 0.f ? 1.f : 0.f) * wx[c];
    }
    acc += row * wy[r];
  }
  mask[i] = acc > 0.5f ? 1 : 0;
}

// out[r] = mask[r] * <a_r / |a_r|, b_r / |b_r|>  (tracking_refiner.py:80-88, fp32 arithmetic on the bf16 tokens)
__global__ void __launch_bounds__(256)
patch_cosine_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, const uint8_t* __restrict__ mask, int rows,
                    int D, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const uint4* pa = reinterpret_cast<const uint4*>(a + size_t(warp) * D);
  const uint4* pb = reinterpret_cast<const uint4*>(b + size_t(warp) * D);
  float saa = 0.f, sbb = 0.f;
  // two passes like the reference: norms first, then the dot product of the normalised vectors
  for (int i = lane; i < D / 8; i += 32) {
    const uint4 ua = pa[i], ub = pb[i];
    const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      saa = fmaf(bf16lo(wa[j]), bf16lo(wa[j]), saa); saa = fmaf(bf16hi(wa[j]), bf16hi(wa[j]), saa);
      sbb = fmaf(bf16lo(wb[j]), bf16lo(wb[j]), sbb); sbb = fmaf(bf16hi(wb[j]), bf16hi(wb[j]), sbb);
    }
  }
  const float na = sqrtf(warp_sum(saa)), nb = sqrtf(warp_sum(sbb));
  float dot = 0.f;
  for (int i = lane; i < D / 8; i += 32) {
    const uint4 ua = pa[i], ub = pb[i];
    const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dot = fmaf(bf16lo(wa[j]) / na, bf16lo(wb[j]) / nb, dot);
      dot = fmaf(bf16hi(wa[j]) / na, bf16hi(wb[j]) / nb, dot);
    }
  }
  dot = warp_sum(dot);
  if (lane == 0) out[warp] = (mask == nullptr || mask[warp]) ? dot : 0.f;
}

int grid_for(size_t total, int threads) {
  size_t blocks = (total + threads - 1) / threads;
  const size_t cap = size_t(sm_count()) * 16;
  return int(blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace

int roi_align(const float* image, int C, int H, int W, const float* boxes, int n, int out_h, int out_w,
              int sampling_ratio, float* out, cudaStream_t stream) {
  FP_REQUIRE(C > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0, "roi_align: bad sizes");
  if (n <= 0) return 0;
  const size_t total = size_t(n) * C * out_h * out_w;
  ProfScope prof(PROF_GEOMETRY, double(total) * 4, 1, stream);
  roi_align_kernel<<<grid_for(total, 256), 256, 0, stream>>>(image, C, H, W, boxes, n, out_h, out_w, sampling_ratio, out);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int depth_mask_cubic(const float* depth, int B, int res, int stride, int g, uint8_t* mask, cudaStream_t stream) {
  FP_REQUIRE(res > 0 && g > 0 && g <= res && stride >= res, "depth_mask_cubic: bad sizes (res=%d, stride=%d, g=%d)", res, stride, g);
  if (B <= 0) return 0;
  ProfScope prof(PROF_GEOMETRY, double(B) * g * g * 65, 1, stream);
  depth_mask_cubic_kernel<<<(B * g * g + 255) / 256, 256, 0, stream>>>(depth, B, res, stride, g, mask);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int patch_cosine(const bf16* a, const bf16* b, const uint8_t* mask, int rows, int D, float* out, cudaStream_t stream) {
  FP_REQUIRE(D > 0 && D % 8 == 0, "patch_cosine: D=%d must be a multiple of 8", D);
  if (rows <= 0) return 0;
  ProfScope prof(PROF_SCORE, 2.0 * double(rows) * D * 2, 1, stream);
  patch_cosine_kernel<<<(rows * 32 + 255) / 256, 256, 0, stream>>>(a, b, mask, rows, D, out);
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fp
