// Geometry around the renders:
//   mask_bbox               depth>0 mask -> bounding box (+ the reference's tiny-mask fallback square)
//   crop_resize_pad         reference CropResizePad (src/utils/bbox_utils.py:20-56): crop, nearest resize,
//                           centred zero pad, second nearest resize -- as ONE gather, writing either the
//                           fp32 CHW crop (API parity) or straight into the normalised bf16 patch matrix
//                           of the patch-embed GEMM (hot path: no fp32 crops in HBM at all)
//   depth_extents           min/max/sum of the back-projected depth map (reference src/pipeline/utils.py:
//                           122-170 only needs these to derive the translation)
#include "common.cuh"
#include "kernels.h"

namespace fp {

namespace {

// ------------------------------------------------------------------------------------------------ bbox
__global__ void __launch_bounds__(256)
mask_bbox_kernel(const float* __restrict__ depth, int res, int fb_lo, int fb_hi, int min_count,
                 int32_t* __restrict__ bbox_out, int32_t* __restrict__ count_out, uint8_t* __restrict__ mask_out) {
  const int b = blockIdx.x;
  const float* d = depth + size_t(b) * res * res;
  int xmin = INT_MAX, ymin = INT_MAX, xmax = -1, ymax = -1, cnt = 0;
  const int quads = res >> 2;
  for (int i = threadIdx.x; i < quads * res; i += blockDim.x) {
    const int y = i / quads, x0 = (i - y * quads) << 2;
    const float4 v = *reinterpret_cast<const float4*>(d + size_t(y) * res + x0);
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (vv[j] > 0.f) {
        ++cnt;
        xmin = min(xmin, x0 + j); xmax = max(xmax, x0 + j);
        ymin = min(ymin, y); ymax = max(ymax, y);
      }
  }
  __shared__ int s[5][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) { s[0][warp] = xmin; s[1][warp] = ymin; s[2][warp] = xmax; s[3][warp] = ymax; s[4][warp] = cnt; }
  __syncthreads();
  __shared__ int use_fallback;
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      xmin = min(xmin, s[0][w]); ymin = min(ymin, s[1][w]);
      xmax = max(xmax, s[2][w]); ymax = max(ymax, s[3][w]);
      cnt += s[4][w];
    }
    const int fb = cnt < min_count;
    if (fb) {  // reference renderer.py:116-117: mask[lo:hi, lo:hi] = True (numpy slice clipping applies)
      const int lo = min(fb_lo, res), hi = min(fb_hi, res);
      if (hi > lo) {
        xmin = min(xmin, lo); ymin = min(ymin, lo);
        xmax = max(xmax, hi - 1); ymax = max(ymax, hi - 1);
      }
    }
    bbox_out[4 * b] = xmin; bbox_out[4 * b + 1] = ymin; bbox_out[4 * b + 2] = xmax; bbox_out[4 * b + 3] = ymax;
    if (count_out) count_out[b] = cnt;
    use_fallback = fb;
  }
  if (mask_out != nullptr) {
    __syncthreads();
    const int fb = use_fallback;
    uint8_t* m = mask_out + size_t(b) * res * res;
    for (int i = threadIdx.x; i < res * res; i += blockDim.x) {
      const int y = i / res, x = i - y * res;
      const bool in_fb = fb && y >= fb_lo && y < fb_hi && x >= fb_lo && x < fb_hi;
      m[i] = (d[i] > 0.f || in_fb) ? 1 : 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------ crop
struct CropParams {
  int x1, y1, w0, h0;   // crop window in the source image
  int w1, h1;           // size after the first nearest resize
  int pad_left, pad_top, padded;
  int s2;               // side of the (square) image entering the second resize
  float inv1, inv2;     // float32(1 / scale) of the two resizes
  int ok;
};

// Restates CropResizePad.__call__ for one box (already extended/clamped; xyxy, x2/y2 exclusive).
__device__ inline CropParams crop_params(int x1, int y1, int x2, int y2, int T) {
  CropParams c;
  c.x1 = x1; c.y1 = y1; c.w0 = x2 - x1; c.h0 = y2 - y1;
  c.ok = (c.w0 > 0 && c.h0 > 0);
  if (!c.ok) { c.w1 = c.h1 = c.pad_left = c.pad_top = c.padded = c.s2 = 0; c.inv1 = c.inv2 = 0.f; return c; }
  const int m = max(c.w0, c.h0);
  // scale_factor = target_max / max(box_sizes): Tensor.__rtruediv__ = reciprocal(tensor) * scalar in float32
  const float scale32 = __fmul_rn(__frcp_rn(float(m)), float(T));
  const double s = double(scale32);                    // scale.item()
  c.h1 = int(floor(double(c.h0) * s));                 // F.interpolate output size
  c.w1 = int(floor(double(c.w0) * s));
  c.inv1 = float(1.0 / s);
  c.ok = c.h1 > 0 && c.w1 > 0 && c.h1 <= T && c.w1 <= T;
  c.padded = (c.w1 != c.h1);                           // target_ratio (1.0) != original_ratio
  if (c.padded) {
    c.pad_top = max((T - c.h1) / 2, 0);
    c.pad_left = max((T - c.w1) / 2, 0);
    c.s2 = T;
    c.inv2 = 1.0f;
  } else {
    c.pad_top = c.pad_left = 0;
    c.s2 = c.h1;
    const double s2 = double(T) / double(c.h1);        // self.target_h / image.shape[1]
    c.inv2 = float(1.0 / s2);
    if (int(floor(double(c.h1) * s2)) != T) c.ok = 0;  // the reference's torch.stack would fail here
  }
  return c;
}

__device__ __forceinline__ int nearest_src(int dst, float inv, int in_size) {
  const int i = int(floorf(__fmul_rn(float(dst), inv)));
  return min(i, in_size - 1);
}

// Maps an output pixel of the T x T crop to a source pixel; returns false for padding.
__device__ __forceinline__ bool crop_source(const CropParams& c, int Y, int X, int& sy, int& sx) {
  int my = Y, mx = X;
  if (!c.padded) {  // square after the first resize: the second resize is a real (near-identity) resample
    my = nearest_src(Y, c.inv2, c.s2);
    mx = nearest_src(X, c.inv2, c.s2);
  }
  const int yy = my - c.pad_top, xx = mx - c.pad_left;
  if (yy < 0 || yy >= c.h1 || xx < 0 || xx >= c.w1) return false;
  sy = c.y1 + nearest_src(yy, c.inv1, c.h0);
  sx = c.x1 + nearest_src(xx, c.inv1, c.w0);
  return true;
}

// One axis of crop_source: output coordinate `o` -> source coordinate along that axis; false for padding.
__device__ __forceinline__ bool crop_source_axis(const CropParams& c, int o, int pad, int size1, int size0, int origin,
                                                 int& src) {
  int m = o;
  if (!c.padded) m = nearest_src(o, c.inv2, c.s2);
  const int t = m - pad;
  if (t < 0 || t >= size1) return false;
  src = origin + nearest_src(t, c.inv1, size0);
  return true;
}

// SRC_U8: source is u8 HWC (a render); else fp32 CHW.  DST_PATCH: write the normalised bf16 patch matrix;
// else the fp32 CHW crop in [0,1].
template <bool SRC_U8, bool DST_PATCH>
__global__ void __launch_bounds__(256)
crop_kernel(const void* __restrict__ src, const int32_t* __restrict__ boxes, int box_is_inclusive,
            const bf16* __restrict__ norm_lut, void* __restrict__ dst, int src_h, int src_w, int T, int Kpad,
            int32_t* __restrict__ status) {
  const int b = blockIdx.y;
  __shared__ CropParams cp;
  if (threadIdx.x == 0) {
    const int32_t* bx = boxes + 4 * b;
    // renders: box = (xmin, ymin, xmax, ymax) of the mask and the reference slices [ymin:ymax, xmin:xmax]
    // (its own off-by-one: the last row/column is dropped) -- so the box is used as-is, exclusive.
    (void)box_is_inclusive;
    cp = crop_params(bx[0], bx[1], bx[2], bx[3], T);
    if (!cp.ok && status) atomicExch(status, b + 1);
  }
  __syncthreads();
  const CropParams c = cp;
  const int g = T / 14;
  if (DST_PATCH) {
    // The composed crop is separable (source row depends on the output row only, source column on the output column
    // only): tabulate both maps once per CTA, then every thread gathers 8 consecutive matrix columns and writes 16 bytes.
    __shared__ short map_y[1024], map_x[1024];
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
      int ty, tx;
      map_y[i] = (c.ok && crop_source_axis(c, i, c.pad_top, c.h1, c.h0, c.y1, ty)) ? short(ty) : short(-1);
      map_x[i] = (c.ok && crop_source_axis(c, i, c.pad_left, c.w1, c.w0, c.x1, tx)) ? short(tx) : short(-1);
    }
    __syncthreads();
    bf16* out = reinterpret_cast<bf16*>(dst) + size_t(b) * g * g * Kpad;
    const int groups_per_row = Kpad / 8;
    const int total = g * g * groups_per_row;
    const uint8_t* img = reinterpret_cast<const uint8_t*>(src) + size_t(b) * src_h * src_w * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int row = i / groups_per_row, col0 = (i - row * groups_per_row) * 8;
      const int py = row / g, px = row - py * g;
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        uint32_t pair = 0;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int col = col0 + j + h2;
          bf16 v = __float2bfloat16_rn(0.f);
          if (col < 588) {
            const int ch = col / 196, rem = col - ch * 196;
            const int ky = rem / 14, kx = rem - ky * 14;
            const int sy = map_y[py * 14 + ky], sx = map_x[px * 14 + kx];
            int val = 0;
            if (SRC_U8 && sy >= 0 && sx >= 0) val = img[(size_t(sy) * src_w + sx) * 3 + ch];
            v = norm_lut[ch * 256 + val];
          }
          pair |= uint32_t(*reinterpret_cast<const uint16_t*>(&v)) << (16 * h2);
        }
        w[j >> 1] = pair;
      }
      *reinterpret_cast<uint4*>(out + size_t(row) * Kpad + col0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else {
    float* out = reinterpret_cast<float*>(dst) + size_t(b) * 3 * T * T;
    const int total = 3 * T * T;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int X = i % T, Y = (i / T) % T, ch = i / (T * T);
      int sy, sx;
      float v = 0.f;
      if (c.ok && crop_source(c, Y, X, sy, sx)) {
        if (SRC_U8) {
          const int u = reinterpret_cast<const uint8_t*>(src)[((size_t(b) * src_h + sy) * src_w + sx) * 3 + ch];
          v = float(double(u) / 255.0);  // torch.from_numpy(img / 255).float()
        } else {
          v = reinterpret_cast<const float*>(src)[((size_t(b) * 3 + ch) * src_h + sy) * src_w + sx];
        }
      }
      out[i] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ extents
__global__ void __launch_bounds__(256)
depth_extents_kernel(const float* __restrict__ depth, const int32_t* __restrict__ view_idx, int res,
                     const double* __restrict__ kinv, double* __restrict__ out) {
  const int o = blockIdx.x;
  const int b = view_idx ? view_idx[o] : o;
  const float* d = depth + size_t(b) * res * res;
  const double k00 = kinv[0], k01 = kinv[1], k02 = kinv[2], k10 = kinv[3], k11 = kinv[4], k12 = kinv[5],
               k20 = kinv[6], k21 = kinv[7], k22 = kinv[8];
  double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY, sx = 0, sy = 0, sz = 0;
  long long cnt = 0;
  for (int i = threadIdx.x; i < res * res; i += blockDim.x) {
    const int v = i / res, u = i - v * res;
    const double dd = double(d[i]);
    // K^-1 [u v 1]^T * d  (reference utils.py:141); rows that are all zero are dropped (utils.py:144)
    const double X = __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(k00, u), __dmul_rn(k01, v)), k02), dd);
    const double Y = __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(k10, u), __dmul_rn(k11, v)), k12), dd);
    const double Z = __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(k20, u), __dmul_rn(k21, v)), k22), dd);
    if (!(X == 0.0 && Y == 0.0 && Z == 0.0)) {
      xmin = fmin(xmin, X); xmax = fmax(xmax, X);
      ymin = fmin(ymin, Y); ymax = fmax(ymax, Y);
      sx += X; sy += Y; sz += Z; ++cnt;
    }
  }
  __shared__ double sm[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, off));
    xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, off));
    ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, off));
    ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, off));
    sx += __shfl_xor_sync(0xffffffffu, sx, off);
    sy += __shfl_xor_sync(0xffffffffu, sy, off);
    sz += __shfl_xor_sync(0xffffffffu, sz, off);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  }
  if (lane == 0) {
    sm[warp][0] = xmin; sm[warp][1] = xmax; sm[warp][2] = ymin; sm[warp][3] = ymax;
    sm[warp][4] = sx; sm[warp][5] = sy; sm[warp][6] = sz; sm[warp][7] = double(cnt);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double c = 0;
    for (int w = 1; w < 8; ++w) {
      xmin = fmin(xmin, sm[w][0]); xmax = fmax(xmax, sm[w][1]);
      ymin = fmin(ymin, sm[w][2]); ymax = fmax(ymax, sm[w][3]);
    }
    sx = sy = sz = 0;
    for (int w = 0; w < 8; ++w) { sx += sm[w][4]; sy += sm[w][5]; sz += sm[w][6]; c += sm[w][7]; }
    double* r = out + size_t(o) * 8;
    r[0] = xmin; r[1] = xmax; r[2] = ymin; r[3] = ymax; r[4] = sx; r[5] = sy; r[6] = sz; r[7] = c;
  }
}

}  // namespace

int mask_bbox(const float* depth, int B, int res, int fallback_lo, int fallback_hi, int min_count,
              int32_t* bbox_out, int32_t* count_out, uint8_t* mask_out, cudaStream_t stream) {
  FP_REQUIRE(res > 0 && res % 4 == 0, "mask_bbox: resolution must be a multiple of 4");
  if (B <= 0) return 0;
  ProfScope prof(PROF_GEOMETRY, double(B) * res * res * 4, 1, stream);
  mask_bbox_kernel<<<B, 256, 0, stream>>>(depth, res, fallback_lo, fallback_hi, min_count, bbox_out, count_out,
                                          mask_out);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int crop_resize_pad(const void* src, int src_is_u8_hwc, const int32_t* boxes, const bf16* norm_lut, void* dst,
                    int dst_is_patches, int B, int src_h, int src_w, int T, int Kpad, int32_t* status,
                    cudaStream_t stream) {
  FP_REQUIRE(T > 0 && (!dst_is_patches || T % 14 == 0), "crop: target size %d must be a multiple of 14", T);
  FP_REQUIRE(!dst_is_patches || (Kpad >= 588 && Kpad % 64 == 0), "crop: bad Kpad %d", Kpad);
  FP_REQUIRE(!dst_is_patches || (norm_lut != nullptr && src_is_u8_hwc),
             "crop: the patch-matrix output needs a u8 source and the normalisation LUT");
  FP_REQUIRE(B <= 65535, "crop: at most 65535 images per call");
  FP_REQUIRE(!dst_is_patches || T <= 1024, "crop: patch-matrix targets above 1024 px are not supported");
  if (B <= 0) return 0;
  const int g = T / 14;
  const int total = dst_is_patches ? g * g * Kpad : 3 * T * T;
  const dim3 grid(min((total + 255) / 256, 64), B);
  ProfScope prof(PROF_GEOMETRY, double(B) * total * (dst_is_patches ? 2 : 4), 1, stream);
  if (src_is_u8_hwc && dst_is_patches)
    crop_kernel<true, true><<<grid, 256, 0, stream>>>(src, boxes, 0, norm_lut, dst, src_h, src_w, T, Kpad, status);
  else if (src_is_u8_hwc)
    crop_kernel<true, false><<<grid, 256, 0, stream>>>(src, boxes, 0, norm_lut, dst, src_h, src_w, T, Kpad, status);
  else
    crop_kernel<false, false><<<grid, 256, 0, stream>>>(src, boxes, 0, norm_lut, dst, src_h, src_w, T, Kpad, status);
  FP_CUDA(cudaGetLastError());
  return 0;
}

int depth_extents(const float* depth, const int32_t* view_idx, int n_out, int res, const double* kinv_dev,
                  double* out, cudaStream_t stream) {
  if (n_out <= 0) return 0;
  ProfScope prof(PROF_GEOMETRY, double(n_out) * res * res * 4, 1, stream);
  depth_extents_kernel<<<n_out, 256, 0, stream>>>(depth, view_idx, res, kinv_dev, out);
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fp
