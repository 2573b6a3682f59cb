// Batched triangle rasteriser for pose-hypothesis views (replaces the per-pose
// pyrender/OpenGL draw + glReadPixels loop of reference src/pipeline/retrieval/renderer.py:43-95).
//
// Semantics restated from the reference's pyrender set-up (renderer.py:37-66):
//   * pinhole camera (fx, fy, cx, cy), OpenCV camera frame (the reference's camera node pose diag(1,-1,-1,1)
//     makes world == OpenCV camera); window coordinate u = fx*X/Z + cx, pixel i covers [i, i+1);
//   * ambient-only lighting (2,2,2): colour = clamp(pow(2 * base_colour, 1/2.2), 0, 1) -> unorm8; transparent
//     black background; SKIP_CULL_FACES (two-sided) unless cull_backfaces;
//   * 4x multisampling (pyrender's offscreen framebuffer): coverage and depth per sample, one shading
//     evaluation per (triangle, pixel) at the pixel centre, box-filter resolve in unorm8; the depth image is
//     sample 0's linear depth (0 = background).  msaa = 1 evaluates everything at the pixel centre.
//
// Everything that decides *which* triangle a sample sees is integer arithmetic (24.8 fixed-point vertices,
// 64-bit edge functions, top-left rule) and the depth test is an order-independent 64-bit atomicMin on
// (depth bits, face id), so results are deterministic and the CPU oracle (oracle/raster_ref.c) restates them
// bit for bit.  All fp32 arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction).
//
// Surfaces (pyrender material paths): per-vertex colours (trimesh ColorVisuals -> COLOR_0), a base-colour texture
// (trimesh TextureVisuals -> baseColorTexture: REPEAT wrap, trilinear filtering over a box-filtered mip chain, filtered
// in the stored sRGB values and linearised with x^2.2 afterwards, as pyrender's mesh.frag does), or both multiplied.
// trimesh.PointCloud inputs (renderer.py:46-51) are GL points of size 1: a one-pixel square sprite centred on the
// projected vertex, flat vertex colour, the vertex's own depth.
//
// Near / far planes and the guard band: a triangle with a vertex that does not project (Z <= znear, Z >= zfar, or more
// than 16384 px off screen) is NOT dropped.  It is rasterised in homogeneous form (Olano & Greer): with camera-space
// vertices P0..P2 and the sample ray d = ((sx - cx)/fx, (sy - cy)/fy, 1), b_i = sign(det) d.(P_{i+1} x P_{i+2}); covered
// iff all b_i >= 0 and sum > 0; depth |det| / sum; the per-sample test znear < z < zfar then removes exactly what GL's
// clipping against the near / far planes removes.  Only views that contain such a vertex run that kernel at all.
//
// Two pipelines behind one entry point.  The GENERAL pipeline is the default for every view; with FP_RASTER_TILE=1 the
// views whose vertices all project (a flag the vertex kernel sets per view) take the TILE pipeline instead -- an
// experiment that cut the DRAM traffic as intended but lost on time (see launch_raster):
//   * TILE pipeline.  Triangles
//     are binned into 16 x 16-pixel tiles (count -> scan -> fill; triangles spanning more than 2 x 2 tiles go to a per-view
//     list instead), then ONE kernel per tile depth-tests the samples in SHARED memory, shades and writes RGB + depth.  No
//     sample-key buffer exists in HBM: the previous pipeline cleared, atomically updated and re-read 32 bytes per pixel
//     (835 MB per 521 views, 12.6 x the 183 MB of output).
//   * GENERAL pipeline (also the only one for views with a vertex behind the near plane / outside the guard band, and
//     for point clouds): 64-bit sample keys in HBM, atomicMin from the triangle / hard-triangle / point kernels, resolve passes.
// Both evaluate the same integer edge functions and the same depth expression and keep the minimum (depth, face) key
// per sample, so they produce identical images.
//
// General pipeline kernels: clear keys -> vertex transform -> triangle (32-bit edge functions for small triangles, warp-cooperative walk
// for large ones) or point scatter -> resolve (one thread per pixel; coalesced depth, shuffle-assembled RGB words).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace fp {

namespace {

constexpr int SUB = 8;                 // sub-pixel bits
constexpr int ONE = 1 << SUB;          // 256
constexpr float ZNEAR_DEFAULT = 0.05f;   // pyrender IntrinsicsCamera defaults (renderer.py:37)
constexpr float ZFAR_DEFAULT = 100.0f;
constexpr int COORD_LIMIT = 1 << 22;   // |fixed-point coordinate| guard (16384 px)

struct __align__(16) ScreenVertex {
  int x, y;      // 24.8 fixed point, image coordinates (y down)
  float z;       // camera-space depth
  float iz;      // 1 / z
};

__constant__ int c_sample_off[2][4][2] = {
    {{128, 128}, {128, 128}, {128, 128}, {128, 128}},   // msaa 1: pixel centre
    {{96, 32}, {224, 96}, {32, 160}, {160, 224}},       // msaa 4: (0.375,0.125) (0.875,0.375) (0.125,0.625) (0.625,0.875)
};

// Screen box of a view: the pixel bounding box of its projected vertices, reduced by the vertex kernel into four
// zero-initialised words with atomicMax -- {BOX_BIG - min px, BOX_BIG - min py, max px + 1, max py + 1}, 0 = no vertex.
// Every sample a triangle / point sprite of the view can cover lies inside it, so the general pipeline clears and
// resolves the 32 bytes per pixel of sample keys only there (a sphere-like object at the benchmark pose covers ~40 % of
// the image; the rest of the key buffer is neither written nor read).  Views with a vertex that does not project
// (view_hard) use the whole image: a clipped triangle can cover anything.
constexpr int BOX_BIG = 1 << 24;
struct ViewBox { int x0, y0, x1, y1; };   // inclusive, clipped to the image; x0 > x1: empty
__device__ __forceinline__ ViewBox view_box(const int* __restrict__ boxes, const int* __restrict__ view_hard, int b, int res) {
  ViewBox vb;
  if (view_hard[b]) { vb.x0 = 0; vb.y0 = 0; vb.x1 = res - 1; vb.y1 = res - 1; return vb; }
  const int4 w = reinterpret_cast<const int4*>(boxes)[b];
  if (w.x == 0) { vb.x0 = 1; vb.y0 = 1; vb.x1 = 0; vb.y1 = 0; return vb; }
  vb.x0 = max(0, BOX_BIG - w.x - 1); vb.y0 = max(0, BOX_BIG - w.y - 1);     // one pixel of margin (point sprites)
  vb.x1 = min(res - 1, w.z - BOX_BIG / 2); vb.y1 = min(res - 1, w.w - BOX_BIG / 2);   // (max px + 1: one pixel of margin)
  return vb;
}

// `route` (per view, written by the vertex kernel): 1 = general pipeline, 0 = tile pipeline; nullptr = all views general.
template <int S>
__global__ void __launch_bounds__(256)
clear_keys_kernel(unsigned long long* __restrict__ keys, int res, const int* __restrict__ route,
                  const int* __restrict__ boxes, const int* __restrict__ view_hard) {
  const int b = blockIdx.y;
  if (route != nullptr && !route[b]) return;
  const ViewBox vb = view_box(boxes, view_hard, b, res);
  if (vb.x0 > vb.x1 || vb.y0 > vb.y1) return;
  ulonglong2* kv = reinterpret_cast<ulonglong2*>(keys + size_t(b) * res * res * S);   // 16-byte stores
  constexpr int PER_PIXEL = S == 4 ? 2 : 1;                                           // ulonglong2 per pixel (S = 1: pixel pairs)
  if (S == 4) {
    const int bw = (vb.x1 - vb.x0 + 1) * PER_PIXEL, bh = vb.y1 - vb.y0 + 1;
    const int n = bw * bh;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const int y = i / bw, x = i - y * bw;
      kv[(size_t(vb.y0 + y) * res + vb.x0) * PER_PIXEL + x] = make_ulonglong2(~0ull, ~0ull);
    }
  } else {
    // one key per pixel: rows of the box widened to even pixel columns (res is a multiple of 4)
    const int xa = vb.x0 & ~1, xb = vb.x1 | 1;
    const int bw = (xb - xa + 1) / 2, bh = vb.y1 - vb.y0 + 1;
    const int n = bw * bh;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const int y = i / bw, x = i - y * bw;
      kv[(size_t(vb.y0 + y) * res + xa) / 2 + x] = make_ulonglong2(~0ull, ~0ull);
    }
  }
}

// camera description shared by the kernels that need camera-space positions again (hard triangles)
struct Camera {
  const float* verts;     // [V,3]
  const float* poses;     // [B,12]
  const float* view_k;    // [B,4] or nullptr
  float fx, fy, cx, cy;
  float znear, zfar;
};

__device__ __forceinline__ void view_intrinsics(const Camera& cam, int b, float& fx, float& fy, float& cx, float& cy) {
  fx = cam.fx; fy = cam.fy; cx = cam.cx; cy = cam.cy;
  if (cam.view_k != nullptr) {
    fx = cam.view_k[4 * b]; fy = cam.view_k[4 * b + 1]; cx = cam.view_k[4 * b + 2]; cy = cam.view_k[4 * b + 3];
  }
}

// cam = R * v + t, evaluated as ((r0*x + r1*y) + r2*z) + t
__device__ __forceinline__ void camera_point(const float* __restrict__ verts, const float* __restrict__ P, int i,
                                             float out[3]) {
  const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
  out[0] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[0], x), __fmul_rn(P[1], y)), __fmul_rn(P[2], z)), P[3]);
  out[1] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[4], x), __fmul_rn(P[5], y)), __fmul_rn(P[6], z)), P[7]);
  out[2] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[8], x), __fmul_rn(P[9], y)), __fmul_rn(P[10], z)), P[11]);
}

__global__ void __launch_bounds__(256)
vertex_kernel(const float* __restrict__ verts, const float* __restrict__ poses, ScreenVertex* __restrict__ sv,
              int V, int B, float fx, float fy, float cx, float cy, const float* __restrict__ view_k, float ZNEAR,
              float ZFAR, int* __restrict__ view_hard, int* __restrict__ boxes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  int bx0 = 0, by0 = 0, bx1 = 0, by1 = 0;   // this vertex's contribution to the view's screen box (0 = none)
  if (i < V) {
  if (view_k != nullptr) {   // per-view intrinsics (the refiner renders every frame at its own cropped K)
    fx = view_k[4 * b]; fy = view_k[4 * b + 1]; cx = view_k[4 * b + 2]; cy = view_k[4 * b + 3];
  }
  const float* P = poses + size_t(b) * 12;
  const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
  // cam = R * v + t, evaluated as ((r0*x + r1*y) + r2*z) + t
  const float X = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[0], x), __fmul_rn(P[1], y)), __fmul_rn(P[2], z)), P[3]);
  const float Y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[4], x), __fmul_rn(P[5], y)), __fmul_rn(P[6], z)), P[7]);
  const float Z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[8], x), __fmul_rn(P[9], y)), __fmul_rn(P[10], z)), P[11]);
  ScreenVertex o;
  if (!(Z > ZNEAR) || !(Z < ZFAR)) {
    o.x = INT_MIN; o.y = INT_MIN; o.z = 0.f; o.iz = 0.f;
  } else {
    const float u = __fadd_rn(__fdiv_rn(__fmul_rn(fx, X), Z), cx);
    const float v = __fadd_rn(__fdiv_rn(__fmul_rn(fy, Y), Z), cy);
    const float uf = floorf(__fadd_rn(__fmul_rn(u, float(ONE)), 0.5f));
    const float vf = floorf(__fadd_rn(__fmul_rn(v, float(ONE)), 0.5f));
    if (!(fabsf(uf) < float(COORD_LIMIT)) || !(fabsf(vf) < float(COORD_LIMIT))) {
      o.x = INT_MIN; o.y = INT_MIN; o.z = 0.f; o.iz = 0.f;
    } else {
      o.x = int(uf); o.y = int(vf); o.z = Z; o.iz = __fdiv_rn(1.0f, Z);
    }
  }
  if (o.x == INT_MIN) {
    view_hard[b] = 1;   // this view has triangles for the homogeneous path (benign race: all write 1)
  } else {
    const int px = o.x >> SUB, py = o.y >> SUB;
    bx0 = BOX_BIG - px; by0 = BOX_BIG - py; bx1 = px + 1 + BOX_BIG / 2; by1 = py + 1 + BOX_BIG / 2;
  }
  sv[size_t(b) * V + i] = o;
  }
  // one atomicMax per CTA and word ("max px + 1" is kept offset by BOX_BIG / 2 so that off-screen negatives stay positive;
  // per-warp atomics -- 320 per word and view -- cost 130 us of contention at 521 views)
  __shared__ int s_box[8][4];
  bx0 = __reduce_max_sync(0xffffffffu, bx0); by0 = __reduce_max_sync(0xffffffffu, by0);
  bx1 = __reduce_max_sync(0xffffffffu, bx1); by1 = __reduce_max_sync(0xffffffffu, by1);
  if ((threadIdx.x & 31) == 0) {
    int* w = s_box[threadIdx.x >> 5];
    w[0] = bx0; w[1] = by0; w[2] = bx1; w[3] = by1;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    int m = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) m = max(m, s_box[w][threadIdx.x]);
    if (m != 0) atomicMax(&boxes[4 * b + threadIdx.x], m);
  }
}

struct TriSetup {
  // edge functions E_i(p) = A_i * px + B_i * py + C_i (64-bit), fill-rule bias folded into C
  long long A[3], Bc[3], C[3];
  float z[3], iz[3];
  float area;
  int xmin, xmax, ymin, ymax;  // inclusive pixel bbox, already clipped
  int valid;
  int fits32;                  // unclipped extent < 64 px: edge functions relative to the box fit 32-bit integers
};

__device__ __forceinline__ bool setup_triangle(const ScreenVertex& a0, const ScreenVertex& a1, const ScreenVertex& a2,
                                               int res, int cull, TriSetup& t) {
  if (a0.x == INT_MIN || a1.x == INT_MIN || a2.x == INT_MIN) return false;
  ScreenVertex v0 = a0, v1 = a1, v2 = a2;
  long long area = (long long)(v1.x - v0.x) * (v2.y - v0.y) - (long long)(v2.x - v0.x) * (v1.y - v0.y);
  if (area == 0) return false;
  // image coordinates have y down: GL-front-facing (CCW in the y-up window) <=> area < 0 here
  if (cull && area > 0) return false;
  if (area < 0) { ScreenVertex tmp = v1; v1 = v2; v2 = tmp; area = -area; }
  const int minx = min(v0.x, min(v1.x, v2.x)), maxx = max(v0.x, max(v1.x, v2.x));
  const int miny = min(v0.y, min(v1.y, v2.y)), maxy = max(v0.y, max(v1.y, v2.y));
  // pixels whose [px, px+1) square can contain a sample inside [min, max]
  t.xmin = max(0, minx >> SUB); t.xmax = min(res - 1, maxx >> SUB);
  t.ymin = max(0, miny >> SUB); t.ymax = min(res - 1, maxy >> SUB);
  if (t.xmin > t.xmax || t.ymin > t.ymax) return false;
  t.fits32 = (maxx - minx) < (64 << SUB) && (maxy - miny) < (64 << SUB);
  const ScreenVertex* e0[3] = {&v1, &v2, &v0};  // edge i runs e0[i] -> e1[i]; weight i belongs to vertex i
  const ScreenVertex* e1[3] = {&v2, &v0, &v1};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const long long dx = e1[i]->x - e0[i]->x, dy = e1[i]->y - e0[i]->y;
    // E(p) = dx * (py - ay) - dy * (px - ax)  > 0 inside (orientation normalised above)
    t.A[i] = -dy;
    t.Bc[i] = dx;
    t.C[i] = dy * e0[i]->x - dx * e0[i]->y;
    const bool topleft = (dy < 0) || (dy == 0 && dx > 0);
    if (!topleft) t.C[i] -= 1;  // then "E >= 0" implements the top-left rule
  }
  t.z[0] = v0.z; t.z[1] = v1.z; t.z[2] = v2.z;
  t.iz[0] = v0.iz; t.iz[1] = v1.iz; t.iz[2] = v2.iz;
  t.area = __ll2float_rn(area);
  t.valid = 1;
  return true;
}

// Depth of the triangle at a sample from its three (biased) edge values -- perspective correct:
// 1/z is affine in screen space.
__device__ __forceinline__ float sample_depth(const TriSetup& t, long long e0, long long e1, long long e2,
                                              const long long* bias) {
  // z = area / (e0/z0 + e1/z1 + e2/z2): ONE division per sample (the form w_i = e_i / area, z = 1 / sum w_i/z_i costs
  // four, and the divisions were a quarter of the triangle kernel); oracle/raster_ref.c evaluates the same expression
  const float f0 = __ll2float_rn(e0 + bias[0]), f1 = __ll2float_rn(e1 + bias[1]), f2 = __ll2float_rn(e2 + bias[2]);
  const float den = __fadd_rn(__fadd_rn(__fmul_rn(f0, t.iz[0]), __fmul_rn(f1, t.iz[1])), __fmul_rn(f2, t.iz[2]));
  return __fdiv_rn(t.area, den);
}

template <int S>
__device__ __forceinline__ void raster_pixel(const TriSetup& t, const long long* bias, int px, int py,
                                             unsigned long long* __restrict__ keys_view, int res, unsigned face,
                                             float ZNEAR, float ZFAR) {
  const long long bx = (long long)px << SUB, by = (long long)py << SUB;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const long long sx = bx + c_sample_off[S == 4][s][0], sy = by + c_sample_off[S == 4][s][1];
    const long long e0 = t.A[0] * sx + t.Bc[0] * sy + t.C[0];
    const long long e1 = t.A[1] * sx + t.Bc[1] * sy + t.C[1];
    const long long e2 = t.A[2] * sx + t.Bc[2] * sy + t.C[2];
    if ((e0 | e1 | e2) >= 0) {
      const float z = sample_depth(t, e0, e1, e2, bias);
      if (z > ZNEAR && z < ZFAR) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(z) << 32) | face;
        atomicMin(&keys_view[(size_t(py) * res + px) * S + s], key);
      }
    }
  }
}

constexpr int BIG_TRI_PIXELS = 64;

template <int S>
__global__ void __launch_bounds__(256, 3)   // 80 registers: 3 CTAs / SM hide more latency than the 44 spilled bytes cost (-9 %)
triangle_kernel(const ScreenVertex* __restrict__ sv, const int* __restrict__ faces,
                unsigned long long* __restrict__ keys, int V, int F, int res, int cull, float ZNEAR, float ZFAR,
                const int* __restrict__ route, int big_pixels) {
  const int b = blockIdx.y;
  if (route != nullptr && !route[b]) return;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const ScreenVertex* svb = sv + size_t(b) * V;
  unsigned long long* keys_view = keys + size_t(b) * res * res * S;
  TriSetup t;
  t.valid = 0;
  long long bias[3] = {0, 0, 0};
  if (f < F) {
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    ScreenVertex v0 = svb[i0], v1 = svb[i1], v2 = svb[i2];
    const bool projects = v0.x != INT_MIN && v1.x != INT_MIN && v2.x != INT_MIN;   // (else: hard_triangle_kernel's business)
    const int minx = min(v0.x, min(v1.x, v2.x)), maxx = max(v0.x, max(v1.x, v2.x));
    const int miny = min(v0.y, min(v1.y, v2.y)), maxy = max(v0.y, max(v1.y, v2.y));
    const bool fits32 = projects && (maxx - minx) < (64 << SUB) && (maxy - miny) < (64 << SUB);
    const int xmin = max(0, minx >> SUB), xmax = min(res - 1, maxx >> SUB);
    const int ymin = max(0, miny >> SUB), ymax = min(res - 1, maxy >> SUB);
    const bool on_screen = xmin <= xmax && ymin <= ymax;
    if (fits32 && on_screen && (xmax - xmin + 1) * (ymax - ymin + 1) <= big_pixels) {
      // ---- FAST PATH (the common case: ~1 px of area).  With vertices less than 64 px apart every quantity of the set-up
      // and every edge value relative to the box origin is bounded by 2 * 2^14 * (2^14 + 2^8) < 2^31: the whole triangle
      // runs in 32-bit integers -- exactly the values of the 64-bit form (setup_triangle / raster_pixel / the oracle) --
      // and never builds the 64-bit TriSetup.
      int area = (v1.x - v0.x) * (v2.y - v0.y) - (v2.x - v0.x) * (v1.y - v0.y);
      // image coordinates have y down: GL-front-facing (CCW in the y-up window) <=> area < 0 here
      if (area != 0 && !(cull && area > 0)) {
        if (area < 0) { const ScreenVertex tmp = v1; v1 = v2; v2 = tmp; area = -area; }
        const float farea = __int2float_rn(area);
        const int ox = xmin << SUB, oy = ymin << SUB;
        const ScreenVertex* ea[3] = {&v1, &v2, &v0};   // edge i runs ea[i] -> eb[i]; weight i belongs to vertex i
        const ScreenVertex* eb[3] = {&v2, &v0, &v1};
        int ax[3], o_row[3], bi[3], by[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int dx = eb[i]->x - ea[i]->x, dy = eb[i]->y - ea[i]->y;
          const bool topleft = (dy < 0) || (dy == 0 && dx > 0);
          bi[i] = topleft ? 0 : 1;                      // "E >= 0" implements the top-left rule with the bias folded in
          ax[i] = -dy;
          by[i] = dx;
          o_row[i] = dx * (oy - ea[i]->y) - dy * (ox - ea[i]->x) - bi[i];   // value at the box origin
        }
        const float iz0 = v0.iz, iz1 = v1.iz, iz2 = v2.iz;
        // e_i(pixel, sample) = origin value + a_i dx + b_i dy: the sample offsets are constants, so every sample is the
        // pixel-origin value plus a per-triangle delta (3 adds instead of 6 multiply-adds; the same integers)
        int ds[3][S];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int s = 0; s < S; ++s) ds[i][s] = ax[i] * c_sample_off[S == 4][s][0] + by[i] * c_sample_off[S == 4][s][1];
        for (int py = ymin; py <= ymax; ++py) {
          int o0 = o_row[0], o1 = o_row[1], o2 = o_row[2];
          for (int px = xmin; px <= xmax; ++px, o0 += ax[0] << SUB, o1 += ax[1] << SUB, o2 += ax[2] << SUB) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const int e0 = o0 + ds[0][s];
              const int e1 = o1 + ds[1][s];
              const int e2 = o2 + ds[2][s];
              if ((e0 | e1 | e2) >= 0) {
                const float f0 = __int2float_rn(e0 + bi[0]), f1 = __int2float_rn(e1 + bi[1]), f2 = __int2float_rn(e2 + bi[2]);
                const float den = __fadd_rn(__fadd_rn(__fmul_rn(f0, iz0), __fmul_rn(f1, iz1)), __fmul_rn(f2, iz2));
                const float z = __fdiv_rn(farea, den);   // (sample_depth's expression)
                if (z > ZNEAR && z < ZFAR)
                  atomicMin(&keys_view[(size_t(py) * res + px) * S + s],
                            ((unsigned long long)__float_as_uint(z) << 32) | unsigned(f));
              }
            }
          }
          o_row[0] += by[0] << SUB; o_row[1] += by[1] << SUB; o_row[2] += by[2] << SUB;
        }
      }
    } else if (projects && on_screen && setup_triangle(v0, v1, v2, res, cull, t)) {
      // ---- everything else keeps the 64-bit set-up: large triangles (warp-cooperative walk below) and triangles that are
      // small on screen but long off screen (clipped by the viewport)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const long long dx = t.Bc[i], dy = -t.A[i];
        bias[i] = ((dy < 0) || (dy == 0 && dx > 0)) ? 0 : 1;  // undo the fill-rule bias for interpolation
      }
    }
  }
  const int w = t.valid ? (t.xmax - t.xmin + 1) : 0;
  const int h = t.valid ? (t.ymax - t.ymin + 1) : 0;
  const bool big = w * h > big_pixels;
  if (t.valid && !big) {
    // small on screen but long off screen (clipped by the viewport): 64-bit evaluation per sample
    for (int py = t.ymin; py <= t.ymax; ++py)
      for (int px = t.xmin; px <= t.xmax; ++px) raster_pixel<S>(t, bias, px, py, keys_view, res, unsigned(f), ZNEAR, ZFAR);
  }
  // large triangles: the whole warp walks the bounding box of one triangle at a time
  unsigned todo = __ballot_sync(0xffffffffu, big);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    TriSetup u;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      u.A[i] = __shfl_sync(0xffffffffu, t.A[i], src);
      u.Bc[i] = __shfl_sync(0xffffffffu, t.Bc[i], src);
      u.C[i] = __shfl_sync(0xffffffffu, t.C[i], src);
      u.z[i] = __shfl_sync(0xffffffffu, t.z[i], src);
      u.iz[i] = __shfl_sync(0xffffffffu, t.iz[i], src);
    }
    long long ub[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) ub[i] = __shfl_sync(0xffffffffu, bias[i], src);
    u.area = __shfl_sync(0xffffffffu, t.area, src);
    u.xmin = __shfl_sync(0xffffffffu, t.xmin, src);
    u.ymin = __shfl_sync(0xffffffffu, t.ymin, src);
    const int uw = __shfl_sync(0xffffffffu, w, src), uh = __shfl_sync(0xffffffffu, h, src);
    const unsigned uf = unsigned(__shfl_sync(0xffffffffu, f, src));
    for (int i = lane; i < uw * uh; i += 32) {
      const int py = u.ymin + i / uw, px = u.xmin + i % uw;
      raster_pixel<S>(u, ub, px, py, keys_view, res, uf, ZNEAR, ZFAR);
    }
  }
}

// GL_POINTS with point size 1 (pyrender.Mesh.from_points): the sprite is the square [x - 0.5, x + 0.5) x [y - 0.5, y + 0.5)
// around the projected vertex; every sample inside it takes the vertex depth.  Key payload = vertex index.
template <int S>
__global__ void __launch_bounds__(256)
point_kernel(const ScreenVertex* __restrict__ sv, unsigned long long* __restrict__ keys, int V, int res, float ZNEAR,
             float ZFAR) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const ScreenVertex v = sv[size_t(b) * V + i];
  if (v.x == INT_MIN) return;
  if (!(v.z > ZNEAR && v.z < ZFAR)) return;
  unsigned long long* keys_view = keys + size_t(b) * res * res * S;
  const int x0 = v.x - ONE / 2, y0 = v.y - ONE / 2;       // inclusive lower corner, exclusive upper = +ONE
  const int pxa = x0 >> SUB, pxb = (x0 + ONE - 1) >> SUB;
  const int pya = y0 >> SUB, pyb = (y0 + ONE - 1) >> SUB;
  const unsigned long long key = ((unsigned long long)__float_as_uint(v.z) << 32) | unsigned(i);
  for (int py = pya; py <= pyb; ++py) {
    if (py < 0 || py >= res) continue;
    for (int px = pxa; px <= pxb; ++px) {
      if (px < 0 || px >= res) continue;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int sx = (px << SUB) + c_sample_off[S == 4][s][0], sy = (py << SUB) + c_sample_off[S == 4][s][1];
        if (sx >= x0 && sx < x0 + ONE && sy >= y0 && sy < y0 + ONE)
          atomicMin(&keys_view[(size_t(py) * res + px) * S + s], key);
      }
    }
  }
}

// ---- hard triangles (see the header): homogeneous rasterisation; oracle/raster_ref.c hard_setup / hard_weights ----------
struct Hard {
  float n0[3], n1[3], n2[3];
  float adet, sgn;
  int xlo, xhi, ylo, yhi;
  int valid;
};

__device__ __forceinline__ void cross3(const float* a, const float* b, float* o) {
  o[0] = __fadd_rn(__fmul_rn(a[1], b[2]), -__fmul_rn(a[2], b[1]));
  o[1] = __fadd_rn(__fmul_rn(a[2], b[0]), -__fmul_rn(a[0], b[2]));
  o[2] = __fadd_rn(__fmul_rn(a[0], b[1]), -__fmul_rn(a[1], b[0]));
}

__device__ __noinline__ void hard_setup(const float* p0, const float* p1, const float* p2, float fx, float fy, float cx,
                                        float cy, float ZNEAR, float ZFAR, int res, int cull, Hard& h) {
  h.valid = 0;
  // entirely in front of the near plane or behind the far plane: nothing survives the per-sample depth test
  if (!(p0[2] > ZNEAR) && !(p1[2] > ZNEAR) && !(p2[2] > ZNEAR)) return;
  if (!(p0[2] < ZFAR) && !(p1[2] < ZFAR) && !(p2[2] < ZFAR)) return;
  cross3(p1, p2, h.n0); cross3(p2, p0, h.n1); cross3(p0, p1, h.n2);
  const float det = __fadd_rn(__fadd_rn(__fmul_rn(p0[0], h.n0[0]), __fmul_rn(p0[1], h.n0[1])), __fmul_rn(p0[2], h.n0[2]));
  if (!(det != 0.f) || !(det == det)) return;
  if (cull && det > 0.f) return;   // sign(det) = sign of the projected area: same rule as the fixed-point path
  h.sgn = det > 0.f ? 1.0f : -1.0f;
  h.adet = fabsf(det);
  // conservative pixel bounds: projections of the vertices in front of the near plane and of the edge / near-plane
  // intersections, +-1 px; anything not finite -> the whole viewport
  const float* v[3] = {p0, p1, p2};
  float umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY;
  bool finite = true;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float* a = v[i];
    const float* b = v[(i + 1) % 3];
    if (a[2] > ZNEAR) {
      const float u = __fadd_rn(__fdiv_rn(__fmul_rn(fx, a[0]), a[2]), cx), w = __fadd_rn(__fdiv_rn(__fmul_rn(fy, a[1]), a[2]), cy);
      umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, w); vmax = fmaxf(vmax, w);
      if (!(fabsf(u) < 1.0e9f) || !(fabsf(w) < 1.0e9f)) finite = false;
    }
    if ((a[2] > ZNEAR) != (b[2] > ZNEAR)) {
      const float t = __fdiv_rn(__fadd_rn(ZNEAR, -a[2]), __fadd_rn(b[2], -a[2]));
      const float X = __fadd_rn(a[0], __fmul_rn(t, __fadd_rn(b[0], -a[0]))), Y = __fadd_rn(a[1], __fmul_rn(t, __fadd_rn(b[1], -a[1])));
      const float u = __fadd_rn(__fdiv_rn(__fmul_rn(fx, X), ZNEAR), cx), w = __fadd_rn(__fdiv_rn(__fmul_rn(fy, Y), ZNEAR), cy);
      umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, w); vmax = fmaxf(vmax, w);
      if (!(fabsf(u) < 1.0e9f) || !(fabsf(w) < 1.0e9f)) finite = false;
    }
  }
  if (finite) {
    h.xlo = int(fmaxf(__fadd_rn(floorf(umin), -1.0f), 0.f)); h.xhi = int(fminf(__fadd_rn(floorf(umax), 1.0f), float(res - 1)));
    h.ylo = int(fmaxf(__fadd_rn(floorf(vmin), -1.0f), 0.f)); h.yhi = int(fminf(__fadd_rn(floorf(vmax), 1.0f), float(res - 1)));
  } else {
    h.xlo = 0; h.xhi = res - 1; h.ylo = 0; h.yhi = res - 1;
  }
  h.valid = h.xlo <= h.xhi && h.ylo <= h.yhi;
}

// weights of the fixed-point sample position (sx, sy); true if covered (b_i >= 0, sum > 0)
__device__ __forceinline__ bool hard_weights(const Hard& h, float fx, float fy, float cx, float cy, long long sx,
                                             long long sy, float b[3], float& sum) {
  const float dx = __fdiv_rn(__fadd_rn(__fmul_rn(__ll2float_rn(sx), 0.00390625f), -cx), fx);
  const float dy = __fdiv_rn(__fadd_rn(__fmul_rn(__ll2float_rn(sy), 0.00390625f), -cy), fy);
  b[0] = __fmul_rn(h.sgn, __fadd_rn(__fadd_rn(__fmul_rn(dx, h.n0[0]), __fmul_rn(dy, h.n0[1])), h.n0[2]));
  b[1] = __fmul_rn(h.sgn, __fadd_rn(__fadd_rn(__fmul_rn(dx, h.n1[0]), __fmul_rn(dy, h.n1[1])), h.n1[2]));
  b[2] = __fmul_rn(h.sgn, __fadd_rn(__fadd_rn(__fmul_rn(dx, h.n2[0]), __fmul_rn(dy, h.n2[1])), h.n2[2]));
  sum = __fadd_rn(__fadd_rn(b[0], b[1]), b[2]);
  return b[0] >= 0.f && b[1] >= 0.f && b[2] >= 0.f && sum > 0.f;
}

__device__ __forceinline__ float hard_interp(const float b[3], float sum, float a0, float a1, float a2) {
  return __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(b[0], a0), __fmul_rn(b[1], a1)), __fmul_rn(b[2], a2)), sum);
}

// One thread per (view, face) like triangle_kernel, but only views flagged by the vertex kernel do anything, and only
// faces with a vertex that did not project.  The warp then walks the pixel bounds of one such triangle at a time.
template <int S>
__global__ void __launch_bounds__(256)
hard_triangle_kernel(const ScreenVertex* __restrict__ sv, const int* __restrict__ faces, const Camera cam,
                     const int* __restrict__ view_hard, unsigned long long* __restrict__ keys, int V, int F, int res,
                     int cull) {
  const int b = blockIdx.y;
  if (!view_hard[b]) return;
  const int lane = threadIdx.x & 31;
  const ScreenVertex* svb = sv + size_t(b) * V;
  unsigned long long* keys_view = keys + size_t(b) * res * res * S;
  float fx, fy, cx, cy;
  view_intrinsics(cam, b, fx, fy, cx, cy);
  // (a small grid that loops over the faces: the CTAs of the views that have no such triangle -- all of them at the
  // benchmark pose -- cost next to nothing to retire)
  for (int blk = blockIdx.x; blk * int(blockDim.x) < F; blk += gridDim.x) {
  const int f = blk * blockDim.x + threadIdx.x;
  Hard h;
  h.valid = 0;
  if (f < F) {
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    if (svb[i0].x == INT_MIN || svb[i1].x == INT_MIN || svb[i2].x == INT_MIN) {
      const float* P = cam.poses + size_t(b) * 12;
      float p0[3], p1[3], p2[3];
      camera_point(cam.verts, P, i0, p0); camera_point(cam.verts, P, i1, p1); camera_point(cam.verts, P, i2, p2);
      hard_setup(p0, p1, p2, fx, fy, cx, cy, cam.znear, cam.zfar, res, cull, h);
    }
  }
  unsigned todo = __ballot_sync(0xffffffffu, h.valid != 0);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    Hard u;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      u.n0[i] = __shfl_sync(0xffffffffu, h.n0[i], src);
      u.n1[i] = __shfl_sync(0xffffffffu, h.n1[i], src);
      u.n2[i] = __shfl_sync(0xffffffffu, h.n2[i], src);
    }
    u.adet = __shfl_sync(0xffffffffu, h.adet, src);
    u.sgn = __shfl_sync(0xffffffffu, h.sgn, src);
    u.xlo = __shfl_sync(0xffffffffu, h.xlo, src); u.xhi = __shfl_sync(0xffffffffu, h.xhi, src);
    u.ylo = __shfl_sync(0xffffffffu, h.ylo, src); u.yhi = __shfl_sync(0xffffffffu, h.yhi, src);
    const unsigned uf = unsigned(__shfl_sync(0xffffffffu, f, src));
    const int uw = u.xhi - u.xlo + 1, uh = u.yhi - u.ylo + 1;
    for (int i = lane; i < uw * uh; i += 32) {
      const int py = u.ylo + i / uw, px = u.xlo + i % uw;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const long long sx = ((long long)px << SUB) + c_sample_off[S == 4][s][0];
        const long long sy = ((long long)py << SUB) + c_sample_off[S == 4][s][1];
        float bw[3], sum;
        if (hard_weights(u, fx, fy, cx, cy, sx, sy, bw, sum)) {
          const float z = __fdiv_rn(u.adet, sum);
          if (z > cam.znear && z < cam.zfar)
            atomicMin(&keys_view[(size_t(py) * res + px) * S + s], ((unsigned long long)__float_as_uint(z) << 32) | uf);
        }
      }
    }
  }
  }
}

struct Surface {
  const uint8_t* colors;    // [V,3] or nullptr
  const float* uv;          // [V,2] or nullptr
  const uint8_t* texture;   // RGBA8 mip chain or nullptr
  const float* srgb_lut;    // [65536] (i/65535)^2.2
  const uint8_t* gamma_lut; // [65536]
  int tex_w, tex_h, tex_levels;
  float ambient;            // scene ambient light (2 in renderer.py:53-55, 5 in tracking_refiner.py:34)
  float ambient_255;        // ambient / 255
};

// linear colour (already x ambient) -> unorm8 through the gamma LUT
__device__ __forceinline__ int to_unorm8(float lin, const uint8_t* __restrict__ lut) {
  lin = fminf(fmaxf(lin, 0.f), 1.f);
  if (!(lin == lin)) lin = 0.f;
  return lut[int(__fadd_rn(__fmul_rn(lin, 65535.0f), 0.5f))];
}

__device__ __forceinline__ int wrap_repeat(int i, int n) {
  int m = i % n;
  return m < 0 ? m + n : m;
}

// One bilinear tap of mip level `lvl` (REPEAT wrap); u, v in texture space with v up (image row 0 is v = 1).
__device__ __forceinline__ void bilinear(const Surface& sf, int lvl, float u, float v, float out[3]) {
  size_t off = 0;
  int W = sf.tex_w, H = sf.tex_h;
  for (int l = 0; l < lvl; ++l) {
    off += size_t(W) * H * 4;
    W = max(1, W >> 1); H = max(1, H >> 1);
  }
  const float x = __fadd_rn(__fmul_rn(u, float(W)), -0.5f);
  const float y = __fadd_rn(__fmul_rn(__fadd_rn(1.0f, -v), float(H)), -0.5f);
  const float xf = floorf(x), yf = floorf(y);
  const float fx = __fadd_rn(x, -xf), fy = __fadd_rn(y, -yf);
  // (coordinates far outside the int range only occur for degenerate UVs; clamp before the conversion)
  const int ix = int(fminf(fmaxf(xf, -1.0e9f), 1.0e9f)), iy = int(fminf(fmaxf(yf, -1.0e9f), 1.0e9f));
  const int x0 = wrap_repeat(ix, W), x1 = wrap_repeat(ix + 1, W);
  const int y0 = wrap_repeat(iy, H), y1 = wrap_repeat(iy + 1, H);
  const uchar4* t = reinterpret_cast<const uchar4*>(sf.texture + off);
  const uchar4 c00 = t[size_t(y0) * W + x0], c10 = t[size_t(y0) * W + x1];
  const uchar4 c01 = t[size_t(y1) * W + x0], c11 = t[size_t(y1) * W + x1];
  const float a00[3] = {float(c00.x), float(c00.y), float(c00.z)}, a10[3] = {float(c10.x), float(c10.y), float(c10.z)};
  const float a01[3] = {float(c01.x), float(c01.y), float(c01.z)}, a11[3] = {float(c11.x), float(c11.y), float(c11.z)};
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float top = __fadd_rn(a00[ch], __fmul_rn(fx, __fadd_rn(a10[ch], -a00[ch])));
    const float bot = __fadd_rn(a01[ch], __fmul_rn(fx, __fadd_rn(a11[ch], -a01[ch])));
    out[ch] = __fadd_rn(top, __fmul_rn(fy, __fadd_rn(bot, -top)));
  }
}

// Level of detail from the squared texel footprint rho^2: lambda = log2(rho), with log2 of the mantissa replaced by its
// chord (lambda = (e + m - 1) / 2 for rho^2 = m * 2^e, m in [1,2)): at most 0.043 levels off, and pure bit arithmetic,
// so the CPU oracle reproduces it exactly (libm's log2f is not specified to the last bit).
__device__ __forceinline__ float lod_from_rho2(float rho2, int levels) {
  const float top = float(levels - 1);
  if (!(rho2 < 1.0e30f)) return top;
  if (!(rho2 > 1.0f)) return 0.f;
  const unsigned bits = __float_as_uint(rho2);
  const float e = float(int(bits >> 23) - 127);
  const float m = __fmul_rn(float(bits & 0x7fffffu), 1.1920928955078125e-07f);   // mantissa fraction, exact
  return fminf(__fmul_rn(0.5f, __fadd_rn(e, m)), top);
}

// Shade triangle `face` at the centre of pixel (px, py): perspective-correct interpolation of vertex colours and / or
// texture coordinates, x2 ambient, gamma LUT -> unorm8.  MODE 0 = vertex colours, 1 = texture (times vertex colours
// when sf.colors is set).
// Interpolated attributes of a HARD triangle at a pixel centre: vertex colour vc[3], texture coordinate (u, v) and its
// finite differences to the neighbouring pixel centres (in texels).  Kept out of line: the common path must not pay
// registers for it.
__device__ __noinline__ void hard_attributes(const Camera& cam, int b, int res, const Surface& sf, int i0, int i1, int i2,
                                             int px, int py, bool want_colors, bool want_uv, float vc[3], float uvd[6]) {
  float fx, fy, cx, cy;
  view_intrinsics(cam, b, fx, fy, cx, cy);
  const float* P = cam.poses + size_t(b) * 12;
  float p0[3], p1[3], p2[3];
  camera_point(cam.verts, P, i0, p0); camera_point(cam.verts, P, i1, p1); camera_point(cam.verts, P, i2, p2);
  Hard h;
  hard_setup(p0, p1, p2, fx, fy, cx, cy, cam.znear, cam.zfar, res, 0, h);
  const long long sx = ((long long)px << SUB) + 128, sy = ((long long)py << SUB) + 128;
  float bw[3], sum;
  hard_weights(h, fx, fy, cx, cy, sx, sy, bw, sum);
  if (want_colors) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      vc[ch] = hard_interp(bw, sum, float(sf.colors[3 * i0 + ch]), float(sf.colors[3 * i1 + ch]), float(sf.colors[3 * i2 + ch]));
  }
  if (want_uv) {
    const float ua = sf.uv[2 * i0], ub = sf.uv[2 * i1], uc = sf.uv[2 * i2];
    const float va = sf.uv[2 * i0 + 1], vb = sf.uv[2 * i1 + 1], vcc = sf.uv[2 * i2 + 1];
    const float u = hard_interp(bw, sum, ua, ub, uc), v = hard_interp(bw, sum, va, vb, vcc);
    float bx[3], sumx, by[3], sumy;
    hard_weights(h, fx, fy, cx, cy, sx + ONE, sy, bx, sumx);
    hard_weights(h, fx, fy, cx, cy, sx, sy + ONE, by, sumy);
    const float fw = float(sf.tex_w), fh = float(sf.tex_h);
    uvd[0] = u; uvd[1] = v;
    uvd[2] = __fmul_rn(__fadd_rn(hard_interp(bx, sumx, ua, ub, uc), -u), fw);
    uvd[3] = __fmul_rn(__fadd_rn(hard_interp(bx, sumx, va, vb, vcc), -v), fh);
    uvd[4] = __fmul_rn(__fadd_rn(hard_interp(by, sumy, ua, ub, uc), -u), fw);
    uvd[5] = __fmul_rn(__fadd_rn(hard_interp(by, sumy, va, vb, vcc), -v), fh);
  }
}

// WITH_HARD = false: the common path (resolve_kernel).  A hard triangle -- which only resolve_hard_kernel can shade --
// yields black there and is overwritten afterwards; keeping that branch out of the common kernel keeps it at 48
// registers without a stack frame (with it: 104 registers + 256 B of local memory, +1.6 ms per 521 views).
template <int MODE, bool WITH_HARD>
__device__ __forceinline__ void shade(const ScreenVertex* __restrict__ svb, const int* __restrict__ faces,
                                      const Surface& sf, const Camera* cam, int b, int res, unsigned face, int px, int py,
                                      int out[3]) {
  const int i0 = faces[3 * face], i1 = faces[3 * face + 1], i2 = faces[3 * face + 2];
  ScreenVertex v0 = svb[i0], v1 = svb[i1], v2 = svb[i2];
  int c0 = i0, c1 = i1, c2 = i2;
  const bool hard = v0.x == INT_MIN || v1.x == INT_MIN || v2.x == INT_MIN;
  float h_vc[3], h_uvd[6];   // (only written and read on the hard path)
  if (hard) {
    if (!WITH_HARD) { out[0] = out[1] = out[2] = 0; return; }
    h_vc[0] = h_vc[1] = h_vc[2] = 255.f;
    hard_attributes(*cam, b, res, sf, i0, i1, i2, px, py, MODE == 0 || sf.colors != nullptr, MODE == 1, h_vc, h_uvd);
  }
  long long area = (long long)(v1.x - v0.x) * (v2.y - v0.y) - (long long)(v2.x - v0.x) * (v1.y - v0.y);
  if (area < 0) { ScreenVertex tmp = v1; v1 = v2; v2 = tmp; int ti = c1; c1 = c2; c2 = ti; area = -area; }
  // perspective weights at (sx, sy) (unbiased edge values: the point may lie outside the triangle)
  // Triangles whose vertices are < 64 px apart (all but a handful): the shading point lies within a pixel of the
  // triangle (one of its samples is covered), every coordinate difference is < 2^14 + 2^9 sub-pixels and every edge value
  // < 2^30: 32-bit integers give exactly the values of the 64-bit form, at a third of the instructions.
  const bool small = !hard &&
                     max(v0.x, max(v1.x, v2.x)) - min(v0.x, min(v1.x, v2.x)) < (64 << SUB) &&
                     max(v0.y, max(v1.y, v2.y)) - min(v0.y, min(v1.y, v2.y)) < (64 << SUB);
  auto weights = [&](long long sx, long long sy, float& w0, float& w1, float& w2, float& wsum) {
    float f0, f1, f2;
    if (small) {
      const int x = int(sx), y = int(sy);
      f0 = __int2float_rn((v2.x - v1.x) * (y - v1.y) - (v2.y - v1.y) * (x - v1.x));
      f1 = __int2float_rn((v0.x - v2.x) * (y - v2.y) - (v0.y - v2.y) * (x - v2.x));
      f2 = __int2float_rn((v1.x - v0.x) * (y - v0.y) - (v1.y - v0.y) * (x - v0.x));
    } else {
      f0 = __ll2float_rn((long long)(v2.x - v1.x) * (sy - v1.y) - (long long)(v2.y - v1.y) * (sx - v1.x));
      f1 = __ll2float_rn((long long)(v0.x - v2.x) * (sy - v2.y) - (long long)(v0.y - v2.y) * (sx - v2.x));
      f2 = __ll2float_rn((long long)(v1.x - v0.x) * (sy - v0.y) - (long long)(v1.y - v0.y) * (sx - v0.x));
    }
    // w_i = e_i / z_i (the common factor 1 / area cancels in the ratio below); `wsum` returns 1 / sum w_i: ONE division
    // per evaluation point instead of three for the weights and one per interpolated attribute
    w0 = __fmul_rn(f0, v0.iz);
    w1 = __fmul_rn(f1, v1.iz);
    w2 = __fmul_rn(f2, v2.iz);
    wsum = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(w0, w1), w2));
  };
  auto interp = [&](float w0, float w1, float w2, float rsum, float a0, float a1, float a2) {
    return __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(w0, a0), __fmul_rn(w1, a1)), __fmul_rn(w2, a2)), rsum);
  };
  const long long sx = ((long long)px << SUB) + 128, sy = ((long long)py << SUB) + 128;
  float w0, w1, w2, wsum;
  weights(sx, sy, w0, w1, w2, wsum);
  float vc[3] = {255.f, 255.f, 255.f};
  if (WITH_HARD && hard) {
    vc[0] = h_vc[0]; vc[1] = h_vc[1]; vc[2] = h_vc[2];
  } else if (MODE == 0 || sf.colors != nullptr) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      vc[ch] = interp(w0, w1, w2, wsum, float(sf.colors[3 * c0 + ch]), float(sf.colors[3 * c1 + ch]),
                      float(sf.colors[3 * c2 + ch]));
  }
  if (MODE == 0) {
    // base colour in [0,1] is c/255; ambient (a,a,a): linear = a*c/255, clamped; LUT index = round(linear*65535)
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) out[ch] = to_unorm8(__fmul_rn(vc[ch], sf.ambient_255), sf.gamma_lut);
    return;
  }
  const float ua = sf.uv[2 * c0], ub = sf.uv[2 * c1], uc = sf.uv[2 * c2];
  const float va = sf.uv[2 * c0 + 1], vb = sf.uv[2 * c1 + 1], vcc = sf.uv[2 * c2 + 1];
  float u = interp(w0, w1, w2, wsum, ua, ub, uc), v = interp(w0, w1, w2, wsum, va, vb, vcc);
  // footprint from the neighbouring pixel centres (the finite differences GL takes inside a 2x2 quad)
  float x0w, x1w, x2w, xs, y0w, y1w, y2w, ys;
  weights(sx + ONE, sy, x0w, x1w, x2w, xs);
  weights(sx, sy + ONE, y0w, y1w, y2w, ys);
  const float fw = float(sf.tex_w), fh = float(sf.tex_h);
  float dux = __fmul_rn(__fadd_rn(interp(x0w, x1w, x2w, xs, ua, ub, uc), -u), fw);
  float dvx = __fmul_rn(__fadd_rn(interp(x0w, x1w, x2w, xs, va, vb, vcc), -v), fh);
  float duy = __fmul_rn(__fadd_rn(interp(y0w, y1w, y2w, ys, ua, ub, uc), -u), fw);
  float dvy = __fmul_rn(__fadd_rn(interp(y0w, y1w, y2w, ys, va, vb, vcc), -v), fh);
  if (WITH_HARD && hard) { u = h_uvd[0]; v = h_uvd[1]; dux = h_uvd[2]; dvx = h_uvd[3]; duy = h_uvd[4]; dvy = h_uvd[5]; }
  const float rx = __fadd_rn(__fmul_rn(dux, dux), __fmul_rn(dvx, dvx));
  const float ry = __fadd_rn(__fmul_rn(duy, duy), __fmul_rn(dvy, dvy));
  const float lod = lod_from_rho2(fmaxf(rx, ry), sf.tex_levels);
  const int l0 = int(lod);
  const float t = __fadd_rn(lod, -float(l0));
  float ca[3];
  bilinear(sf, l0, u, v, ca);
  if (t > 0.f) {
    float cb[3];
    bilinear(sf, min(l0 + 1, sf.tex_levels - 1), u, v, cb);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) ca[ch] = __fadd_rn(ca[ch], __fmul_rn(t, __fadd_rn(cb[ch], -ca[ch])));
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float cn = __fdiv_rn(ca[ch], 255.0f);
    cn = fminf(fmaxf(cn, 0.f), 1.f);
    if (!(cn == cn)) cn = 0.f;
    float lin = sf.srgb_lut[int(__fadd_rn(__fmul_rn(cn, 65535.0f), 0.5f))];           // sRGB -> linear after filtering
    if (sf.colors != nullptr) lin = __fmul_rn(lin, __fdiv_rn(vc[ch], 255.0f));         // COLOR_0 multiplier
    out[ch] = to_unorm8(__fmul_rn(lin, sf.ambient), sf.gamma_lut);
  }
}

// MODE 0 / 1: triangles (key payload = face), 2: points (key payload = vertex, flat colour).
// One thread per pixel, a warp = 32 consecutive pixels of a view: covered pixels come in runs along a row, so warps are
// mostly all-covered or all-background (the earlier 4-pixels-per-thread mapping spread a warp over 128 pixels and ran with
// 12 of 32 lanes active).  Depth goes out as one coalesced float per lane; the 96 RGB bytes of a warp are assembled into 24
// words with two shuffles per lane.
template <int S, int MODE>
__global__ void __launch_bounds__(256)
resolve_kernel(const unsigned long long* __restrict__ keys, const ScreenVertex* __restrict__ sv,
               const int* __restrict__ faces, const Surface sf, uint8_t* __restrict__ rgb, float* __restrict__ depth,
               int V, int res, const int* __restrict__ route, const int* __restrict__ boxes,
               const int* __restrict__ view_hard) {
  const int b = blockIdx.y;
  if (route != nullptr && !route[b]) return;
  const int npix = res * res;
  const ViewBox vb = view_box(boxes, view_hard, b, res);   // outside: background, the sample keys are not even read
  // (the grid may be smaller than the image: a routed launch keeps it small so that the CTAs of views that are not
  // its business cost nothing to retire)
  for (int blk = blockIdx.x; blk * int(blockDim.x) < npix; blk += gridDim.x) {
  const int p = blk * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp_base = p - lane;
  if (warp_base >= npix) return;
  const bool live = p < npix;
  const int py = live ? p / res : 0, px = live ? p - py * res : 0;
  const ScreenVertex* svb = sv + size_t(b) * V;
  int acc[3] = {0, 0, 0};
  float dep = 0.f;
  if (live && !(px >= vb.x0 && px <= vb.x1 && py >= vb.y0 && py <= vb.y1)) depth[size_t(b) * npix + p] = 0.f;
  if (live && px >= vb.x0 && px <= vb.x1 && py >= vb.y0 && py <= vb.y1) {
    const unsigned long long* kv = keys + (size_t(b) * npix + p) * S;
    unsigned long long k[S];
    if (S == 4) {
      const ulonglong2 k01 = *reinterpret_cast<const ulonglong2*>(kv);
      const ulonglong2 k23 = *reinterpret_cast<const ulonglong2*>(kv + 2);
      k[0] = k01.x; k[1] = k01.y; k[2 % S] = k23.x; k[3 % S] = k23.y;
    } else {
      k[0] = kv[0];
    }
    unsigned last_face = 0xffffffffu;
    int col[3] = {0, 0, 0};
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (k[s] != ~0ull) {
        const unsigned face = unsigned(k[s] & 0xffffffffu);
        if (face != last_face) {
          if (MODE == 2) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
              col[ch] = to_unorm8(__fmul_rn(float(sf.colors[3 * face + ch]), sf.ambient_255), sf.gamma_lut);
          } else {
            shade<MODE, false>(svb, faces, sf, nullptr, b, res, face, px, py, col);
          }
          last_face = face;
        }
        acc[0] += col[0]; acc[1] += col[1]; acc[2] += col[2];
      }
    }
    if (S == 4) { acc[0] = (acc[0] + 2) >> 2; acc[1] = (acc[1] + 2) >> 2; acc[2] = (acc[2] + 2) >> 2; }
    dep = (k[0] != ~0ull) ? __uint_as_float(unsigned(k[0] >> 32)) : 0.f;
    depth[size_t(b) * npix + p] = dep;
  }
  const unsigned c24 = unsigned(acc[0]) | (unsigned(acc[1]) << 8) | (unsigned(acc[2]) << 16);
  uint8_t* out = rgb + (size_t(b) * npix + warp_base) * 3;      // 4-byte aligned: npix and warp_base are multiples of 4
  if (warp_base + 32 <= npix) {
    // word w (0..23) holds bytes 4w..4w+3 = the tail of pixel a = 4w/3 and the head of pixel a+1
    const int a = (4 * lane) / 3, o = (4 * lane) - 3 * a;
    const unsigned ca = __shfl_sync(0xffffffffu, c24, a & 31);
    const unsigned cb = __shfl_sync(0xffffffffu, c24, (a + 1) & 31);
    if (lane < 24) reinterpret_cast<uint32_t*>(out)[lane] = (ca >> (8 * o)) | (cb << (24 - 8 * o));
  } else if (live) {
    out[3 * lane] = uint8_t(acc[0]); out[3 * lane + 1] = uint8_t(acc[1]); out[3 * lane + 2] = uint8_t(acc[2]);
  }
  }
}

// Second resolve pass for the views flagged by the vertex kernel: pixels with at least one sample won by a hard triangle
// are shaded again (hard triangles through hard_attributes, the others as before) and their RGB overwritten.  Depth is
// already final.  Views without such vertices return at once.
template <int S, int MODE>
__global__ void __launch_bounds__(256)
resolve_hard_kernel(const unsigned long long* __restrict__ keys, const ScreenVertex* __restrict__ sv,
                    const int* __restrict__ faces, const Surface sf, const Camera cam, const int* __restrict__ view_hard,
                    uint8_t* __restrict__ rgb, int V, int res) {
  const int b = blockIdx.y;
  if (!view_hard[b]) return;
  const int npix = res * res;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
  const int py = p / res, px = p - py * res;
  const ScreenVertex* svb = sv + size_t(b) * V;
  const unsigned long long* kv = keys + (size_t(b) * npix + p) * S;
  bool any_hard = false;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const unsigned long long k = kv[s];
    if (k != ~0ull) {
      const unsigned face = unsigned(k & 0xffffffffu);
      any_hard |= svb[faces[3 * face]].x == INT_MIN || svb[faces[3 * face + 1]].x == INT_MIN || svb[faces[3 * face + 2]].x == INT_MIN;
    }
  }
  if (!any_hard) continue;
  int acc[3] = {0, 0, 0};
  unsigned last_face = 0xffffffffu;
  int col[3] = {0, 0, 0};
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const unsigned long long k = kv[s];
    if (k != ~0ull) {
      const unsigned face = unsigned(k & 0xffffffffu);
      if (face != last_face) {
        shade<MODE, true>(svb, faces, sf, &cam, b, res, face, px, py, col);
        last_face = face;
      }
      acc[0] += col[0]; acc[1] += col[1]; acc[2] += col[2];
    }
  }
  if (S == 4) { acc[0] = (acc[0] + 2) >> 2; acc[1] = (acc[1] + 2) >> 2; acc[2] = (acc[2] + 2) >> 2; }
  uint8_t* o = rgb + (size_t(b) * npix + p) * 3;
  o[0] = uint8_t(acc[0]); o[1] = uint8_t(acc[1]); o[2] = uint8_t(acc[2]);
  }
}

// ================================================================================================================
// TILE pipeline
// ================================================================================================================
constexpr int TILE = 16;
constexpr int TILE_SHIFT = 4;

// pixel box of a projected triangle, exactly setup_triangle's (including its rejections)
__device__ __forceinline__ bool face_box(const ScreenVertex& v0, const ScreenVertex& v1, const ScreenVertex& v2, int res,
                                         int cull, int& xmin, int& xmax, int& ymin, int& ymax) {
  const long long area = (long long)(v1.x - v0.x) * (v2.y - v0.y) - (long long)(v2.x - v0.x) * (v1.y - v0.y);
  if (area == 0) return false;
  if (cull && area > 0) return false;
  const int minx = min(v0.x, min(v1.x, v2.x)), maxx = max(v0.x, max(v1.x, v2.x));
  const int miny = min(v0.y, min(v1.y, v2.y)), maxy = max(v0.y, max(v1.y, v2.y));
  xmin = max(0, minx >> SUB); xmax = min(res - 1, maxx >> SUB);
  ymin = max(0, miny >> SUB); ymax = min(res - 1, maxy >> SUB);
  return xmin <= xmax && ymin <= ymax;
}

// PASS 0: count the triangles of every tile.  PASS 1: write them (tile_count is then the zeroed cursor array); triangles
// spanning more than 2 x 2 tiles go to the view's own list with their pixel box.
template <int PASS>
__global__ void __launch_bounds__(256)
bin_kernel(const ScreenVertex* __restrict__ sv, const int* __restrict__ faces, const int* __restrict__ route, int V, int F,
           int res, int cull, int ntx, int nty, int* __restrict__ tile_count, const int* __restrict__ tile_offset,
           int* __restrict__ bin_list, int* __restrict__ big_count, int4* __restrict__ big_list) {
  const int b = blockIdx.y;
  if (route[b]) return;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const ScreenVertex* svb = sv + size_t(b) * V;
  const ScreenVertex v0 = svb[faces[3 * f]], v1 = svb[faces[3 * f + 1]], v2 = svb[faces[3 * f + 2]];
  int xmin, xmax, ymin, ymax;
  if (!face_box(v0, v1, v2, res, cull, xmin, xmax, ymin, ymax)) return;
  const int tx0 = xmin >> TILE_SHIFT, tx1 = xmax >> TILE_SHIFT, ty0 = ymin >> TILE_SHIFT, ty1 = ymax >> TILE_SHIFT;
  if (tx1 - tx0 <= 1 && ty1 - ty0 <= 1) {
    for (int ty = ty0; ty <= ty1; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) {
        const int tile = (b * nty + ty) * ntx + tx;
        const int pos = atomicAdd(&tile_count[tile], 1);
        if (PASS == 1) bin_list[tile_offset[tile] + pos] = f;
      }
  } else if (PASS == 1) {
    const int pos = atomicAdd(&big_count[b], 1);
    big_list[size_t(b) * F + pos] = make_int4(f, xmin | (xmax << 16), ymin | (ymax << 16), 0);
  }
}

// exclusive prefix sum of n counters (one CTA; n is views x tiles, ~1e5): offset[i] = sum(count[0..i)), offset[n] = total
__global__ void __launch_bounds__(1024)
tile_scan_kernel(const int* __restrict__ count, int* __restrict__ offset, int n) {
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + 1023) / 1024;
  const int lo = min(n, tid * per), hi = min(n, lo + per);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += count[i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  int run = incl - sum + (warp > 0 ? warp_tot[warp - 1] : 0);
  for (int i = lo; i < hi; ++i) { offset[i] = run; run += count[i]; }
  if (tid == 1023) offset[n] = warp_tot[31];
}

// every sample of pixels [xa, xb] x [ya, yb] against triangle t; the keys live at slot((px, py)) = base + ((py - oy) * pitch
// + (px - ox)) * S (shared memory here)
template <int S>
__device__ __forceinline__ void raster_box(const TriSetup& t, const long long* bias, int xa, int xb, int ya, int yb,
                                           unsigned long long* base, int pitch, int ox, int oy, unsigned face, float ZNEAR,
                                           float ZFAR) {
  if (t.fits32) {
    // (see triangle_kernel: vertices less than 64 px apart -> the edge functions relative to the box origin fit 32 bits)
    const long long bx = (long long)t.xmin << SUB, by = (long long)t.ymin << SUB;
    int a[3], b[3], c[3], bi[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      a[i] = int(t.A[i]); b[i] = int(t.Bc[i]);
      c[i] = int(t.A[i] * bx + t.Bc[i] * by + t.C[i]);
      bi[i] = int(bias[i]);
    }
    for (int py = ya; py <= yb; ++py) {
      const int dy0 = (py - t.ymin) << SUB;
      for (int px = xa; px <= xb; ++px) {
        const int dx0 = (px - t.xmin) << SUB;
        unsigned long long* slot = base + (size_t(py - oy) * pitch + (px - ox)) * S;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int dx = dx0 + c_sample_off[S == 4][s][0], dy = dy0 + c_sample_off[S == 4][s][1];
          const int e0 = a[0] * dx + b[0] * dy + c[0];
          const int e1 = a[1] * dx + b[1] * dy + c[1];
          const int e2 = a[2] * dx + b[2] * dy + c[2];
          if ((e0 | e1 | e2) >= 0) {
            const float f0 = __int2float_rn(e0 + bi[0]), f1 = __int2float_rn(e1 + bi[1]), f2 = __int2float_rn(e2 + bi[2]);
            const float den = __fadd_rn(__fadd_rn(__fmul_rn(f0, t.iz[0]), __fmul_rn(f1, t.iz[1])), __fmul_rn(f2, t.iz[2]));
            const float z = __fdiv_rn(t.area, den);   // (sample_depth's expression)
            if (z > ZNEAR && z < ZFAR) atomicMin(&slot[s], ((unsigned long long)__float_as_uint(z) << 32) | face);
          }
        }
      }
    }
  } else {
    for (int py = ya; py <= yb; ++py)
      for (int px = xa; px <= xb; ++px) {
        unsigned long long* slot = base + (size_t(py - oy) * pitch + (px - ox)) * S;
        const long long bx = (long long)px << SUB, by = (long long)py << SUB;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const long long sx = bx + c_sample_off[S == 4][s][0], sy = by + c_sample_off[S == 4][s][1];
          const long long e0 = t.A[0] * sx + t.Bc[0] * sy + t.C[0];
          const long long e1 = t.A[1] * sx + t.Bc[1] * sy + t.C[1];
          const long long e2 = t.A[2] * sx + t.Bc[2] * sy + t.C[2];
          if ((e0 | e1 | e2) >= 0) {
            const float z = sample_depth(t, e0, e1, e2, bias);
            if (z > ZNEAR && z < ZFAR) atomicMin(&slot[s], ((unsigned long long)__float_as_uint(z) << 32) | face);
          }
        }
      }
  }
}

__device__ __forceinline__ void fill_rule_bias(const TriSetup& t, long long (&bias)[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const long long dx = t.Bc[i], dy = -t.A[i];
    bias[i] = ((dy < 0) || (dy == 0 && dx > 0)) ? 0 : 1;   // undo the fill-rule bias for interpolation
  }
}

// One CTA per (tile, view): depth test in shared memory, then shade + write.  Outputs of tiles without triangles were
// zeroed by a memset.  MODE 0 = vertex colours, 1 = texture.
template <int S, int MODE, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS)
tile_kernel(const ScreenVertex* __restrict__ sv, const int* __restrict__ faces, const Surface sf,
            const int* __restrict__ route, const int* __restrict__ tile_offset, const int* __restrict__ bin_list,
            const int* __restrict__ big_count, const int4* __restrict__ big_list, uint8_t* __restrict__ rgb,
            float* __restrict__ depth, int V, int F, int res, int ntx, int nty, int cull, float ZNEAR, float ZFAR) {
  const int b = blockIdx.y;
  if (route[b]) return;
  const int tile = blockIdx.x;
  const int gt = b * ntx * nty + tile;
  const int beg = tile_offset[gt], n_small = tile_offset[gt + 1] - beg, n_big = big_count[b];
  const int ty = tile / ntx, tx = tile - ty * ntx;
  const int px0 = tx << TILE_SHIFT, py0 = ty << TILE_SHIFT;
  const int px1 = min(px0 + TILE - 1, res - 1), py1 = min(py0 + TILE - 1, res - 1);
  // does any spanning triangle of the view touch this tile?  (uniform over the CTA)
  __shared__ unsigned long long skeys[TILE * TILE * S];
  __shared__ int s_any_big;
  if (threadIdx.x == 0) s_any_big = 0;
  for (int i = threadIdx.x; i < TILE * TILE * S; i += blockDim.x) skeys[i] = ~0ull;
  __syncthreads();
  const int4* bl = big_list + size_t(b) * F;
  for (int e = threadIdx.x; e < n_big; e += blockDim.x) {
    const int4 en = bl[e];
    const int xmin = en.y & 0xffff, xmax = en.y >> 16, ymin = en.z & 0xffff, ymax = en.z >> 16;
    if (xmin <= px1 && xmax >= px0 && ymin <= py1 && ymax >= py0) s_any_big = 1;
  }
  __syncthreads();
  if (n_small == 0 && !s_any_big) return;
  const ScreenVertex* svb = sv + size_t(b) * V;
  // ---- binned triangles: one per thread, clipped to the tile
  for (int e = threadIdx.x; e < n_small; e += blockDim.x) {
    const int f = bin_list[beg + e];
    TriSetup t;
    if (!setup_triangle(svb[faces[3 * f]], svb[faces[3 * f + 1]], svb[faces[3 * f + 2]], res, cull, t)) continue;
    long long bias[3];
    fill_rule_bias(t, bias);
    raster_box<S>(t, bias, max(t.xmin, px0), min(t.xmax, px1), max(t.ymin, py0), min(t.ymax, py1), skeys, TILE, px0, py0,
                  unsigned(f), ZNEAR, ZFAR);
  }
  // ---- spanning triangles: the CTA takes them one at a time, a thread per pixel of the tile
  if (s_any_big) {
    const int lx = threadIdx.x & (TILE - 1), ly = threadIdx.x >> TILE_SHIFT;
    const int px = px0 + lx, py = py0 + ly;
    for (int e = 0; e < n_big; ++e) {
      const int4 en = bl[e];
      const int xmin = en.y & 0xffff, xmax = en.y >> 16, ymin = en.z & 0xffff, ymax = en.z >> 16;
      if (!(xmin <= px1 && xmax >= px0 && ymin <= py1 && ymax >= py0)) continue;     // uniform
      if (px < xmin || px > xmax || py < ymin || py > ymax || px >= res || py >= res) continue;
      const int f = en.x;
      TriSetup t;
      if (!setup_triangle(svb[faces[3 * f]], svb[faces[3 * f + 1]], svb[faces[3 * f + 2]], res, cull, t)) continue;
      long long bias[3];
      fill_rule_bias(t, bias);
      t.fits32 = 0;                                   // (a spanning triangle never fits)
      raster_box<S>(t, bias, px, px, py, py, skeys, TILE, px0, py0, unsigned(f), ZNEAR, ZFAR);
    }
  }
  __syncthreads();
  // ---- resolve: a thread per pixel
  const int lx = threadIdx.x & (TILE - 1), ly = threadIdx.x >> TILE_SHIFT;
  const int px = px0 + lx, py = py0 + ly;
  if (px >= res || py >= res) return;
  const unsigned long long* k = skeys + threadIdx.x * S;
  int acc[3] = {0, 0, 0};
  unsigned last_face = 0xffffffffu;
  int col[3] = {0, 0, 0};
#pragma unroll
  for (int s = 0; s < S; ++s) {
    if (k[s] != ~0ull) {
      const unsigned face = unsigned(k[s] & 0xffffffffu);
      if (face != last_face) {
        shade<MODE, false>(svb, faces, sf, nullptr, b, res, face, px, py, col);
        last_face = face;
      }
      acc[0] += col[0]; acc[1] += col[1]; acc[2] += col[2];
    }
  }
  if (S == 4) { acc[0] = (acc[0] + 2) >> 2; acc[1] = (acc[1] + 2) >> 2; acc[2] = (acc[2] + 2) >> 2; }
  const size_t pix = (size_t(b) * res + py) * res + px;
  depth[pix] = (k[0] != ~0ull) ? __uint_as_float(unsigned(k[0] >> 32)) : 0.f;
  uint8_t* o = rgb + pix * 3;
  o[0] = uint8_t(acc[0]); o[1] = uint8_t(acc[1]); o[2] = uint8_t(acc[2]);
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// workspace carve-up shared by raster_workspace_bytes and rasterize
struct RasterWorkspace {
  size_t sv, keys, flags, counters, offsets, bins, big, total;
  int ntx, nty;
};
RasterWorkspace raster_layout(int B, int V, int F, int res, int msaa) {
  RasterWorkspace w;
  w.ntx = (res + TILE - 1) / TILE; w.nty = w.ntx;
  const size_t ntiles = size_t(B) * w.ntx * w.nty;
  size_t off = 0;
  w.sv = off;       off += align_up(size_t(B) * V * sizeof(ScreenVertex), 256);
  w.keys = off;     off += align_up(size_t(B) * res * res * msaa * 8, 256);          // general pipeline only
  w.flags = off;    off += align_up(size_t(B) * 5 * sizeof(int) + 16, 256);           // per-view route | per-view screen box (int4)
  w.counters = off; off += align_up((2 * ntiles + size_t(B)) * sizeof(int), 256);     // tile counts | tile cursors | big counts
  w.offsets = off;  off += align_up((ntiles + 1) * sizeof(int), 256);
  w.bins = off;     off += align_up(size_t(B) * size_t(F > 0 ? F : 0) * 4 * sizeof(int), 256);   // <= 4 tiles per binned triangle
  w.big = off;      off += align_up(size_t(B) * size_t(F > 0 ? F : 0) * sizeof(int4), 256);
  w.total = off + 256;
  return w;
}

template <int S>
int launch_raster(const RasterArgs& a, uint8_t* ws, const RasterWorkspace& w, float znear, float zfar, cudaStream_t stream) {
  ScreenVertex* sv = reinterpret_cast<ScreenVertex*>(ws + w.sv);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws + w.keys);
  int* view_hard = reinterpret_cast<int*>(ws + w.flags);
  const int* boxes = reinterpret_cast<const int*>((reinterpret_cast<uintptr_t>(view_hard + a.B) + 15) & ~uintptr_t(15));
  Camera cam;
  cam.verts = a.verts; cam.poses = a.poses; cam.view_k = a.view_k;
  cam.fx = a.fx; cam.fy = a.fy; cam.cx = a.cx; cam.cy = a.cy; cam.znear = znear; cam.zfar = zfar;
  Surface sf;
  sf.ambient = a.ambient > 0.f ? a.ambient : 2.0f;
  sf.ambient_255 = sf.ambient / 255.0f;
  sf.colors = a.colors; sf.uv = a.uv; sf.texture = a.texture; sf.srgb_lut = a.srgb_lut; sf.gamma_lut = a.gamma_lut;
  sf.tex_w = a.tex_w; sf.tex_h = a.tex_h; sf.tex_levels = a.tex_levels;
  const size_t per_view = size_t(a.res) * a.res * S;
  const dim3 rgrid((a.res * a.res + 255) / 256, a.B);
  const dim3 cgrid(unsigned((per_view + 256 * 16 - 1) / (256 * 16)), a.B);
  if (a.primitive == 1) {
    // point clouds: general pipeline for every view
    clear_keys_kernel<S><<<cgrid, 256, 0, stream>>>(keys, a.res, nullptr, boxes, view_hard);
    FP_CUDA(cudaGetLastError());
    point_kernel<S><<<dim3((a.V + 255) / 256, a.B), 256, 0, stream>>>(sv, keys, a.V, a.res, znear, zfar);
    FP_CUDA(cudaGetLastError());
    resolve_kernel<S, 2><<<rgrid, 256, 0, stream>>>(keys, sv, a.faces, sf, a.rgb, a.depth, a.V, a.res, nullptr, boxes, view_hard);
    FP_CUDA(cudaGetLastError());
    return 0;
  }
  const dim3 fgrid((a.F + 255) / 256, a.B);
  // FP_RASTER_TILE=1 selects the tile pipeline for the views it can take.  It is NOT the default: measured at the benchmark
  // shape (521 views x 20 480 faces, 224^2, ~1.3 px per triangle) it moves 4 x less DRAM but takes 3.7 ms against 2.2 ms --
  // both pipelines are bound by the per-triangle instruction stream (set-up in 64-bit integers, IEEE divisions in the depth
  // interpolation), which the tile pipeline runs three times (two binning passes + the tile kernel) instead of once, and
  // whose latency its 2-4 resident CTAs per SM hide less well than the 24 warps of triangle_kernel (profiles/r02e_*).
  static const int use_tile = [] { const char* e = getenv("FP_RASTER_TILE"); return e ? atoi(e) : 0; }();
  static const int big_pixels = [] { const char* e = getenv("FP_RASTER_BIG"); return e ? atoi(e) : BIG_TRI_PIXELS; }();   // perf experiments
  if (!use_tile) {
    clear_keys_kernel<S><<<cgrid, 256, 0, stream>>>(keys, a.res, nullptr, boxes, view_hard);
    FP_CUDA(cudaGetLastError());
    triangle_kernel<S><<<fgrid, 256, 0, stream>>>(sv, a.faces, keys, a.V, a.F, a.res, a.cull_backfaces, znear, zfar, nullptr, big_pixels);
    FP_CUDA(cudaGetLastError());
    hard_triangle_kernel<S><<<dim3(min(fgrid.x, 16u), a.B), 256, 0, stream>>>(sv, a.faces, cam, view_hard, keys, a.V, a.F, a.res, a.cull_backfaces);
    FP_CUDA(cudaGetLastError());
    const dim3 hgrid(min((a.res * a.res + 255) / 256, 24), a.B);
    if (a.texture != nullptr) {
      resolve_kernel<S, 1><<<rgrid, 256, 0, stream>>>(keys, sv, a.faces, sf, a.rgb, a.depth, a.V, a.res, nullptr, boxes, view_hard);
      resolve_hard_kernel<S, 1><<<hgrid, 256, 0, stream>>>(keys, sv, a.faces, sf, cam, view_hard, a.rgb, a.V, a.res);
    } else {
      resolve_kernel<S, 0><<<rgrid, 256, 0, stream>>>(keys, sv, a.faces, sf, a.rgb, a.depth, a.V, a.res, nullptr, boxes, view_hard);
      resolve_hard_kernel<S, 0><<<hgrid, 256, 0, stream>>>(keys, sv, a.faces, sf, cam, view_hard, a.rgb, a.V, a.res);
    }
    FP_CUDA(cudaGetLastError());
    return 0;
  }
  // ---- tile pipeline (views with route == 0)
  int* tile_count = reinterpret_cast<int*>(ws + w.counters);
  const size_t ntiles = size_t(a.B) * w.ntx * w.nty;
  int* tile_cursor = tile_count + ntiles;
  int* big_count = tile_cursor + ntiles;
  int* tile_offset = reinterpret_cast<int*>(ws + w.offsets);
  int* bin_list = reinterpret_cast<int*>(ws + w.bins);
  int4* big_list = reinterpret_cast<int4*>(ws + w.big);
  FP_REQUIRE(ntiles < (size_t(1) << 30), "raster: too many tiles");
  FP_CUDA(cudaMemsetAsync(tile_count, 0, (2 * ntiles + size_t(a.B)) * sizeof(int), stream));
  FP_CUDA(cudaMemsetAsync(a.rgb, 0, size_t(a.B) * a.res * a.res * 3, stream));
  FP_CUDA(cudaMemsetAsync(a.depth, 0, size_t(a.B) * a.res * a.res * sizeof(float), stream));
  bin_kernel<0><<<fgrid, 256, 0, stream>>>(sv, a.faces, view_hard, a.V, a.F, a.res, a.cull_backfaces, w.ntx, w.nty, tile_count,
                                           nullptr, nullptr, nullptr, nullptr);
  FP_CUDA(cudaGetLastError());
  tile_scan_kernel<<<1, 1024, 0, stream>>>(tile_count, tile_offset, int(ntiles));
  FP_CUDA(cudaGetLastError());
  bin_kernel<1><<<fgrid, 256, 0, stream>>>(sv, a.faces, view_hard, a.V, a.F, a.res, a.cull_backfaces, w.ntx, w.nty, tile_cursor,
                                           tile_offset, bin_list, big_count, big_list);
  FP_CUDA(cudaGetLastError());
  const dim3 tgrid(w.ntx * w.nty, a.B);
  static const int occ = [] { const char* e = getenv("FP_TILE_OCC"); return e ? atoi(e) : 4; }();   // perf experiments
#define FP_TILE_LAUNCH(MODE_, OCC_)                                                                                              \
  tile_kernel<S, MODE_, OCC_><<<tgrid, 256, 0, stream>>>(sv, a.faces, sf, view_hard, tile_offset, bin_list, big_count, big_list, \
                                                         a.rgb, a.depth, a.V, a.F, a.res, w.ntx, w.nty, a.cull_backfaces, znear, zfar)
  if (a.texture != nullptr) {
    if (occ == 2) FP_TILE_LAUNCH(1, 2); else if (occ == 3) FP_TILE_LAUNCH(1, 3); else FP_TILE_LAUNCH(1, 4);
  } else {
    if (occ == 2) FP_TILE_LAUNCH(0, 2); else if (occ == 3) FP_TILE_LAUNCH(0, 3); else FP_TILE_LAUNCH(0, 4);
  }
#undef FP_TILE_LAUNCH
  FP_CUDA(cudaGetLastError());
  // ---- general pipeline (views with route == 1: every kernel returns at once for the others; small grids, the kernels
  //      loop, so that those returns cost next to nothing)
  const dim3 rgrid_routed(min((a.res * a.res + 255) / 256, 24), a.B);
  clear_keys_kernel<S><<<dim3(min(cgrid.x, 8u), a.B), 256, 0, stream>>>(keys, a.res, view_hard, boxes, view_hard);
  FP_CUDA(cudaGetLastError());
  triangle_kernel<S><<<fgrid, 256, 0, stream>>>(sv, a.faces, keys, a.V, a.F, a.res, a.cull_backfaces, znear, zfar, view_hard, big_pixels);
  FP_CUDA(cudaGetLastError());
  hard_triangle_kernel<S><<<dim3(min(fgrid.x, 16u), a.B), 256, 0, stream>>>(sv, a.faces, cam, view_hard, keys, a.V, a.F, a.res, a.cull_backfaces);
  FP_CUDA(cudaGetLastError());
  if (a.texture != nullptr) {
    resolve_kernel<S, 1><<<rgrid_routed, 256, 0, stream>>>(keys, sv, a.faces, sf, a.rgb, a.depth, a.V, a.res, view_hard, boxes, view_hard);
    resolve_hard_kernel<S, 1><<<rgrid_routed, 256, 0, stream>>>(keys, sv, a.faces, sf, cam, view_hard, a.rgb, a.V, a.res);
  } else {
    resolve_kernel<S, 0><<<rgrid_routed, 256, 0, stream>>>(keys, sv, a.faces, sf, a.rgb, a.depth, a.V, a.res, view_hard, boxes, view_hard);
    resolve_hard_kernel<S, 0><<<rgrid_routed, 256, 0, stream>>>(keys, sv, a.faces, sf, cam, view_hard, a.rgb, a.V, a.res);
  }
  FP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int raster_workspace_bytes(int B, int V, int F, int res, int msaa, size_t* bytes) {
  FP_REQUIRE(msaa == 1 || msaa == 4, "raster: msaa must be 1 or 4");
  FP_REQUIRE(B >= 0 && V >= 0 && F >= 0 && res > 0, "raster: bad sizes");
  *bytes = raster_layout(B, V, F, res, msaa).total;
  return 0;
}

int rasterize(const RasterArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  FP_REQUIRE(a.res > 0 && a.res % 4 == 0, "raster: resolution %d must be a positive multiple of 4", a.res);
  FP_REQUIRE(a.res <= 16384, "raster: resolution %d too large", a.res);
  FP_REQUIRE(a.msaa == 1 || a.msaa == 4, "raster: msaa must be 1 or 4");
  FP_REQUIRE(a.primitive == 0 || a.primitive == 1, "raster: primitive must be 0 (triangles) or 1 (points)");
  FP_REQUIRE(a.V > 0 && (a.primitive == 1 || a.F > 0), "raster: empty mesh (V=%d, F=%d)", a.V, a.F);
  FP_REQUIRE(a.B <= 65535, "raster: at most 65535 views per call");
  if (a.texture != nullptr) {
    FP_REQUIRE(a.primitive == 0 && a.uv != nullptr && a.srgb_lut != nullptr, "raster: a texture needs triangles, uv and srgb_lut");
    FP_REQUIRE(a.tex_w > 0 && a.tex_h > 0 && a.tex_levels > 0 && a.tex_levels <= 16, "raster: bad texture size %dx%d, %d levels",
               a.tex_w, a.tex_h, a.tex_levels);
  } else {
    FP_REQUIRE(a.colors != nullptr, "raster: neither vertex colours nor a texture");
  }
  if (a.B <= 0) return 0;
  const RasterWorkspace w = raster_layout(a.B, a.V, a.primitive == 1 ? 0 : a.F, a.res, a.msaa);
  FP_REQUIRE(workspace_bytes >= w.total, "raster: workspace too small (%zu < %zu)", workspace_bytes, w.total);
  FP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "raster: workspace must be 256-byte aligned");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  ScreenVertex* sv = reinterpret_cast<ScreenVertex*>(ws + w.sv);
  int* view_hard = reinterpret_cast<int*>(ws + w.flags);
  // algorithmic bytes: RGB u8 + depth f32 out.  Launches: vertex + (points: clear, point, resolve | triangles: 2 bin passes,
  // scan, tile + the general pipeline's clear, triangle, hard triangle, 2 resolves, which return at once for tile views)
  ProfScope prof(PROF_RASTER, double(a.B) * a.res * a.res * 7.0, a.primitive == 1 ? 4 : 10, stream);
  int* boxes = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(view_hard + a.B) + 15) & ~uintptr_t(15));   // int4 per view
  FP_CUDA(cudaMemsetAsync(view_hard, 0, size_t(a.B) * 5 * sizeof(int) + 16, stream));
  const float znear = a.znear > 0.f ? a.znear : ZNEAR_DEFAULT, zfar = a.zfar > 0.f ? a.zfar : ZFAR_DEFAULT;
  vertex_kernel<<<dim3((a.V + 255) / 256, a.B), 256, 0, stream>>>(a.verts, a.poses, sv, a.V, a.B, a.fx, a.fy, a.cx, a.cy,
                                                                  a.view_k, znear, zfar, view_hard, boxes);
  FP_CUDA(cudaGetLastError());
  return a.msaa == 4 ? launch_raster<4>(a, ws, w, znear, zfar, stream) : launch_raster<1>(a, ws, w, znear, zfar, stream);
}

}  // namespace fp
