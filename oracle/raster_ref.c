/* ORACLE (test infrastructure, never shipped): plain-C restatement of the rasterisation semantics that the
 * reference obtains from pyrender/OpenGL (reference src/pipeline/retrieval/renderer.py:37-66): pinhole camera
 * in the OpenCV frame, two-sided triangles, ambient-only (2,2,2) lighting with 1/2.2 gamma, transparent black
 * background, 4x multisampling with one shading sample per (triangle, pixel), linear depth of sample 0.
 * Surfaces: vertex colours, a base-colour texture (REPEAT wrap, trilinear over the caller's box-filtered mip chain,
 * x^2.2 after filtering, optional vertex-colour multiplier) or 1-pixel point sprites (trimesh.PointCloud inputs).
 *
 * pyrender / PyOpenGL / EGL are not installed and OpenGL rasterisation is not bit-specified, so this oracle
 * cannot be pinned against the real renderer: PARITY UNPINNED for row R.  What it pins is that the
 * CUDA rasteriser implements exactly this written-down specification: sequential loops, a classic z-buffer
 * (no atomics, no binning), compiled with -ffp-contract=off so every fp32 operation rounds once.
 *
 * Near / far planes and the coordinate guard band.  A triangle whose three vertices all project inside the guard band
 * with znear < Z < zfar takes the fixed-point path above.  Any other triangle ("hard": it straddles the near plane, has
 * a vertex behind the camera, beyond zfar, or further than 16384 px off screen) is NOT dropped: GL clips it, and so
 * does this specification, per sample, by rasterising it in homogeneous form (Olano & Greer 1997): with camera-space
 * vertices P0,P1,P2 and the sample ray d = ((sx - cx)/fx, (sy - cy)/fy, 1), b_i = sign(det) * d . (P_{i+1} x P_{i+2}),
 * the sample is covered iff all b_i >= 0 and their sum > 0, its depth is |det| / sum(b_i), and the existing per-sample
 * test znear < z < zfar cuts the part in front of the near plane (and behind the far plane) away -- exactly what
 * clipping the triangle against those planes would leave.  Attributes interpolate with the weights b_i / sum.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC oracle/raster_ref.c -o oracle/_build/libraster_ref.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SUB 8
#define ONE 256
static float ZNEAR = 0.05f, ZFAR = 100.0f;   /* pyrender IntrinsicsCamera defaults; raster_ref3 may override per call */
#define COORD_LIMIT (1 << 22)

typedef struct { int x, y, ok; float z, iz; float X, Y, Z; } SV;

static const int OFF1[1][2] = {{128, 128}};
static const int OFF4[4][2] = {{96, 32}, {224, 96}, {32, 160}, {160, 224}};

static void project(const float* verts, const float* P, int V, float fx, float fy, float cx, float cy, SV* sv) {
  for (int i = 0; i < V; ++i) {
    float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    float X = ((P[0] * x + P[1] * y) + P[2] * z) + P[3];
    float Y = ((P[4] * x + P[5] * y) + P[6] * z) + P[7];
    float Z = ((P[8] * x + P[9] * y) + P[10] * z) + P[11];
    sv[i].ok = 0;
    sv[i].X = X; sv[i].Y = Y; sv[i].Z = Z;
    if (!(Z > ZNEAR) || !(Z < ZFAR)) continue;
    float u = (fx * X) / Z + cx;
    float v = (fy * Y) / Z + cy;
    float uf = floorf(u * (float)ONE + 0.5f), vf = floorf(v * (float)ONE + 0.5f);
    if (!(fabsf(uf) < (float)COORD_LIMIT) || !(fabsf(vf) < (float)COORD_LIMIT)) continue;
    sv[i].x = (int)uf; sv[i].y = (int)vf; sv[i].z = Z; sv[i].iz = 1.0f / Z; sv[i].ok = 1;
  }
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

typedef struct {
  const uint8_t* colors;    /* [V,3] or NULL */
  const float* uv;          /* [V,2] or NULL */
  const uint8_t* texture;   /* RGBA8 mip chain or NULL */
  const float* srgb_lut;    /* [65536] */
  const uint8_t* gamma_lut; /* [65536] */
  int tex_w, tex_h, tex_levels;
  int points;
  float ambient, ambient_255;
} Surface;

static int to_unorm8(float lin, const uint8_t* lut) {
  lin = fminf(fmaxf(lin, 0.f), 1.f);
  if (!(lin == lin)) lin = 0.f;
  return lut[(int)(lin * 65535.0f + 0.5f)];
}

static int wrap_repeat(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

static void bilinear(const Surface* sf, int lvl, float u, float v, float out[3]) {
  size_t off = 0;
  int W = sf->tex_w, H = sf->tex_h;
  for (int l = 0; l < lvl; ++l) { off += (size_t)W * H * 4; W = imax(1, W >> 1); H = imax(1, H >> 1); }
  float x = u * (float)W + -0.5f;
  float y = (1.0f + -v) * (float)H + -0.5f;
  float xf = floorf(x), yf = floorf(y);
  float fx = x + -xf, fy = y + -yf;
  int ix = (int)fminf(fmaxf(xf, -1.0e9f), 1.0e9f), iy = (int)fminf(fmaxf(yf, -1.0e9f), 1.0e9f);
  int x0 = wrap_repeat(ix, W), x1 = wrap_repeat(ix + 1, W), y0 = wrap_repeat(iy, H), y1 = wrap_repeat(iy + 1, H);
  const uint8_t* t = sf->texture + off;
  for (int ch = 0; ch < 3; ++ch) {
    float a00 = (float)t[((size_t)y0 * W + x0) * 4 + ch], a10 = (float)t[((size_t)y0 * W + x1) * 4 + ch];
    float a01 = (float)t[((size_t)y1 * W + x0) * 4 + ch], a11 = (float)t[((size_t)y1 * W + x1) * 4 + ch];
    float top = a00 + fx * (a10 + -a00);
    float bot = a01 + fx * (a11 + -a01);
    out[ch] = top + fy * (bot + -top);
  }
}

static float lod_from_rho2(float rho2, int levels) {
  float top = (float)(levels - 1);
  if (!(rho2 < 1.0e30f)) return top;
  if (!(rho2 > 1.0f)) return 0.f;
  uint32_t bits;
  memcpy(&bits, &rho2, 4);
  float e = (float)((int)(bits >> 23) - 127);
  float m = (float)(bits & 0x7fffffu) * 1.1920928955078125e-07f;
  return fminf(0.5f * (e + m), top);
}

typedef struct { float w0, w1, w2, wsum; } Weights;

static Weights persp_weights(SV v0, SV v1, SV v2, float fa, int64_t sx, int64_t sy) {
  int64_t e0 = (int64_t)(v2.x - v1.x) * (sy - v1.y) - (int64_t)(v2.y - v1.y) * (sx - v1.x);
  int64_t e1 = (int64_t)(v0.x - v2.x) * (sy - v2.y) - (int64_t)(v0.y - v2.y) * (sx - v2.x);
  int64_t e2 = (int64_t)(v1.x - v0.x) * (sy - v0.y) - (int64_t)(v1.y - v0.y) * (sx - v0.x);
  Weights w;
  /* w_i = e_i / z_i (1 / area cancels in the ratio); wsum holds 1 / sum w_i */
  (void)fa;
  w.w0 = (float)e0 * v0.iz; w.w1 = (float)e1 * v1.iz; w.w2 = (float)e2 * v2.iz;
  w.wsum = 1.0f / ((w.w0 + w.w1) + w.w2);
  return w;
}
static float interp(Weights w, float a0, float a1, float a2) { return ((w.w0 * a0 + w.w1 * a1) + w.w2 * a2) * w.wsum; }

/* ---- hard triangles: homogeneous rasterisation ------------------------------------------------------------------ */
typedef struct { float n0[3], n1[3], n2[3]; float adet, sgn; int xlo, xhi, ylo, yhi; int valid; } Hard;

static void cross3(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

static Hard hard_setup(SV v0, SV v1, SV v2, float fx, float fy, float cx, float cy, int res, int cull) {
  Hard h;
  h.valid = 0;
  /* entirely in front of the near plane or behind the far plane: nothing survives the per-sample depth test */
  if (!(v0.Z > ZNEAR) && !(v1.Z > ZNEAR) && !(v2.Z > ZNEAR)) return h;
  if (!(v0.Z < ZFAR) && !(v1.Z < ZFAR) && !(v2.Z < ZFAR)) return h;
  float p0[3] = {v0.X, v0.Y, v0.Z}, p1[3] = {v1.X, v1.Y, v1.Z}, p2[3] = {v2.X, v2.Y, v2.Z};
  cross3(p1, p2, h.n0); cross3(p2, p0, h.n1); cross3(p0, p1, h.n2);
  float det = (p0[0] * h.n0[0] + p0[1] * h.n0[1]) + p0[2] * h.n0[2];
  if (!(det != 0.f) || !(det == det)) return h;
  if (cull && det > 0.f) return h;           /* sign(det) = sign of the projected area: same rule as the fixed-point path */
  h.sgn = det > 0.f ? 1.0f : -1.0f;
  h.adet = fabsf(det);
  /* conservative pixel bounds: projections of the vertices in front of the near plane and of the edge / near-plane
   * intersections, +-1 px; anything not finite -> the whole viewport */
  const SV* v[3] = {&v0, &v1, &v2};
  float umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY;
  int finite = 1;
  for (int i = 0; i < 3; ++i) {
    const SV* a = v[i];
    const SV* b = v[(i + 1) % 3];
    if (a->Z > ZNEAR) {
      float u = (fx * a->X) / a->Z + cx, w = (fy * a->Y) / a->Z + cy;
      umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, w); vmax = fmaxf(vmax, w);
      if (!(fabsf(u) < 1.0e9f) || !(fabsf(w) < 1.0e9f)) finite = 0;
    }
    if ((a->Z > ZNEAR) != (b->Z > ZNEAR)) {
      float t = (ZNEAR - a->Z) / (b->Z - a->Z);
      float X = a->X + t * (b->X - a->X), Y = a->Y + t * (b->Y - a->Y);
      float u = (fx * X) / ZNEAR + cx, w = (fy * Y) / ZNEAR + cy;
      umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, w); vmax = fmaxf(vmax, w);
      if (!(fabsf(u) < 1.0e9f) || !(fabsf(w) < 1.0e9f)) finite = 0;
    }
  }
  if (finite) {
    h.xlo = (int)fmaxf(floorf(umin) - 1.0f, 0.f); h.xhi = (int)fminf(floorf(umax) + 1.0f, (float)(res - 1));
    h.ylo = (int)fmaxf(floorf(vmin) - 1.0f, 0.f); h.yhi = (int)fminf(floorf(vmax) + 1.0f, (float)(res - 1));
  } else {
    h.xlo = 0; h.xhi = res - 1; h.ylo = 0; h.yhi = res - 1;
  }
  h.valid = h.xlo <= h.xhi && h.ylo <= h.yhi;
  return h;
}

/* weights of the fixed-point sample position (sx, sy): returns 1 if covered (b_i >= 0, sum > 0) */
static int hard_weights(const Hard* h, float fx, float fy, float cx, float cy, int64_t sx, int64_t sy, float b[3],
                        float* sum) {
  float dx = ((float)sx * 0.00390625f - cx) / fx, dy = ((float)sy * 0.00390625f - cy) / fy;
  b[0] = h->sgn * ((dx * h->n0[0] + dy * h->n0[1]) + h->n0[2]);
  b[1] = h->sgn * ((dx * h->n1[0] + dy * h->n1[1]) + h->n1[2]);
  b[2] = h->sgn * ((dx * h->n2[0] + dy * h->n2[1]) + h->n2[2]);
  *sum = (b[0] + b[1]) + b[2];
  return b[0] >= 0.f && b[1] >= 0.f && b[2] >= 0.f && *sum > 0.f;
}

static float hard_interp(const float b[3], float sum, float a0, float a1, float a2) {
  return ((b[0] * a0 + b[1] * a1) + b[2] * a2) / sum;
}

typedef struct { float fx, fy, cx, cy; int cull, res; } Cam;
static Cam CAM;

/* colour of triangle f at the centre of pixel (px, py) -> out[3] unorm8 */
static void shade(const Surface* sf, const int32_t* faces, const SV* sv, int f, int px, int py, int out[3]) {
  int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
  SV v0 = sv[i0], v1 = sv[i1], v2 = sv[i2];
  int64_t sx = ((int64_t)px << SUB) + 128, sy = ((int64_t)py << SUB) + 128;
  const int hard = !v0.ok || !v1.ok || !v2.ok;
  float vc[3] = {255.f, 255.f, 255.f};
  float ua = 0.f, ub = 0.f, uc = 0.f, va = 0.f, vb = 0.f, vcc = 0.f, u = 0.f, v = 0.f;
  float dux = 0.f, dvx = 0.f, duy = 0.f, dvy = 0.f;
  float fw = (float)sf->tex_w, fh = (float)sf->tex_h;
  if (hard) {
    Hard h = hard_setup(v0, v1, v2, CAM.fx, CAM.fy, CAM.cx, CAM.cy, CAM.res, 0);
    float b[3], sum;
    hard_weights(&h, CAM.fx, CAM.fy, CAM.cx, CAM.cy, sx, sy, b, &sum);
    if (sf->texture == NULL || sf->colors != NULL)
      for (int ch = 0; ch < 3; ++ch)
        vc[ch] = hard_interp(b, sum, (float)sf->colors[3 * i0 + ch], (float)sf->colors[3 * i1 + ch], (float)sf->colors[3 * i2 + ch]);
    if (sf->texture != NULL) {
      ua = sf->uv[2 * i0]; ub = sf->uv[2 * i1]; uc = sf->uv[2 * i2];
      va = sf->uv[2 * i0 + 1]; vb = sf->uv[2 * i1 + 1]; vcc = sf->uv[2 * i2 + 1];
      u = hard_interp(b, sum, ua, ub, uc); v = hard_interp(b, sum, va, vb, vcc);
      float bx[3], sumx, by[3], sumy;
      hard_weights(&h, CAM.fx, CAM.fy, CAM.cx, CAM.cy, sx + ONE, sy, bx, &sumx);
      hard_weights(&h, CAM.fx, CAM.fy, CAM.cx, CAM.cy, sx, sy + ONE, by, &sumy);
      dux = (hard_interp(bx, sumx, ua, ub, uc) + -u) * fw; dvx = (hard_interp(bx, sumx, va, vb, vcc) + -v) * fh;
      duy = (hard_interp(by, sumy, ua, ub, uc) + -u) * fw; dvy = (hard_interp(by, sumy, va, vb, vcc) + -v) * fh;
    }
  } else {
    int64_t area = (int64_t)(v1.x - v0.x) * (v2.y - v0.y) - (int64_t)(v2.x - v0.x) * (v1.y - v0.y);
    if (area < 0) { SV t = v1; v1 = v2; v2 = t; int ti = i1; i1 = i2; i2 = ti; area = -area; }
    float fa = (float)area;
    Weights w = persp_weights(v0, v1, v2, fa, sx, sy);
    if (sf->texture == NULL || sf->colors != NULL)
      for (int ch = 0; ch < 3; ++ch)
        vc[ch] = interp(w, (float)sf->colors[3 * i0 + ch], (float)sf->colors[3 * i1 + ch], (float)sf->colors[3 * i2 + ch]);
    if (sf->texture != NULL) {
      ua = sf->uv[2 * i0]; ub = sf->uv[2 * i1]; uc = sf->uv[2 * i2];
      va = sf->uv[2 * i0 + 1]; vb = sf->uv[2 * i1 + 1]; vcc = sf->uv[2 * i2 + 1];
      u = interp(w, ua, ub, uc); v = interp(w, va, vb, vcc);
      Weights wx = persp_weights(v0, v1, v2, fa, sx + ONE, sy), wy = persp_weights(v0, v1, v2, fa, sx, sy + ONE);
      dux = (interp(wx, ua, ub, uc) + -u) * fw; dvx = (interp(wx, va, vb, vcc) + -v) * fh;
      duy = (interp(wy, ua, ub, uc) + -u) * fw; dvy = (interp(wy, va, vb, vcc) + -v) * fh;
    }
  }
  if (sf->texture == NULL) {
    for (int ch = 0; ch < 3; ++ch) out[ch] = to_unorm8(vc[ch] * sf->ambient_255, sf->gamma_lut);
    return;
  }
  float rx = dux * dux + dvx * dvx, ry = duy * duy + dvy * dvy;
  float lod = lod_from_rho2(fmaxf(rx, ry), sf->tex_levels);
  int l0 = (int)lod;
  float t = lod + -(float)l0;
  float ca[3];
  bilinear(sf, l0, u, v, ca);
  if (t > 0.f) {
    float cb[3];
    bilinear(sf, imin(l0 + 1, sf->tex_levels - 1), u, v, cb);
    for (int ch = 0; ch < 3; ++ch) ca[ch] = ca[ch] + t * (cb[ch] + -ca[ch]);
  }
  for (int ch = 0; ch < 3; ++ch) {
    float cn = ca[ch] / 255.0f;
    cn = fminf(fmaxf(cn, 0.f), 1.f);
    if (!(cn == cn)) cn = 0.f;
    float lin = sf->srgb_lut[(int)(cn * 65535.0f + 0.5f)];
    if (sf->colors != NULL) lin = lin * (vc[ch] / 255.0f);
    out[ch] = to_unorm8(lin * sf->ambient, sf->gamma_lut);
  }
}

/* One view.  rgb: res*res*3 u8, depth: res*res f32. */
static void render_view(const float* verts, const int32_t* faces, const Surface* sf, int V, int F,
                        const float* P, float fx, float fy, float cx, float cy, int res, int msaa, int cull,
                        uint8_t* rgb, float* depth, SV* sv, float* zbuf, int32_t* fbuf) {
  const int S = msaa;
  const int (*off)[2] = (S == 4) ? OFF4 : OFF1;
  project(verts, P, V, fx, fy, cx, cy, sv);
  CAM.fx = fx; CAM.fy = fy; CAM.cx = cx; CAM.cy = cy; CAM.cull = cull; CAM.res = res;
  const size_t ns = (size_t)res * res * S;
  for (size_t i = 0; i < ns; ++i) { zbuf[i] = INFINITY; fbuf[i] = -1; }
  /* GL_POINTS, size 1: square sprite [x-.5, x+.5) x [y-.5, y+.5), flat depth; lower vertex index wins depth ties */
  for (int i = 0; sf->points && i < V; ++i) {
    if (!sv[i].ok || !(sv[i].z > ZNEAR && sv[i].z < ZFAR)) continue;
    int bx = sv[i].x - ONE / 2, by = sv[i].y - ONE / 2;
    for (int py = by >> SUB; py <= (by + ONE - 1) >> SUB; ++py)
      for (int px = bx >> SUB; px <= (bx + ONE - 1) >> SUB; ++px) {
        if (px < 0 || px >= res || py < 0 || py >= res) continue;
        for (int s = 0; s < S; ++s) {
          int sx = (px << SUB) + off[s][0], sy = (py << SUB) + off[s][1];
          if (sx < bx || sx >= bx + ONE || sy < by || sy >= by + ONE) continue;
          size_t k = ((size_t)py * res + px) * S + s;
          if (sv[i].z < zbuf[k]) { zbuf[k] = sv[i].z; fbuf[k] = i; }
        }
      }
  }
  for (int f = 0; !sf->points && f < F; ++f) {
    SV v0 = sv[faces[3 * f]], v1 = sv[faces[3 * f + 1]], v2 = sv[faces[3 * f + 2]];
    if (!v0.ok || !v1.ok || !v2.ok) {
      /* hard triangle: homogeneous rasterisation, clipped per sample by the depth test */
      Hard h = hard_setup(v0, v1, v2, fx, fy, cx, cy, res, cull);
      if (!h.valid) continue;
      for (int py = h.ylo; py <= h.yhi; ++py)
        for (int px = h.xlo; px <= h.xhi; ++px)
          for (int s = 0; s < S; ++s) {
            int64_t sx = ((int64_t)px << SUB) + off[s][0], sy = ((int64_t)py << SUB) + off[s][1];
            float b[3], sum;
            if (!hard_weights(&h, fx, fy, cx, cy, sx, sy, b, &sum)) continue;
            float z = h.adet / sum;
            if (!(z > ZNEAR && z < ZFAR)) continue;
            size_t k = ((size_t)py * res + px) * S + s;
            if (z < zbuf[k]) { zbuf[k] = z; fbuf[k] = f; }
          }
      continue;
    }
    int64_t area = (int64_t)(v1.x - v0.x) * (v2.y - v0.y) - (int64_t)(v2.x - v0.x) * (v1.y - v0.y);
    if (area == 0) continue;
    if (cull && area > 0) continue;          /* y-down image: GL front faces have negative area here */
    if (area < 0) { SV t = v1; v1 = v2; v2 = t; area = -area; }
    int minx = imin(v0.x, imin(v1.x, v2.x)), maxx = imax(v0.x, imax(v1.x, v2.x));
    int miny = imin(v0.y, imin(v1.y, v2.y)), maxy = imax(v0.y, imax(v1.y, v2.y));
    int x0 = imax(0, minx >> SUB), x1 = imin(res - 1, maxx >> SUB);
    int y0 = imax(0, miny >> SUB), y1 = imin(res - 1, maxy >> SUB);
    if (x0 > x1 || y0 > y1) continue;
    const SV* ea[3] = {&v1, &v2, &v0};
    const SV* eb[3] = {&v2, &v0, &v1};
    const float iz[3] = {v0.iz, v1.iz, v2.iz};
    const float farea = (float)area;
    for (int py = y0; py <= y1; ++py)
      for (int px = x0; px <= x1; ++px)
        for (int s = 0; s < S; ++s) {
          int64_t sx = ((int64_t)px << SUB) + off[s][0], sy = ((int64_t)py << SUB) + off[s][1];
          int64_t e[3];
          int inside = 1;
          for (int i = 0; i < 3; ++i) {
            int64_t dx = eb[i]->x - ea[i]->x, dy = eb[i]->y - ea[i]->y;
            e[i] = dx * (sy - ea[i]->y) - dy * (sx - ea[i]->x);
            int topleft = (dy < 0) || (dy == 0 && dx > 0);
            if (e[i] < 0 || (e[i] == 0 && !topleft)) inside = 0;
          }
          if (!inside) continue;
          /* perspective-correct depth: 1/z is affine in screen space, z = area / (e0/z0 + e1/z1 + e2/z2) */
          float den = ((float)e[0] * iz[0] + (float)e[1] * iz[1]) + (float)e[2] * iz[2];
          float z = farea / den;
          if (!(z > ZNEAR && z < ZFAR)) continue;
          size_t k = ((size_t)py * res + px) * S + s;
          /* GL_LESS with primitives in order: strictly nearer wins; on equal depth the earlier face stays */
          if (z < zbuf[k]) { zbuf[k] = z; fbuf[k] = f; }
        }
  }
  for (int py = 0; py < res; ++py)
    for (int px = 0; px < res; ++px) {
      int acc[3] = {0, 0, 0};
      size_t k0 = ((size_t)py * res + px) * S;
      for (int s = 0; s < S; ++s) {
        int f = fbuf[k0 + s];
        if (f < 0) continue;
        int col[3];
        if (sf->points)
          for (int ch = 0; ch < 3; ++ch) col[ch] = to_unorm8((float)sf->colors[3 * f + ch] * sf->ambient_255, sf->gamma_lut);
        else
          shade(sf, faces, sv, f, px, py, col);
        for (int ch = 0; ch < 3; ++ch) acc[ch] += col[ch];
      }
      uint8_t* o = rgb + ((size_t)py * res + px) * 3;
      for (int ch = 0; ch < 3; ++ch) o[ch] = (uint8_t)(S == 4 ? (acc[ch] + 2) >> 2 : acc[ch]);
      depth[(size_t)py * res + px] = fbuf[k0] >= 0 ? zbuf[k0] : 0.f;
    }
}

int raster_ref3(const float* verts, const int32_t* faces, const uint8_t* colors, int V, int F, const float* poses,
                int B, float fx, float fy, float cx, float cy, int res, int msaa, int cull, const uint8_t* lut,
                uint8_t* rgb, float* depth, int points, const float* uv, const uint8_t* texture, int tex_w, int tex_h,
                int tex_levels, const float* srgb_lut, float ambient, float znear, float zfar, const float* view_k) {
  if (msaa != 1 && msaa != 4) return -1;
  if (texture == NULL && colors == NULL) return -3;
  Surface sf = {colors, uv, texture, srgb_lut, lut, tex_w, tex_h, tex_levels, points, 0.f, 0.f};
  sf.ambient = ambient > 0.f ? ambient : 2.0f;
  sf.ambient_255 = sf.ambient / 255.0f;
  ZNEAR = znear > 0.f ? znear : 0.05f;
  ZFAR = zfar > 0.f ? zfar : 100.0f;
  SV* sv = (SV*)malloc(sizeof(SV) * (size_t)V);
  float* zbuf = (float*)malloc(sizeof(float) * (size_t)res * res * msaa);
  int32_t* fbuf = (int32_t*)malloc(sizeof(int32_t) * (size_t)res * res * msaa);
  if (!sv || !zbuf || !fbuf) return -2;
  for (int b = 0; b < B; ++b) {
    if (view_k != NULL) { fx = view_k[4 * b]; fy = view_k[4 * b + 1]; cx = view_k[4 * b + 2]; cy = view_k[4 * b + 3]; }
    render_view(verts, faces, &sf, V, F, poses + (size_t)b * 12, fx, fy, cx, cy, res, msaa, cull,
                rgb + (size_t)b * res * res * 3, depth + (size_t)b * res * res, sv, zbuf, fbuf);
  }
  free(sv); free(zbuf); free(fbuf);
  ZNEAR = 0.05f; ZFAR = 100.0f;
  return 0;
}

int raster_ref2(const float* verts, const int32_t* faces, const uint8_t* colors, int V, int F, const float* poses,
                int B, float fx, float fy, float cx, float cy, int res, int msaa, int cull, const uint8_t* lut,
                uint8_t* rgb, float* depth, int points, const float* uv, const uint8_t* texture, int tex_w, int tex_h,
                int tex_levels, const float* srgb_lut) {
  return raster_ref3(verts, faces, colors, V, F, poses, B, fx, fy, cx, cy, res, msaa, cull, lut, rgb, depth, points, uv,
                     texture, tex_w, tex_h, tex_levels, srgb_lut, 0.f, 0.f, 0.f, NULL);
}

int raster_ref(const float* verts, const int32_t* faces, const uint8_t* colors, int V, int F, const float* poses,
               int B, float fx, float fy, float cx, float cy, int res, int msaa, int cull, const uint8_t* lut,
               uint8_t* rgb, float* depth) {
  return raster_ref2(verts, faces, colors, V, F, poses, B, fx, fy, cx, cy, res, msaa, cull, lut, rgb, depth, 0, NULL,
                     NULL, 0, 0, 0, NULL);
}
