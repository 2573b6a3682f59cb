/* ORACLE (test infrastructure, never shipped): plain-C restatement of the rasterisation semantics that the
 * reference obtains from pyrender/OpenGL (reference src/pipeline/retrieval/renderer.py:37-66): pinhole camera
 * in the OpenCV frame, two-sided triangles, ambient-only (2,2,2) lighting with 1/2.2 gamma, transparent black
 * background, 4x multisampling with one shading sample per (triangle, pixel), linear depth of sample 0.
 *
 * pyrender / PyOpenGL / EGL are not installed and OpenGL rasterisation is not bit-specified, so this oracle
 * cannot be pinned against the real renderer: PARITY UNPINNED for row R.  What it pins is that the
 * CUDA rasteriser implements exactly this written-down specification: sequential loops, a classic z-buffer
 * (no atomics, no binning), compiled with -ffp-contract=off so every fp32 operation rounds once.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC oracle/raster_ref.c -o oracle/_build/libraster_ref.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SUB 8
#define ONE 256
static const float ZNEAR = 0.05f, ZFAR = 100.0f;
#define COORD_LIMIT (1 << 22)

typedef struct { int x, y, ok; float z, iz; } SV;

static const int OFF1[1][2] = {{128, 128}};
static const int OFF4[4][2] = {{96, 32}, {224, 96}, {32, 160}, {160, 224}};

static void project(const float* verts, const float* P, int V, float fx, float fy, float cx, float cy, SV* sv) {
  for (int i = 0; i < V; ++i) {
    float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    float X = ((P[0] * x + P[1] * y) + P[2] * z) + P[3];
    float Y = ((P[4] * x + P[5] * y) + P[6] * z) + P[7];
    float Z = ((P[8] * x + P[9] * y) + P[10] * z) + P[11];
    sv[i].ok = 0;
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

/* One view.  rgb: res*res*3 u8, depth: res*res f32. */
static void render_view(const float* verts, const int32_t* faces, const uint8_t* colors, int V, int F,
                        const float* P, float fx, float fy, float cx, float cy, int res, int msaa, int cull,
                        const uint8_t* lut, uint8_t* rgb, float* depth, SV* sv, float* zbuf, int32_t* fbuf) {
  const int S = msaa;
  const int (*off)[2] = (S == 4) ? OFF4 : OFF1;
  project(verts, P, V, fx, fy, cx, cy, sv);
  const size_t ns = (size_t)res * res * S;
  for (size_t i = 0; i < ns; ++i) { zbuf[i] = INFINITY; fbuf[i] = -1; }
  for (int f = 0; f < F; ++f) {
    SV v0 = sv[faces[3 * f]], v1 = sv[faces[3 * f + 1]], v2 = sv[faces[3 * f + 2]];
    if (!v0.ok || !v1.ok || !v2.ok) continue;
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
          float w0 = (float)e[0] / farea, w1 = (float)e[1] / farea, w2 = (float)e[2] / farea;
          float izs = (w0 * iz[0] + w1 * iz[1]) + w2 * iz[2];
          float z = 1.0f / izs;
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
        int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
        SV v0 = sv[i0], v1 = sv[i1], v2 = sv[i2];
        int64_t area = (int64_t)(v1.x - v0.x) * (v2.y - v0.y) - (int64_t)(v2.x - v0.x) * (v1.y - v0.y);
        if (area < 0) { SV t = v1; v1 = v2; v2 = t; int ti = i1; i1 = i2; i2 = ti; area = -area; }
        int64_t sx = ((int64_t)px << SUB) + 128, sy = ((int64_t)py << SUB) + 128;
        int64_t e0 = (int64_t)(v2.x - v1.x) * (sy - v1.y) - (int64_t)(v2.y - v1.y) * (sx - v1.x);
        int64_t e1 = (int64_t)(v0.x - v2.x) * (sy - v2.y) - (int64_t)(v0.y - v2.y) * (sx - v2.x);
        int64_t e2 = (int64_t)(v1.x - v0.x) * (sy - v0.y) - (int64_t)(v1.y - v0.y) * (sx - v0.x);
        float fa = (float)area;
        float w0 = ((float)e0 / fa) * v0.iz, w1 = ((float)e1 / fa) * v1.iz, w2 = ((float)e2 / fa) * v2.iz;
        float wsum = (w0 + w1) + w2;
        for (int ch = 0; ch < 3; ++ch) {
          float a0 = (float)colors[3 * i0 + ch], a1 = (float)colors[3 * i1 + ch], a2 = (float)colors[3 * i2 + ch];
          float c = ((w0 * a0 + w1 * a1) + w2 * a2) / wsum;
          float lin = c * (2.0f / 255.0f);
          lin = fminf(fmaxf(lin, 0.f), 1.f);
          if (!(lin == lin)) lin = 0.f;
          int idx = (int)(lin * 65535.0f + 0.5f);
          acc[ch] += lut[idx];
        }
      }
      uint8_t* o = rgb + ((size_t)py * res + px) * 3;
      for (int ch = 0; ch < 3; ++ch) o[ch] = (uint8_t)(S == 4 ? (acc[ch] + 2) >> 2 : acc[ch]);
      depth[(size_t)py * res + px] = fbuf[k0] >= 0 ? zbuf[k0] : 0.f;
    }
}

int raster_ref(const float* verts, const int32_t* faces, const uint8_t* colors, int V, int F, const float* poses,
               int B, float fx, float fy, float cx, float cy, int res, int msaa, int cull, const uint8_t* lut,
               uint8_t* rgb, float* depth) {
  if (msaa != 1 && msaa != 4) return -1;
  SV* sv = (SV*)malloc(sizeof(SV) * (size_t)V);
  float* zbuf = (float*)malloc(sizeof(float) * (size_t)res * res * msaa);
  int32_t* fbuf = (int32_t*)malloc(sizeof(int32_t) * (size_t)res * res * msaa);
  if (!sv || !zbuf || !fbuf) return -2;
  for (int b = 0; b < B; ++b)
    render_view(verts, faces, colors, V, F, poses + (size_t)b * 12, fx, fy, cx, cy, res, msaa, cull, lut,
                rgb + (size_t)b * res * res * 3, depth + (size_t)b * res * res, sv, zbuf, fbuf);
  free(sv); free(zbuf); free(fbuf);
  return 0;
}
