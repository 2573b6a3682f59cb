"""Seeded synthetic inputs for tests and the benchmark: there is no network, so no
Objaverse meshes, no DINOv2 checkpoint and no BOP images -- shapes and value ranges match the real ones."""
from __future__ import annotations

import numpy as np

from .pipeline.utils import Mesh


def icosphere(subdivisions: int = 4):
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(subdivisions):
        cache = {}
        verts = list(v)

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = (verts[a] + verts[b]) / 2.0
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        nf = []
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        v, f = np.array(verts), np.array(nf, dtype=np.int64)
    return v, f


def synthetic_mesh(seed: int = 0, subdivisions: int = 5, scale: float = 0.25) -> Mesh:
    """Closed bumpy blob, ~20k faces at subdivisions=5, normalised like reference resize_meshes.py:18-23
    (bbox centred, max half-extent 1) and then scaled by the rendering scale 0.25.  Vertex colours are smooth,
    seeded and kept below 128 so that the reference's x2 ambient light does not saturate them."""
    rng = np.random.default_rng(seed)
    v, f = icosphere(subdivisions)
    # low-frequency radial bumps make the silhouette and shading pose-dependent
    dirs = rng.normal(size=(6, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    amp = rng.uniform(0.15, 0.45, size=6)
    radial = 1.0 + sum(a * np.maximum(v @ d, 0.0) ** 2 for a, d in zip(amp, dirs))
    v = v * radial[:, None] * rng.uniform(0.6, 1.0, size=3)[None, :]
    lo, hi = v.min(0), v.max(0)
    v = v - (lo + hi) / 2.0
    v = v / np.abs(v).max()
    cdirs = rng.normal(size=(3, 3))
    col = 20.0 + 100.0 * (0.5 + 0.5 * np.sin(3.0 * (v @ cdirs.T) + rng.uniform(0, 6.28, size=3)))
    return Mesh(v * scale, f, np.clip(np.rint(col), 0, 255).astype(np.uint8))


def synthetic_texture(seed: int = 0, width: int = 512, height: int = 256) -> np.ndarray:
    """(H,W,3) u8: smooth colour ramps + stripes + seeded per-texel noise (the noise is what makes mip selection and
    bilinear weights visible in the output)."""
    rng = np.random.default_rng(seed)
    v, u = np.meshgrid(np.linspace(0, 1, height, endpoint=False), np.linspace(0, 1, width, endpoint=False), indexing="ij")
    base = np.stack([0.5 + 0.5 * np.sin(2 * np.pi * (3 * u + v)), u, 0.5 + 0.5 * np.cos(2 * np.pi * 5 * v)], axis=2)
    stripes = ((np.floor(u * 32) + np.floor(v * 16)) % 2)[:, :, None]
    img = 0.55 * base + 0.25 * stripes + 0.2 * rng.uniform(size=(height, width, 3))
    return np.clip(np.rint(255 * img), 0, 255).astype(np.uint8)


def synthetic_textured_mesh(seed: int = 0, subdivisions: int = 4, tex_size=(512, 256), with_vertex_colors=False) -> Mesh:
    """The bumpy blob of :func:`synthetic_mesh` with spherical texture coordinates (longitude, latitude of the vertex
    direction; the seam triangles wrap around in u, which exercises REPEAT addressing and large footprints)."""
    m = synthetic_mesh(seed, subdivisions)
    d = m.vertices / np.linalg.norm(m.vertices, axis=1, keepdims=True)
    uv = np.stack([0.5 + np.arctan2(d[:, 1], d[:, 0]) / (2 * np.pi), 0.5 + np.arcsin(np.clip(d[:, 2], -1, 1)) / np.pi], axis=1)
    tex = synthetic_texture(seed, *tex_size)
    return Mesh(m.vertices, m.faces, m.vertex_colors if with_vertex_colors else None, uv, tex)


def synthetic_point_cloud(seed: int = 0, n: int = 20000, scale: float = 0.25) -> Mesh:
    """Seeded coloured points on the surface of the blob (a stand-in for reconstructed point-cloud models)."""
    rng = np.random.default_rng(seed)
    m = synthetic_mesh(seed, 4, scale)
    tri = m.vertices[m.faces[rng.integers(0, len(m.faces), size=n)]]
    w = rng.dirichlet(np.ones(3), size=n)
    pts = (tri * w[:, :, None]).sum(1)
    col = np.clip(np.rint(60 + 60 * (1 + np.sin(20 * pts @ rng.normal(size=(3, 3))))), 0, 255).astype(np.uint8)
    return Mesh(pts, None, col)


def camera_for(resolution: int):
    """Reference template camera K = [[600,0,210],[0,600,210],[0,0,1]] at 420 px (renderer.py:37,
    template.py:96), scaled with the resolution so the object fills the same image fraction."""
    f = 600.0 * resolution / 420.0
    return f, f, resolution / 2.0, resolution / 2.0
