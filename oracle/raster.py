"""ORACLE (test infrastructure): ctypes wrapper of oracle/raster_ref.c (see its header for scope and the
"parity unpinned" note for row R)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "libraster_ref.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _DIR / "raster_ref.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        _SO.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", str(src), "-o",
                        str(_SO), "-lm"], check=True)
    return _SO


def gamma_lut() -> np.ndarray:
    i = np.arange(65536, dtype=np.float64) / 65535.0
    return np.floor(255.0 * np.power(i, 1.0 / 2.2) + 0.5).astype(np.uint8)


def render(verts, faces, colors, poses, fx, fy, cx, cy, res, msaa=4, cull=False, uv=None, texture=None, points=False,
           ambient=0.0, znear=0.0, zfar=0.0, view_k=None):
    """verts (V,3), faces (F,3) | None, colors (V,3) u8 | None, poses (B,4,4) -> rgb u8 (B,res,res,3), depth f32
    (B,res,res).  ``texture`` = (RGBA8 mip chain bytes, w, h, levels) with ``uv`` (V,2); ``points`` = 1-px sprites."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.raster_ref3.restype = C.c_int
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    faces = np.zeros((0, 3), np.int32) if faces is None else np.ascontiguousarray(faces, dtype=np.int32)
    colors = None if colors is None else np.ascontiguousarray(np.asarray(colors)[:, :3], dtype=np.uint8)
    p = np.ascontiguousarray(np.asarray(poses, dtype=np.float32)[:, :3, :4])
    B = p.shape[0]
    lut = gamma_lut()
    rgb = np.zeros((B, res, res, 3), dtype=np.uint8)
    depth = np.zeros((B, res, res), dtype=np.float32)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    chain, tw, th, tl = texture if texture is not None else (None, 0, 0, 0)
    if texture is not None:
        chain = np.ascontiguousarray(chain, dtype=np.uint8)
        uv = np.ascontiguousarray(uv, dtype=np.float32)
        slut = np.power(np.arange(65536, dtype=np.float64) / 65535.0, 2.2).astype(np.float32)
    else:
        uv = slut = None
    if view_k is not None:
        view_k = np.ascontiguousarray(view_k, dtype=np.float32).reshape(B, 4)
    rc = _lib.raster_ref3(vp(verts), vp(faces), vp(colors), C.c_int(verts.shape[0]), C.c_int(faces.shape[0]), vp(p),
                          C.c_int(B), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_int(res),
                          C.c_int(msaa), C.c_int(int(cull)), vp(lut), vp(rgb), vp(depth), C.c_int(int(points)), vp(uv),
                          vp(chain), C.c_int(tw), C.c_int(th), C.c_int(tl), vp(slut), C.c_float(ambient), C.c_float(znear),
                          C.c_float(zfar), vp(view_k))
    if rc != 0:
        raise RuntimeError(f"raster_ref failed: {rc}")
    return rgb, depth


def render_mesh(mesh, poses, fx, fy, cx, cy, res, msaa=4, cull=False, **kw):
    """Any freepose_b200.pipeline.utils.Mesh (vertex colours, texture, point cloud) through the C restatement."""
    from freepose_b200.pipeline.utils import build_mip_chain
    colors = mesh.vertex_colors
    if mesh.texture is not None and mesh.uv is not None and mesh.faces is not None:
        chain, levels = build_mip_chain(mesh.texture)
        return render(mesh.vertices, mesh.faces, colors, poses, fx, fy, cx, cy, res, msaa, cull, uv=mesh.uv,
                      texture=(chain, mesh.texture.shape[1], mesh.texture.shape[0], levels), **kw)
    if colors is None:
        colors = np.full((len(mesh.vertices), 3), 255, np.uint8)
    return render(mesh.vertices, mesh.faces, colors, poses, fx, fy, cx, cy, res, msaa, cull, points=mesh.faces is None, **kw)
