"""Host-side geometry of the hot path: pose set, mesh container, translation from depth extents.

Mirrors the names of reference ``src/pipeline/utils.py`` (depthmap_to_pointcloud, get_z_from_pointcloud,
mask_to_bbox) and ``DinoPoseEstimator.generate_poses``; the estimators use the fused CUDA reductions
(``ops.depth_extents``) and only the O(1) scalar tail below runs on the host, in float64 like the reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

RENDERING_SCALE = 0.25  # reference render_templates.py:61 / online_pose_estimator.py:23


def generate_poses(n_poses: int = 600) -> list:
    """Super-Fibonacci SO(3) samples at distance 1.1 (reference pose_estimator.py:121-147, duplicated in
    renderer.py:13-35).  Quaternion (x,y,z,w) -> matrix uses the same expression order as
    scipy.spatial.transform.Rotation.from_quat(...).as_matrix()."""
    phi = np.sqrt(2.0)
    psi = 1.533751168755204288118041
    s = np.arange(n_poses, dtype=np.float64) + 0.5
    r = np.sqrt(s / n_poses)
    R = np.sqrt(1.0 - s / n_poses)
    alpha = 2.0 * np.pi * s / phi
    beta = 2.0 * np.pi * s / psi
    q = np.stack((r * np.sin(alpha), r * np.cos(alpha), R * np.sin(beta), R * np.cos(beta)), axis=1)
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    poses = np.zeros((n_poses, 4, 4), dtype=np.float64)
    poses[:, 0, 0] = x2 - y2 - z2 + w2
    poses[:, 1, 0] = 2 * (xy + zw)
    poses[:, 2, 0] = 2 * (xz - yw)
    poses[:, 0, 1] = 2 * (xy - zw)
    poses[:, 1, 1] = -x2 + y2 - z2 + w2
    poses[:, 2, 1] = 2 * (yz + xw)
    poses[:, 0, 2] = 2 * (xz + yw)
    poses[:, 1, 2] = 2 * (yz - xw)
    poses[:, 2, 2] = -x2 - y2 + z2 + w2
    poses[:, 2, 3] = 1.1
    poses[:, 3, 3] = 1.0
    return [p for p in poses]


@dataclass
class Mesh:
    """Minimal renderable (the reference passes ``trimesh.Trimesh`` / ``trimesh.PointCloud``; anything exposing
    ``vertices`` plus ``faces`` / ``visual`` or ``colors`` is accepted through :func:`as_mesh`).

    ``faces is None`` marks a point cloud (reference renderer.py:46-51).  A base-colour texture is given as ``uv``
    (V,2, v up) + ``texture`` (H,W,3|4 u8, row 0 = top of the image, i.e. v = 1)."""

    vertices: np.ndarray                      # (V,3)
    faces: np.ndarray | None                  # (F,3) int, None = point cloud
    vertex_colors: np.ndarray | None = None   # (V,3|4) u8
    uv: np.ndarray | None = None              # (V,2) float
    texture: np.ndarray | None = None         # (H,W,3|4) u8
    _device_cache: dict = field(default_factory=dict, repr=False, compare=False)

    @property
    def is_point_cloud(self) -> bool:
        return self.faces is None

    def apply_scale(self, s: float):
        self.vertices = np.asarray(self.vertices, dtype=np.float64) * s
        self._device_cache.clear()
        return self

    def copy(self):
        cp = lambda a: None if a is None else np.array(a)
        return Mesh(np.array(self.vertices), cp(self.faces), cp(self.vertex_colors), cp(self.uv), cp(self.texture))


def _texture_image(material):
    """The base-colour image of a trimesh material (SimpleMaterial.image / PBRMaterial.baseColorTexture) as an
    (H,W,3|4) u8 array, or None."""
    for name in ("image", "baseColorTexture"):
        img = getattr(material, name, None)
        if img is None:
            continue
        if hasattr(img, "convert"):          # PIL image
            img = img.convert("RGBA")
        arr = np.asarray(img)
        if arr.ndim == 2:
            arr = np.repeat(arr[:, :, None], 3, axis=2)
        if arr.ndim == 3 and arr.shape[2] in (3, 4) and arr.dtype == np.uint8:
            return arr
    return None


def as_mesh(obj) -> Mesh:
    """trimesh-like object -> :class:`Mesh`, following what ``pyrender.Mesh.from_trimesh`` / ``from_points`` read:
    ``visual.kind == 'vertex'`` -> vertex colours, ``'texture'`` -> uv + material image, point clouds -> ``colors``
    (white when empty, renderer.py:47-50)."""
    if isinstance(obj, Mesh):
        return obj
    if hasattr(obj, "vertices") and getattr(obj, "faces", None) is not None:
        colors = uv = tex = None
        vis = getattr(obj, "visual", None)
        kind = getattr(vis, "kind", None)
        if kind == "vertex" and getattr(vis, "vertex_colors", None) is not None:
            colors = np.asarray(vis.vertex_colors)
        elif kind == "texture" and getattr(vis, "uv", None) is not None:
            tex = _texture_image(getattr(vis, "material", None))
            if tex is not None:
                uv = np.asarray(vis.uv, dtype=np.float64)
        return Mesh(np.asarray(obj.vertices), np.asarray(obj.faces), colors, uv, tex)
    if hasattr(obj, "vertices"):
        colors = getattr(obj, "colors", None)
        colors = None if colors is None or np.size(colors) == 0 else np.asarray(colors)
        return Mesh(np.asarray(obj.vertices), None, colors)
    raise TypeError("expected a triangle mesh (vertices/faces[/visual]) or a point cloud (vertices[/colors])")


def build_mip_chain(image: np.ndarray):
    """(H,W,3|4) u8 -> (flat RGBA8 bytes of every level, level 0 first; number of levels).  Each level halves both
    sides (floor, min 1) with a 2x2 box filter in the stored (sRGB) values, rounded half up -- what glGenerateMipmap
    does for pyrender's non-sRGB GL_RGBA textures (odd sides: the last source texel is reused)."""
    img = np.asarray(image)
    if img.shape[2] == 3:
        img = np.concatenate([img, np.full(img.shape[:2] + (1,), 255, np.uint8)], axis=2)
    levels = [np.ascontiguousarray(img, dtype=np.uint8)]
    while levels[-1].shape[0] > 1 or levels[-1].shape[1] > 1:
        src = levels[-1].astype(np.uint16)
        H, W = src.shape[:2]
        nh, nw = max(1, H // 2), max(1, W // 2)
        y0, y1 = np.minimum(2 * np.arange(nh), H - 1), np.minimum(2 * np.arange(nh) + 1, H - 1)
        x0, x1 = np.minimum(2 * np.arange(nw), W - 1), np.minimum(2 * np.arange(nw) + 1, W - 1)
        acc = src[y0][:, x0] + src[y0][:, x1] + src[y1][:, x0] + src[y1][:, x1]
        levels.append(((acc + 2) >> 2).astype(np.uint8))
    return np.concatenate([l.reshape(-1) for l in levels]), len(levels)


def srgb_to_linear_lut() -> np.ndarray:
    """[65536] fp32 (i/65535)^2.2: pyrender's ``srgb_to_linear`` (mesh.frag), applied to the filtered texel."""
    return np.power(np.arange(65536, dtype=np.float64) / 65535.0, 2.2).astype(np.float32)


def mesh_to_device(mesh: Mesh, device):
    """fp32 vertices, int32 faces, u8 RGB colours on the device (white when the mesh has neither colours nor a
    texture, which is what pyrender's default material renders; None when a texture replaces them)."""
    key = str(device)
    if key not in mesh._device_cache:
        v = torch.from_numpy(np.ascontiguousarray(np.asarray(mesh.vertices, dtype=np.float32))).to(device)
        if mesh.faces is None:
            f = torch.zeros((0, 3), dtype=torch.int32, device=device)
        else:
            f_host = np.ascontiguousarray(np.asarray(mesh.faces, dtype=np.int32))
            if f_host.size and (f_host.min() < 0 or f_host.max() >= v.shape[0]):   # validated once, on the host
                raise ValueError("mesh faces index outside the vertex array")
            f = torch.from_numpy(f_host).to(device)
        if mesh.vertex_colors is not None:
            c = torch.from_numpy(np.ascontiguousarray(np.asarray(mesh.vertex_colors)[:, :3].astype(np.uint8))).to(device)
        elif mesh.texture is not None and mesh.uv is not None:
            c = None
        else:
            c = torch.full((v.shape[0], 3), 255, dtype=torch.uint8, device=device)
        mesh._device_cache[key] = (v, f, c)
    return mesh._device_cache[key]


def mesh_texture_to_device(mesh: Mesh, device):
    """-> dict(uv (V,2) fp32, chain u8, w, h, levels) on the device, or None for untextured meshes."""
    if mesh.texture is None or mesh.uv is None or mesh.faces is None:
        return None
    key = "tex:" + str(device)
    if key not in mesh._device_cache:
        chain, levels = build_mip_chain(mesh.texture)
        mesh._device_cache[key] = dict(
            uv=torch.from_numpy(np.ascontiguousarray(np.asarray(mesh.uv, dtype=np.float32))).to(device),
            chain=torch.from_numpy(chain).to(device), w=int(mesh.texture.shape[1]), h=int(mesh.texture.shape[0]),
            levels=levels)
    return mesh._device_cache[key]


# -------------------------------------------------------------------------------------------------
# Reference-named helpers (numpy, float64 -- same arithmetic as src/pipeline/utils.py:122-181)
# -------------------------------------------------------------------------------------------------
def depthmap_to_pointcloud(depth_map, K):
    """Back-project every pixel with K^-1 and drop all-zero rows (reference utils.py:122-145)."""
    K_inv = np.linalg.inv(K)
    h, w = depth_map.shape[:2]
    u, v = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
    hom = np.stack((u, v, np.ones_like(u)), axis=2).reshape(-1, 3)
    pc = (np.dot(K_inv, hom.T) * depth_map.flatten()).T
    return pc[~np.all(pc == 0, axis=1)]


def tco_from_extents(bbox, deltax_3d, deltay_3d, K, TCO_init):
    """Tail of reference get_z_from_pointcloud (utils.py:148-170) given the cloud's X/Y extents."""
    bbox = np.asarray(bbox)
    TCO = np.array(TCO_init, dtype=np.float64, copy=True)
    K = np.asarray(K)
    fxfy = K[[0, 1], [0, 1]]
    cxcy = K[[0, 1], [2, 2]]
    bb_xy_centers = (bbox[0:2] + bbox[2:4]) / 2
    bb_deltax = (bbox[2] - bbox[0]) + 1
    bb_deltay = (bbox[3] - bbox[1]) + 1
    z_from_dx = fxfy[0] * deltax_3d / bb_deltax
    z_from_dy = fxfy[1] * deltay_3d / bb_deltay
    z = (z_from_dy + z_from_dx) / 2
    TCO[:2, 3] = ((bb_xy_centers - cxcy) * z) / fxfy
    TCO[2, 3] = z
    return TCO


def get_z_from_pointcloud(bbox, pointcloud, K, TCO_init):
    dx = pointcloud[:, 0].max() - pointcloud[:, 0].min()
    dy = pointcloud[:, 1].max() - pointcloud[:, 1].min()
    return tco_from_extents(bbox, dx, dy, K, TCO_init)


def rescaled_extents(ext_row: np.ndarray, est_scale: float, recentre: bool):
    """X/Y extents of the point cloud after the reference's rescaling, from one fp_depth_extents row
    (xmin,xmax,ymin,ymax,sum_x,sum_y,sum_z,count).

    coarse (pose_estimator.py:104-111): pc -= mean; pc /= 0.25; pc *= est_scale; pc += mean
    fine   (online_pose_estimator.py:82-85): pc /= 0.25; pc *= est_scale
    Both are increasing maps for est_scale > 0, so min/max commute with them and the same float64 operations
    are applied to the extreme coordinates."""
    xmin, xmax, ymin, ymax, sx, sy, _, cnt = [np.float64(v) for v in ext_row]
    if cnt == 0:
        raise ValueError("empty depth map: the selected view has no rendered pixels")
    if recentre:
        mx, my = sx / cnt, sy / cnt
        f = lambda v, m: ((v - m) / RENDERING_SCALE) * est_scale + m
        xs = (f(xmin, mx), f(xmax, mx))
        ys = (f(ymin, my), f(ymax, my))
    else:
        f = lambda v: (v / RENDERING_SCALE) * est_scale
        xs = (f(xmin), f(xmax))
        ys = (f(ymin), f(ymax))
    return max(xs) - min(xs), max(ys) - min(ys)


def mask_to_bbox(mask):
    ys, xs = np.nonzero(mask)
    return np.array([xs.min(), ys.min(), xs.max(), ys.max()])
