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
    """Minimal triangle mesh (the reference passes ``trimesh.Trimesh``; anything exposing ``vertices``,
    ``faces`` and optionally ``visual.vertex_colors`` is accepted through :func:`as_mesh`)."""

    vertices: np.ndarray                      # (V,3)
    faces: np.ndarray                         # (F,3) int
    vertex_colors: np.ndarray | None = None   # (V,3|4) u8
    _device_cache: dict = field(default_factory=dict, repr=False, compare=False)

    def apply_scale(self, s: float):
        self.vertices = np.asarray(self.vertices, dtype=np.float64) * s
        self._device_cache.clear()
        return self

    def copy(self):
        return Mesh(np.array(self.vertices), np.array(self.faces),
                    None if self.vertex_colors is None else np.array(self.vertex_colors))


def as_mesh(obj) -> Mesh:
    if isinstance(obj, Mesh):
        return obj
    if hasattr(obj, "vertices") and hasattr(obj, "faces"):
        colors = None
        vis = getattr(obj, "visual", None)
        if vis is not None and getattr(vis, "kind", None) == "vertex" and getattr(vis, "vertex_colors", None) is not None:
            colors = np.asarray(vis.vertex_colors)
        return Mesh(np.asarray(obj.vertices), np.asarray(obj.faces), colors)
    raise TypeError("textured meshes and point clouds are not rasterised by this build; pass a triangle mesh with "
                    "vertices/faces (and optional per-vertex colours)")


def mesh_to_device(mesh: Mesh, device):
    """fp32 vertices, int32 faces, u8 RGB colours on the device (white when the mesh has no colours, which is
    what pyrender's default material renders)."""
    key = str(device)
    if key not in mesh._device_cache:
        v = torch.from_numpy(np.ascontiguousarray(np.asarray(mesh.vertices, dtype=np.float32))).to(device)
        f = torch.from_numpy(np.ascontiguousarray(np.asarray(mesh.faces, dtype=np.int32))).to(device)
        if mesh.vertex_colors is None:
            c = torch.full((v.shape[0], 3), 255, dtype=torch.uint8, device=device)
        else:
            c = torch.from_numpy(np.ascontiguousarray(np.asarray(mesh.vertex_colors)[:, :3].astype(np.uint8))).to(device)
        if f.numel() and (int(f.min()) < 0 or int(f.max()) >= v.shape[0]):
            raise ValueError("mesh faces index outside the vertex array")
        mesh._device_cache[key] = (v, f, c)
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
