"""Drop-in for reference ``src/pipeline/retrieval/renderer.py:11-129`` backed by the CUDA rasteriser."""
from __future__ import annotations

import numpy as np
import torch

from ... import ops
from ..._lib import on_device
from ..bbox_utils import CropResizePad
from ..utils import Mesh, as_mesh, generate_poses, mask_to_bbox, mesh_to_device


class MeshRenderer:
    """Same surface as the reference class: ``render``, ``render_from_poses``, ``mask_to_bbox``,
    ``generate_proposals``; plus ``*_device`` variants that keep everything in HBM (used by the estimators).

    Camera: ``pyrender.IntrinsicsCamera(fx=600, fy=600, cx=res/2, cy=res/2)`` (renderer.py:37).  ``focal``
    defaults to 600 at the reference's 420 px and scales with the resolution otherwise.
    """

    def __init__(self, n_poses, resolution=420, focal=None, msaa=4, device="cuda"):
        self.mesh_poses = generate_poses(n_poses)
        self.rotations = [p[:3, :3] for p in self.mesh_poses]
        self.resolution = resolution
        self.focal = float(focal) if focal is not None else 600.0 * resolution / 420.0
        self.msaa = msaa
        self.device = torch.device(device)
        self._poses_dev = None

    # ------------------------------------------------------------------ device path
    @on_device
    def render_device(self, mesh, poses=None, cull_faces=False):
        """-> rgb u8 (B,res,res,3), depth fp32 (B,res,res) CUDA tensors."""
        m = as_mesh(mesh)
        P = self.poses_device(poses)
        r = self.resolution
        return ops.rasterize_mesh(m, P, self.focal, self.focal, r / 2, r / 2, r, msaa=self.msaa,
                                  cull_backfaces=cull_faces)

    def poses_device(self, poses=None) -> torch.Tensor:
        """(B,4,4) fp32 on the device: the renderer's own hypothesis set (uploaded once), a list / array of 4x4
        matrices, or a device tensor passed through."""
        if poses is None:
            if self._poses_dev is None:
                self._poses_dev = torch.from_numpy(np.array(self.mesh_poses)).to(self.device, torch.float32)
            return self._poses_dev
        if torch.is_tensor(poses):
            return poses.to(self.device, torch.float32, non_blocking=True)
        # (a non-blocking copy from pageable memory is staged by the driver before the call returns: safe for H2D, and
        # the host does not wait for the work already queued on the stream)
        return torch.as_tensor(np.asarray(poses), dtype=torch.float32).to(self.device, non_blocking=True)

    @on_device
    def proposals_device(self, rgb, depth, resolution=None, to_patches=True, out=None, want_mask=True):
        """Device version of generate_proposals: mask -> bbox -> CropResizePad.  Returns
        (patch matrix | fp32 crops, bbox (B,4) int32, masks u8 (B,res,res) | None, status).  ``want_mask=False`` skips
        writing the masks (a second pass over the depth maps) for callers that only need the crops."""
        res = rgb.shape[1]
        T = resolution or res
        lo, hi = (105, 315) if res == 420 else (res // 4, res - res // 4)  # renderer.py:117 is hard-coded for 420
        bbox, count, mask = ops.mask_bbox(depth, fallback=(lo, hi), min_count=100, return_mask=want_mask)
        out, status = ops.crop_resize_pad(rgb, bbox, T, to_patches=to_patches, out=out)
        return out, bbox, mask, status

    # ------------------------------------------------------------------ reference-shaped API (host results)
    def render(self, mesh, cull_faces=False):
        rgb, depth = self.render_device(mesh, None, cull_faces)
        rgb, depth = rgb.cpu().numpy(), depth.cpu().numpy()
        return [(rgb[i], depth[i], self.mesh_poses[i][:3, :3]) for i in range(len(self.mesh_poses))]

    def render_from_poses(self, mesh, poses, cull_faces=False):
        rgb, depth = self.render_device(mesh, poses, cull_faces)
        rgb, depth = rgb.cpu().numpy(), depth.cpu().numpy()
        return [(rgb[i], depth[i], poses[i]) for i in range(len(poses))]

    mask_to_bbox = staticmethod(mask_to_bbox)

    @staticmethod
    def generate_proposals(res, resolution=420, bbox_extend=0):
        """(rgb, depth, pose) list -> (templates (B,3,T,T) fp32 CPU, poses, masks) like renderer.py:110-129."""
        imgs, boxes, poses, masks = [], [], [], []
        for img, depth, pose in res:
            mask = depth > 0
            if mask.sum() < 100:
                mask[105:315, 105:315] = True
            boxes.append(mask_to_bbox(mask))
            imgs.append(torch.from_numpy(img / 255).float())
            poses.append(pose)
            masks.append(mask)
        templates = torch.stack(imgs).permute(0, 3, 1, 2)
        h, w = templates.shape[-2:]
        proc = CropResizePad(resolution, (h, w), bbox_extend=bbox_extend)
        return proc(templates, torch.tensor(np.array(boxes))).cpu(), poses, masks
