"""The pose-confidence half of reference ``src/pipeline/estimators/tracking_refiner.py`` (TrackingRefiner.__init__,
_render, _crop_image, _get_threshold_for_confidence, pose_confidence, n_inliers_per_pose: lines 19-100) on the B200
engine: roi_align photo crop at 518^2, one render per frame at the cropped intrinsics, DINOv2 ViT-B/14-reg on both,
masked per-patch cosine, histogram threshold.  SURVEY.md section 8f row 3.

The CoTracker / PnP refinement itself (the rest of the reference class) is outside the hot path and not provided.

Deliberate differences: the reference runs the ViT in fp32 (no autocast in scripts/smooth_poses_video.py:93-95); the
engine computes in bf16 like the rest of the path, so confidences agree to bf16 noise, not bit for bit.  The reference
re-seeds numpy's GLOBAL generator on every crop (tracking_refiner.py:47); here the same 100 vertex indices are drawn from
a private RandomState(42).
"""
from __future__ import annotations

import numpy as np
import torch

from ... import ops
from ..._lib import on_device
from ...vit_engine import ViTEngine
from ...vit_weights import VITB14_REG, load_state_dict_file, synthetic_state_dict
from .. import refiner_utils
from ..utils import as_mesh

bf16 = torch.bfloat16


class TrackingRefiner:
    def __init__(self, dino_model="dinov2_vitb14_reg", dino_device="cuda", cotracker_device="cuda", weights=None,
                 chunk: int = 32):
        if dino_model != "dinov2_vitb14_reg":
            raise ValueError("the confidence pass is built for dinov2_vitb14_reg (reference default)")
        if not torch.cuda.is_available() or torch.device(dino_device).type != "cuda":
            raise RuntimeError("TrackingRefiner needs a CUDA device (no CPU fallback)")
        self.dino_device = torch.device(dino_device)
        self.device = self.dino_device
        self.cotracker_device = cotracker_device
        if weights is None:
            import warnings
            warnings.warn("no DINOv2 ViT-B checkpoint given: using seeded synthetic weights")
            weights = synthetic_state_dict(VITB14_REG, seed=0)
        elif isinstance(weights, str):
            weights = load_state_dict_file(weights)
        self.dinov2 = ViTEngine(weights, VITB14_REG, device=self.dino_device, chunk=chunk)
        self.image_size = 518          # int(sqrt(1370 - 1) * 14), tracking_refiner.py:26
        self.patch_size = 14
        self.feats_size = self.image_size // self.patch_size   # 37
        self._target = (self.image_size + 3) // 4 * 4           # the rasteriser's targets are multiples of 4 px wide

    # ------------------------------------------------------------------ device pieces
    def _render_device(self, mesh, Ks: torch.Tensor, transforms: torch.Tensor):
        """Ks (n,3,3), transforms (n,4,4) host tensors -> rgb u8 (n,S,S,3), depth (n,S,S) on the device, S = 520; the
        518 x 518 image is the top-left corner.  pyrender set-up of tracking_refiner.py:31-45: ambient (5,5,5), znear
        1e-4 / zfar 9999, default render flags (back faces culled)."""
        m = as_mesh(mesh)
        dev = self.dino_device
        view_k = torch.stack([Ks[:, 0, 0], Ks[:, 1, 1], Ks[:, 0, 2], Ks[:, 1, 2]], dim=1).to(dev, torch.float32).contiguous()
        P = transforms.to(dev, torch.float32)
        return ops.rasterize_mesh(m, P, 1.0, 1.0, 0.0, 0.0, self._target, msaa=4, cull_backfaces=True, ambient=5.0,
                                  znear=1e-4, zfar=9999.0, view_k=view_k)

    def _crop_boxes(self, mesh, K, transforms):
        vertices = np.asarray(as_mesh(mesh).vertices)
        idx = np.random.RandomState(42).choice(np.arange(len(vertices)), 100)      # np.random.seed(42); np.random.choice
        pts = torch.from_numpy(np.pad(vertices[idx], ((0, 0), (0, 1)), constant_values=1.).copy()).float()
        Kt = torch.from_numpy(np.asarray(K, dtype=np.float64)).view(3, 3).float()
        Ts = torch.from_numpy(np.asarray(transforms, dtype=np.float64)).view(-1, 4, 4).float()
        boxes = refiner_utils.crop_boxes(Ts, pts, Kt, self.image_size, self.image_size)
        return boxes, refiner_utils.update_K_with_crop(Kt, boxes, self.image_size, self.image_size), Ts

    def _to_image_tensor(self, image):
        """PIL image / HWC u8 array / CHW float tensor -> (3,H,W) fp32 in [0,1] on the device (MaybeToTensor)."""
        if isinstance(image, torch.Tensor):
            return image.to(self.dino_device, torch.float32)
        arr = np.asarray(image)
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(self.dino_device)
        if t.dim() == 2:
            t = t[:, :, None]
        t = t.permute(2, 0, 1).contiguous()
        return t.float().div(255) if arr.dtype == np.uint8 else t.float()

    @on_device
    def pose_confidences(self, mesh, frames, K, transforms) -> torch.Tensor:
        """All frames in one batch: (n,37,37) fp32 on the device (row i = reference pose_confidence(mesh, frames[i], K,
        transforms[i]))."""
        n = len(frames)
        S, g = self.image_size, self.feats_size
        boxes, new_K, Ts = self._crop_boxes(mesh, K, np.stack([np.asarray(t) for t in transforms]))
        dev = self.dino_device
        photos = torch.empty(n, 3, S, S, dtype=torch.float32, device=dev)
        for i, frame in enumerate(frames):   # frames may differ in size; one roi_align per frame, one box each
            photos[i:i + 1] = ops.roi_align(self._to_image_tensor(frame), boxes[i:i + 1].to(dev), S, S, 2)
        rgb, depth = self._render_device(mesh, new_K, Ts)
        mask = ops.depth_mask_cubic(depth, g, res=S)
        full = torch.tensor([[0, 0, S, S]], dtype=torch.int32, device=dev).repeat(n, 1)
        patches, status = ops.crop_resize_pad(rgb, full, S, to_patches=True)      # u8 render -> normalised patch matrix
        L = self.dinov2.depth
        f_photo = self.dinov2.forward(photos, layer=L, feature_type="patch")
        f_render = self.dinov2.forward(patches, layer=L, feature_type="patch", res=S)
        return ops.patch_cosine(f_photo, f_render, mask.reshape(n, g * g)).view(n, g, g)

    # ------------------------------------------------------------------ reference-shaped API
    @on_device
    def _render(self, mesh, width, height, K, transform):
        assert width == self.image_size and height == self.image_size
        rgb, depth = self._render_device(mesh, torch.from_numpy(np.asarray(K, dtype=np.float64)).view(1, 3, 3).float(),
                                         torch.from_numpy(np.asarray(transform, dtype=np.float64)).view(1, 4, 4).float())
        return rgb[0, :height, :width].cpu().numpy(), depth[0, :height, :width].cpu().numpy()

    @on_device
    def _crop_image(self, mesh, image, K, transform):
        boxes, new_K, _ = self._crop_boxes(mesh, K, np.asarray(transform)[None])
        crop = ops.roi_align(self._to_image_tensor(image), boxes.to(self.dino_device), self.image_size, self.image_size, 2)
        return crop[0], boxes[0], new_K[0]

    def _get_threshold_for_confidence(self, similarity_matrices, top_quantile=0.2):
        """tracking_refiner.py:59-68: lower edge of the histogram bin (50 bins over the positive similarities) at which
        the count accumulated from the top exceeds ``top_quantile`` of all positive entries."""
        counts, edges = np.histogram(similarity_matrices[similarity_matrices > 0], bins=50)
        from_top = np.cumsum(counts[::-1])
        over = from_top > counts.sum() * top_quantile
        k = int(np.argmax(over)) if over.any() else len(counts) - 1     # the reference's loop falls through to bin 0
        return edges[:-1][::-1][k]

    @on_device
    def pose_confidence(self, mesh, photo, K, transform):
        return self.pose_confidences(mesh, [photo], K, [transform])[0].cpu().numpy()

    def n_inliers_per_pose(self, mesh, frames, K, transforms):
        confidences = self.pose_confidences(mesh, list(frames), K, list(transforms)).cpu().numpy()
        thr = self._get_threshold_for_confidence(confidences)
        return (confidences > thr).sum(-1).sum(-1), thr
