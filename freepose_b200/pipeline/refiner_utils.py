"""Host geometry of the refiner's confidence pass, named after reference ``src/pipeline/refiner_utils.py``: the crop box
around the projected object (``crop_image``, :92-137) and the intrinsics of the cropped view (``update_K_with_crop``,
:140-176).  Both are O(100)-point fp32 torch expressions evaluated on the host exactly as the reference does; the
resampling itself (torchvision ``roi_align`` there) is ``ops.roi_align`` on the device."""
from __future__ import annotations

import torch

from .. import ops


def crop_boxes(Ts: torch.Tensor, points: torch.Tensor, K: torch.Tensor, render_width: int, render_height: int,
               lamb: float = 1.4) -> torch.Tensor:
    """(n,4,4) poses, (m,4) homogeneous object points, (3,3) K -> (n,4) fp32 crop boxes x1,y1,x2,y2
    (reference crop_image, refiner_utils.py:98-124)."""
    assert Ts.shape[1:] == (4, 4) and points.shape[1:] == (4,) and K.shape == (3, 3)
    T = torch.matmul(torch.nn.functional.pad(K, (0, 1, 0, 0), value=0.).unsqueeze(0), Ts)
    pts = torch.matmul(points.unsqueeze(0), T.permute(0, 2, 1))
    uv = pts[:, :, :2] / torch.maximum(pts[:, :, [2]], torch.tensor(0.01))
    bboxes = torch.cat([uv.min(dim=1).values, uv.max(dim=1).values], dim=1)
    ctr = torch.matmul(torch.mean(points, dim=0, keepdim=True).unsqueeze(0), T.permute(0, 2, 1)).squeeze(1)
    ctr_uv = ctr[:, :2] / torch.maximum(ctr[:, [2]], torch.tensor(0.01))
    dists = torch.maximum((bboxes[:, [0, 1]] - ctr_uv).abs_(), (bboxes[:, [2, 3]] - ctr_uv).abs_())
    xd, yd = dists[:, 0], dists[:, 1]
    r = render_width / render_height
    width = torch.max(xd, yd * r) * 2 * lamb
    height = torch.max(xd / r, yd) * 2 * lamb
    x1, y1 = ctr_uv[:, 0] - width / 2, ctr_uv[:, 1] - height / 2
    x2, y2 = ctr_uv[:, 0] + width / 2, ctr_uv[:, 1] + height / 2
    return torch.stack([x1, y1, x2, y2], dim=1)


def crop_image(image: torch.Tensor, Ts, points, K, render_width, render_height, lamb=1.4):
    """Reference signature (refiner_utils.py:92): (C,H,W) fp32 image -> (crops (n,C,h,w) on the image's device, boxes).
    The image must live on the GPU (no CPU fallback)."""
    assert image.dim() == 3 and image.shape[0] in (1, 3, 4) and image.dtype == torch.float32
    boxes = crop_boxes(Ts, points, K, render_width, render_height, lamb)
    crops = ops.roi_align(image, boxes.to(image.device), render_height, render_width, sampling_ratio=2)
    return crops, boxes


def update_K_with_crop(K: torch.Tensor, bboxes: torch.Tensor, render_width: int, render_height: int) -> torch.Tensor:
    """Intrinsics after cropping to ``bboxes`` and resizing to (render_width, render_height); skew is not handled
    (reference refiner_utils.py:140-176)."""
    assert K.shape == (3, 3) and bboxes.shape[1:] == (4,)
    new_K = K.unsqueeze(0).repeat(len(bboxes), 1, 1)
    cw, ch = bboxes[:, 2] - bboxes[:, 0], bboxes[:, 3] - bboxes[:, 1]
    ccx, ccy = (bboxes[:, 0] + bboxes[:, 2]) / 2, (bboxes[:, 1] + bboxes[:, 3]) / 2
    cx = K[0, 2] + (cw - 1) / 2 - ccx
    cy = K[1, 2] + (ch - 1) / 2 - ccy
    dx, dy = cx - (cw - 1) / 2, cy - (ch - 1) / 2
    sx, sy = render_width / cw, render_height / ch
    new_K[:, 0, 0] = sx * K[0, 0]
    new_K[:, 1, 1] = sy * K[1, 1]
    new_K[:, 0, 2] = (render_width - 1) / 2 + sx * dx
    new_K[:, 1, 2] = (render_height - 1) / 2 + sy * dy
    return new_K
