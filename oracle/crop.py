"""ORACLE (test infrastructure): numpy restatement of reference ``CropResizePad.__call__``
(src/utils/bbox_utils.py:20-56) and ``MeshRenderer.generate_proposals`` (src/pipeline/retrieval/renderer.py:110-129).
Pinned against the reference's own class (imported unmodified) in tests/golden/make_golden.py -> the committed
fixtures tests/golden/crop_*.npz.

torch nearest interpolation with ``scale_factor`` (what the reference calls): out = floor(in * s) (double),
src = min(floor(float32(dst) * float32(1/s)), in-1).  (torch additionally short-cuts to the identity when
out_h + out_w <= 128; targets here are >= 224 so that path is never taken -- see DESIGN.md.)
"""
from __future__ import annotations

import math

import numpy as np


def nearest_indices(in_size: int, scale: float):
    out = int(math.floor(float(in_size) * scale))
    inv = np.float32(1.0 / scale)
    idx = np.floor(np.arange(out, dtype=np.float32) * inv).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def extend_box(box, bbox_extend, w, h):
    """bbox_utils.py:21-28 on an integer box.  ``bbox_extend * box_w`` is Python-scalar x int64 0-d tensor ->
    float32; the sum/difference stays float32; assignment into the int64 box truncates toward zero."""
    x1, y1, x2, y2 = [int(v) for v in box]
    f = np.float32
    bw, bh = x2 - x1, y2 - y1   # box_w / box_h are read BEFORE the box is modified (they are 0-d copies)

    def trunc(v):
        return int(v)  # toward zero

    if isinstance(bbox_extend, int):
        ew, eh = bbox_extend * bw, bbox_extend * bh  # int path (bbox_extend=0 default): exact
        nx1, nx2 = max(0, x1 - ew), min(w, x2 + ew)
        ny1, ny2 = max(0, y1 - eh), min(h, y2 + eh)
        return nx1, ny1, nx2, ny2
    ew = f(f(bbox_extend) * f(bw))
    eh = f(f(bbox_extend) * f(bh))
    nx1 = trunc(max(f(0), f(f(x1) - ew)))
    nx2 = trunc(min(f(w), f(f(x2) + ew)))
    ny1 = trunc(max(f(0), f(f(y1) - eh)))
    ny2 = trunc(min(f(h), f(f(y2) + eh)))
    return nx1, ny1, nx2, ny2


def crop_resize_pad(images: np.ndarray, boxes, target: int, bbox_extend=0, orig_size=None):
    """images (B,3,H,W) float32, boxes (B,4) int -> (B,3,target,target) float32."""
    B, _, H, W = images.shape
    oh, ow = orig_size if orig_size is not None else (H, W)
    out = np.zeros((B, 3, target, target), dtype=np.float32)
    for b in range(B):
        x1, y1, x2, y2 = extend_box(boxes[b], bbox_extend, ow, oh)
        # ``target / int64_tensor`` is Tensor.__rtruediv__ = reciprocal(tensor) * target, evaluated in float32
        scale = float(np.float32(np.float32(1.0) / np.float32(max(x2 - x1, y2 - y1))) * np.float32(target))
        img = images[b][:, y1:y2, x1:x2]
        iy = nearest_indices(img.shape[1], scale)
        ix = nearest_indices(img.shape[2], scale)
        img = img[:, iy][:, :, ix]
        h1, w1 = img.shape[1:]
        if w1 != h1:
            pt = max((target - h1) // 2, 0)
            pl = max((target - w1) // 2, 0)
            padded = np.zeros((3, target, target), dtype=np.float32)
            padded[:, pt:pt + h1, pl:pl + w1] = img
            img = padded
        assert img.shape[1] == img.shape[2]
        s2 = target / img.shape[1]
        i2 = nearest_indices(img.shape[1], s2)
        img = img[:, i2][:, :, i2]
        out[b] = img
    return out


def mask_to_bbox(mask):
    ys, xs = np.nonzero(mask)
    return np.array([xs.min(), ys.min(), xs.max(), ys.max()])


def generate_proposals(renders, resolution: int, fallback=(105, 315)):
    """renders: list of (rgb u8 HWC, depth f32) -> (templates (B,3,T,T) f32 in [0,1], boxes, masks)."""
    imgs, boxes, masks = [], [], []
    for rgb, depth in renders:
        mask = depth > 0
        if mask.sum() < 100:
            mask = mask.copy()
            mask[fallback[0]:fallback[1], fallback[0]:fallback[1]] = True
        boxes.append(mask_to_bbox(mask))
        imgs.append((rgb / 255).astype(np.float32))
        masks.append(mask)
    x = np.stack(imgs).transpose(0, 3, 1, 2)
    return crop_resize_pad(x, np.array(boxes), resolution), np.array(boxes), np.array(masks)
