"""Drop-in for reference ``src/utils/bbox_utils.py:9-56`` (CropResizePad) running as one CUDA gather."""
from __future__ import annotations

from typing import Tuple, Union

import torch

from .. import ops


def extend_boxes(boxes: torch.Tensor, bbox_extend, w: int, h: int) -> torch.Tensor:
    """Reference bbox_utils.py:21-28: per-box extension with *truncating assignment into the integer box
    tensor* (``box[0] = max(0, box[0] - ext * box_w)`` on an int tensor) -- restated with the same dtype rules:
    the right-hand side is computed as a Python/tensor float and truncated toward zero on assignment."""
    boxes = boxes.clone()
    for box in boxes:
        box_w = box[2] - box[0]
        box_h = box[3] - box[1]
        box[0] = max(0, box[0] - bbox_extend * box_w)
        box[2] = min(w, box[2] + bbox_extend * box_w)
        box[1] = max(0, box[1] - bbox_extend * box_h)
        box[3] = min(h, box[3] + bbox_extend * box_h)
    return boxes


class CropResizePad:
    def __init__(self, target_size: Union[Tuple, int], orig_size: Union[Tuple, int], bbox_extend: int = 0):
        if isinstance(target_size, int):
            target_size = (target_size, target_size)
        if target_size[0] != target_size[1]:
            raise ValueError("only square targets are built (the reference only uses square targets)")
        self.target_size = target_size
        self.target_h, self.target_w = target_size
        self.bbox_extend = bbox_extend
        self.h, self.w = orig_size

    def __call__(self, images: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
        """images (B,3,H,W) float, boxes (B,4) int xyxy -> (B,3,T,T) fp32 on the images' device side of the
        engine (CUDA).  Raises like the reference when a box degenerates."""
        boxes_ext = extend_boxes(boxes.detach().cpu(), self.bbox_extend, self.w, self.h)
        dev = images.device if images.is_cuda else torch.device("cuda")
        src = images.to(dev, dtype=torch.float32).contiguous()
        out, status = ops.crop_resize_pad(src, boxes_ext.to(dev), self.target_h, to_patches=False)
        bad = int(status.item())
        if bad:
            raise RuntimeError(f"CropResizePad: box {bad - 1} = {boxes_ext[bad - 1].tolist()} is degenerate")
        return out
