"""Drop-in for reference ``Proposals`` (src/pipeline/utils.py:18-69): full frame + per-proposal masks/boxes ->
masked, cropped, resized and padded query crops -- on the device."""
from __future__ import annotations

import numpy as np
import torch

from .bbox_utils import CropResizePad


def mask_to_rle(mask: np.ndarray) -> dict:
    """Uncompressed column-major run-length encoding ({'size': [h, w], 'counts': [...]}), the format
    sam2.utils.amg.mask_to_rle_pytorch emits and the reference stores in its BOP-style proposal JSON
    (src/pipeline/utils.py:62): runs alternate background/foreground starting with background."""
    h, w = mask.shape
    flat = np.asarray(mask, dtype=bool).T.reshape(-1)  # Fortran order
    change = np.nonzero(flat[1:] != flat[:-1])[0] + 1
    idx = np.concatenate(([0], change, [h * w]))
    counts = np.diff(idx).tolist()
    if flat.size and flat[0]:
        counts = [0] + counts
    return {"size": [h, w], "counts": counts}


def rle_to_mask(rle: dict) -> np.ndarray:
    h, w = rle["size"]
    flat = np.zeros(h * w, dtype=bool)
    pos, val = 0, False
    for c in rle["counts"]:
        flat[pos:pos + c] = val
        pos += c
        val = not val
    return flat.reshape(w, h).T


_unit_tables = {}


def _unit_table(dev: torch.device) -> torch.Tensor:
    key = str(dev)
    if key not in _unit_tables:
        _unit_tables[key] = (torch.arange(256, dtype=torch.uint8).float() / 255).to(dev)
    return _unit_tables[key]


class Proposals:
    def __init__(self, image, detections_output, target_size=350, scene_id=None, frame_id=None, bbox_extend=0.2,
                 mask_rgb=True, device="cuda"):
        dev = torch.device(device)
        img = torch.as_tensor(image)
        if img.dtype == torch.uint8:
            # the frame travels as bytes (a quarter of the fp32 H2D copy) and `.float() / 255` becomes a 256-entry table
            # filled by the very same host expression, so the values are the reference's bit for bit
            self.image = _unit_table(dev)[img.to(dev, non_blocking=True).long()].permute(2, 0, 1)
        else:
            self.image = (img.float() / 255).permute(2, 0, 1).to(dev)
        self.masks = torch.as_tensor(detections_output["masks"]).bool().to(dev)
        self.boxes = torch.as_tensor(detections_output["boxes"]).int()
        self.rgb_proposal_processor = CropResizePad(target_size=target_size, orig_size=(image.shape[0], image.shape[1]),
                                                    bbox_extend=bbox_extend)
        self.proposals, self.proposals_masks = self.extract_proposals(mask_rgb=mask_rgb)
        self.features = None
        self.scores = []
        self.meshes = []
        self.scene_id = scene_id
        self.frame_id = frame_id

    def extract_proposals(self, mask_rgb=True):
        n = len(self.masks)
        rgbs = self.image.unsqueeze(0).expand(n, -1, -1, -1)
        m = self.masks.unsqueeze(1)
        masked = (rgbs * m) if mask_rgb else rgbs
        crops = self.rgb_proposal_processor(masked.contiguous(), self.boxes)
        mask_imgs = m.expand(-1, 3, -1, -1).float().contiguous()
        crop_masks = self.rgb_proposal_processor(mask_imgs, self.boxes)[:, 0] > 0.5
        return crops, crop_masks

    def to_bop_dict(self):
        out = []
        masks = self.masks.cpu().numpy()
        for i in range(len(self.boxes)):
            b = self.boxes[i].cpu().numpy().tolist()
            out.append({"bbox": [b[0], b[1], b[2] - b[0], b[3] - b[1]], "segmentation": mask_to_rle(masks[i]),
                        "mesh": self.meshes[i], "score": self.scores[i], "scene_id": int(self.scene_id),
                        "image_id": int(self.frame_id), "time": 0.01})
        return out
