"""ORACLE (test infrastructure, never imported by the product): CPU restatement of the refiner's pose-confidence pass,
reference src/pipeline/estimators/tracking_refiner.py:45-100 and src/pipeline/refiner_utils.py:92-176.

  crop box + roi_align (torchvision's own CPU op, the one the reference calls)      refiner_utils.py:92-137
  intrinsics of the crop                                                           refiner_utils.py:140-176
  render at the cropped K (oracle/raster_ref.c: ambient 5, znear 1e-4, zfar 9999)  tracking_refiner.py:31-45
  37x37 validity mask = cv2.resize(depth > 0, INTER_CUBIC) > 0.5 (cv2 itself)      tracking_refiner.py:75
  ViT-B/14-reg x_norm_patchtokens (oracle/vit.py) on photo crop and render         tracking_refiner.py:77-86
  masked per-patch cosine, histogram threshold                                     tracking_refiner.py:59-68,88-100

Pinned by tests/golden/refiner.npz: the reference's OWN ``TrackingRefiner.pose_confidence``, ``crop_image``,
``update_K_with_crop`` and ``_get_threshold_for_confidence`` executed unmodified (tests/golden/make_golden.py), with
``torch.hub.load`` returning the hub-shaped oracle ViT and ``_render`` (pyrender: not installable) replaced by the C
raster restatement -- so everything except the rasteriser is the reference's arithmetic.  PARITY UNPINNED for the render
step, as for row R."""
from __future__ import annotations

import cv2
import numpy as np
import torch
import torchvision

from freepose_b200.vit_weights import VITB14_REG

from . import raster as oraster
from .vit import OracleViT

IMAGE_SIZE, PATCH, G = 518, 14, 37


def crop_boxes(Ts, points, K, render_width, render_height, lamb=1.4):
    T = torch.matmul(torch.nn.functional.pad(K, (0, 1, 0, 0), value=0.).unsqueeze(0), Ts)
    pts = torch.matmul(points.unsqueeze(0), T.permute(0, 2, 1))
    uv = pts[:, :, :2] / torch.maximum(pts[:, :, [2]], torch.tensor(0.01))
    bboxes = torch.cat([uv.min(dim=1).values, uv.max(dim=1).values], dim=1)
    c = torch.matmul(torch.mean(points, dim=0, keepdim=True).unsqueeze(0), T.permute(0, 2, 1)).squeeze(1)
    cuv = c[:, :2] / torch.maximum(c[:, [2]], torch.tensor(0.01))
    d = torch.maximum((bboxes[:, [0, 1]] - cuv).abs_(), (bboxes[:, [2, 3]] - cuv).abs_())
    r = render_width / render_height
    w = torch.max(d[:, 0], d[:, 1] * r) * 2 * lamb
    h = torch.max(d[:, 0] / r, d[:, 1]) * 2 * lamb
    return torch.stack([cuv[:, 0] - w / 2, cuv[:, 1] - h / 2, cuv[:, 0] + w / 2, cuv[:, 1] + h / 2], dim=1)


def roi_crops(image, boxes, out_h, out_w):
    b = torch.cat([torch.zeros((len(boxes), 1)), boxes], 1)
    return torchvision.ops.roi_align(image.unsqueeze(0), b, output_size=(out_h, out_w), sampling_ratio=2)


def update_K_with_crop(K, bboxes, W, H):
    new_K = K.unsqueeze(0).repeat(len(bboxes), 1, 1)
    cw, ch = bboxes[:, 2] - bboxes[:, 0], bboxes[:, 3] - bboxes[:, 1]
    cx = K[0, 2] + (cw - 1) / 2 - (bboxes[:, 0] + bboxes[:, 2]) / 2
    cy = K[1, 2] + (ch - 1) / 2 - (bboxes[:, 1] + bboxes[:, 3]) / 2
    new_K[:, 0, 0] = W / cw * K[0, 0]
    new_K[:, 1, 1] = H / ch * K[1, 1]
    new_K[:, 0, 2] = (W - 1) / 2 + W / cw * (cx - (cw - 1) / 2)
    new_K[:, 1, 2] = (H - 1) / 2 + H / ch * (cy - (ch - 1) / 2)
    return new_K


def crop_points(vertices):
    np.random.seed(42)                                            # tracking_refiner.py:47-48
    v = np.asarray(vertices)[np.random.choice(np.arange(len(vertices)), 100)]
    return torch.from_numpy(np.pad(v, ((0, 0), (0, 1)), constant_values=1.).copy()).float()


def render(mesh, K, transform, size=IMAGE_SIZE):
    """tracking_refiner.py:31-45 through the C raster restatement: (size,size,3) u8, (size,size) fp32."""
    S = (size + 3) // 4 * 4
    k = np.asarray(K, dtype=np.float32)
    rgb, depth = oraster.render_mesh(mesh, np.asarray(transform, dtype=np.float64)[None], float(k[0, 0]), float(k[1, 1]),
                                     float(k[0, 2]), float(k[1, 2]), S, msaa=4, cull=True, ambient=5.0, znear=1e-4,
                                     zfar=9999.0)
    return rgb[0, :size, :size], depth[0, :size, :size]


def valid_mask(depth):
    return cv2.resize((depth > 0).astype(np.float32), (G, G), interpolation=cv2.INTER_CUBIC) > 0.5


def threshold_for_confidence(sim, top_quantile=0.2):
    counts, values = np.histogram(sim[sim > 0], bins=50)
    cutoff = counts.sum() * top_quantile
    cum, v = 0, values[0]
    for c, v in zip(counts[::-1], values[:-1][::-1]):
        cum += c
        if cum > cutoff:
            break
    return v


class OracleRefiner:
    def __init__(self, state_dict, mode="fp32"):
        """mode 'fp32' = the reference's arithmetic; 'contract' = the engine's bf16 rounding contract."""
        self.mode = mode
        self.vit = OracleViT(state_dict, VITB14_REG, contract=(mode == "contract"))
        self.norm = torchvision.transforms.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))

    @torch.no_grad()
    def tokens(self, image01):
        """(3,518,518) fp32 in [0,1] -> (37,37,768) x_norm_patchtokens."""
        if self.mode == "contract":
            x = self.norm(image01.to(torch.bfloat16)).float()[None]
        else:
            x = self.norm(image01)[None]
        t = self.vit.forward_features(x, len(self.vit.blocks))[0, 5:]
        return t.view(G, G, -1)

    @torch.no_grad()
    def pose_confidence(self, mesh, photo01, K, transform):
        """photo01: (3,H,W) fp32 in [0,1] -> (37,37) fp32 and the intermediate products (for stage-level checks)."""
        Kt = torch.from_numpy(np.asarray(K, dtype=np.float64)).view(3, 3).float()
        Tt = torch.from_numpy(np.asarray(transform, dtype=np.float64)).view(1, 4, 4).float()
        boxes = crop_boxes(Tt, crop_points(mesh.vertices), Kt, IMAGE_SIZE, IMAGE_SIZE)
        crop = roi_crops(photo01, boxes, IMAGE_SIZE, IMAGE_SIZE)[0]
        new_K = update_K_with_crop(Kt, boxes, IMAGE_SIZE, IMAGE_SIZE)[0]
        rgb, depth = render(mesh, new_K.numpy(), transform)
        mask = valid_mask(depth)
        fa = self.tokens(crop)
        fb = self.tokens(torch.from_numpy(rgb.astype(np.float32) / 255).permute(2, 0, 1).contiguous())
        fa = fa / torch.linalg.norm(fa, dim=-1, keepdim=True)
        fb = fb / torch.linalg.norm(fb, dim=-1, keepdim=True)
        conf = ((fa * fb).sum(-1) * torch.from_numpy(mask).float()).numpy()
        return conf, dict(boxes=boxes[0].numpy(), crop=crop.numpy(), new_K=new_K.numpy(), rgb=rgb, depth=depth, mask=mask)


def synthetic_case(i: int, seed: int = 11):
    """Seeded test scene shared by tests/golden/make_golden.py and the tests: (mesh, frame u8 (480,640,3), K, T).  The
    "photo" is the object rendered at a slightly perturbed pose over seeded noise."""
    from freepose_b200.synthetic import synthetic_mesh
    mesh = synthetic_mesh(4, subdivisions=3, scale=0.1)
    rng = np.random.default_rng(seed + 100 * i)
    H, W = 480, 640
    K = np.array([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    T = np.eye(4)
    T[:3, :3] = q
    T[:3, 3] = [rng.uniform(-0.08, 0.08), rng.uniform(-0.05, 0.05), rng.uniform(0.5, 0.7)]
    T2 = T.copy()
    T2[:3, 3] += rng.normal(scale=0.004, size=3)
    rgb, depth = oraster.render_mesh(mesh, T2[None], 800.0, 800.0, 320.0, 240.0, 640, msaa=4, ambient=3.0)
    frame = rng.integers(0, 90, size=(H, W, 3), dtype=np.uint8)
    m = depth[0, :H, :W] > 0
    frame[m] = rgb[0, :H, :W][m]
    return mesh, frame, K, T
