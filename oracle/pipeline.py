"""ORACLE (test infrastructure): the whole hot path restated on the CPU, stage by stage, in the order of the
reference call stack:

  render (oracle/raster_ref.c ~ renderer.py:43-95) -> generate_proposals (oracle/crop.py ~ renderer.py:110-129)
  -> T.Normalize in bf16 (dino.py:12,16) -> ViT-L/14-reg to `layer` + norm + patch slice (oracle/vit.py ~ dino.py:16-30)
  -> score / top-k (oracle/score.py ~ pose_estimator.py:85-92) -> translation (pose_estimator.py:103-113).

Also the CPU baseline arm of bench.py (`--impl reference` and `cpu_baseline`): pyrender is not installable here,
so the timed CPU path is this restatement with PyTorch-eager ViT -- bench.py reports kind="port".
"""
from __future__ import annotations

import numpy as np
import torch
import torchvision.transforms as T

from freepose_b200.pipeline import utils as U
from freepose_b200.vit_weights import VITL14_REG

from . import crop as ocrop
from . import raster as oraster
from . import score as oscore
from .vit import OracleViT

_NORM = T.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))


def reference_normalize(images: torch.Tensor) -> torch.Tensor:
    """``self.transform(images)`` of dino.py:16 -- torchvision's own Normalize, in the tensor's dtype."""
    return _NORM(images)


class OraclePipeline:
    def __init__(self, state_dict, resolution=224, focal=None, msaa=4, mode="contract", layer=22):
        """mode: 'contract' (explicit rounding contract, fp32 math on bf16 values), 'eager' (native PyTorch bf16,
        what the reference executes) or 'fp32'."""
        self.res = resolution
        self.focal = focal if focal is not None else 600.0 * resolution / 420.0
        self.msaa = msaa
        self.mode = mode
        self.layer = layer
        if mode == "contract":
            self.vit = OracleViT(state_dict, VITL14_REG, contract=True)
        elif mode == "eager":
            self.vit = OracleViT(state_dict, VITL14_REG).to(torch.bfloat16)
        else:
            self.vit = OracleViT(state_dict, VITL14_REG).float()

    # -- stages ---------------------------------------------------------------------------------
    def render(self, mesh, poses):
        r = self.res
        return oraster.render(mesh.vertices, mesh.faces, _colors(mesh), np.asarray(poses), self.focal, self.focal,
                              r / 2, r / 2, r, msaa=self.msaa)

    def proposals(self, rgb, depth):
        r = self.res
        fb = (105, 315) if r == 420 else (r // 4, r - r // 4)
        return ocrop.generate_proposals(list(zip(rgb, depth)), r, fallback=fb)

    @torch.no_grad()
    def features(self, images_f32: torch.Tensor, batch_size=16) -> torch.Tensor:
        """(B,3,T,T) fp32 in [0,1] -> (B,P,1024) patch tokens (bf16 for the bf16 modes)."""
        outs = []
        for i in range(0, len(images_f32), batch_size):
            x = images_f32[i:i + batch_size]
            if self.mode == "fp32":
                t = self.vit.forward_features(reference_normalize(x), self.layer)
            else:
                xn = reference_normalize(x.to(torch.bfloat16))      # Normalize runs on the bf16 tensor
                t = self.vit.forward_features(xn if self.mode == "eager" else xn.float(), self.layer)
                t = t.to(torch.bfloat16)
            outs.append(t[:, 1 + self.vit.num_register_tokens:])
        return torch.cat(outs)

    def score(self, feats_t, feat_q, k=3, engine_order=False):
        if engine_order:
            s = oscore.engine_order_scores(feats_t, feat_q)
        else:
            s = oscore.reference_scores(feats_t, feat_q.reshape(1, *feats_t.shape[1:])).float().numpy()
        idx, vals = oscore.stable_topk(s, k)
        return s, idx, vals

    def translation(self, depth, K_template, bbox, K, pose, est_scale, recentre=True):
        pc = U.depthmap_to_pointcloud(depth, K_template)
        if recentre:
            m = pc.mean(axis=0); pc -= m; pc /= 0.25; pc *= est_scale; pc += m
        else:
            pc /= 0.25; pc *= est_scale
        return U.get_z_from_pointcloud(np.asarray(bbox), pc, np.asarray(K), pose)

    # -- the per-proposal hot path -----------------------------------------------------------------
    def forward(self, proposal_f32, mesh, K, bbox, est_scale, poses, k=3):
        rgb, depth = self.render(mesh, poses)
        templates, boxes, masks = self.proposals(rgb, depth)
        feats_t = self.features(torch.from_numpy(templates))
        feat_q = self.features(proposal_f32[None])
        s, idx, vals = self.score(feats_t, feat_q, k)
        r = self.res
        K_t = np.array([[self.focal, 0, r / 2], [0, self.focal, r / 2], [0, 0, 1]])
        tco = [self.translation(depth[i], K_t, bbox, K, np.asarray(poses[i]), est_scale) for i in idx]
        return {"TCO": tco, "scores": vals, "top_indices": idx, "all_scores": s, "rgb": rgb, "depth": depth,
                "templates": templates, "feats_t": feats_t, "feat_q": feat_q}


def _colors(mesh):
    if mesh.vertex_colors is None:
        return np.full((len(mesh.vertices), 3), 255, dtype=np.uint8)
    return np.asarray(mesh.vertex_colors)[:, :3].astype(np.uint8)


def synthetic_query(mesh, resolution, seed=1, msaa=4, noise=0.02):
    """Query crop of SURVEY.md section 8d: a render of the same mesh at a held-out rotation + Gaussian noise,
    cropped like a proposal.  Returns ((3,T,T) fp32 in [0,1], pose)."""
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    pose = np.eye(4)
    pose[:3, :3] = q
    pose[2, 3] = 1.1
    f = 600.0 * resolution / 420.0
    rgb, depth = oraster.render(mesh.vertices, mesh.faces, _colors(mesh), pose[None], f, f, resolution / 2,
                                resolution / 2, resolution, msaa=msaa)
    crops, _, _ = ocrop.generate_proposals(list(zip(rgb, depth)), resolution,
                                           fallback=(resolution // 4, resolution - resolution // 4))
    img = crops[0] + rng.normal(scale=noise, size=crops[0].shape).astype(np.float32)
    return torch.from_numpy(np.clip(img, 0.0, 1.0).astype(np.float32)), pose
