"""Drop-in for reference ``src/pipeline/estimators/online_pose_estimator.py:16-95`` (coarse -> fine)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ..._lib import on_device
from ..utils import as_mesh, rescaled_extents, tco_from_extents
from .pose_estimator import DinoPoseEstimator


def _rotvec_angle(R: np.ndarray) -> np.ndarray:
    """|rotation vector| of a batch of rotation matrices = geodesic angle (what the reference obtains through
    scipy ``Rotation.from_matrix(diffs).as_rotvec()`` + norm, online_pose_estimator.py:30)."""
    tr = np.clip((np.trace(R, axis1=1, axis2=2) - 1.0) / 2.0, -1.0, 1.0)
    s = 0.5 * np.sqrt((R[:, 2, 1] - R[:, 1, 2]) ** 2 + (R[:, 0, 2] - R[:, 2, 0]) ** 2 + (R[:, 1, 0] - R[:, 0, 1]) ** 2)
    return np.arctan2(s, tr)


class DinoOnlinePoseEstimator(nn.Module):
    def __init__(self, n_coarse_poses=600, n_fine_poses=10000, cache_size=50, save_all=False,
                 cache_dir="./data/cache", resolution=420, **extractor_kwargs):
        super().__init__()
        self.coarse_estimator = DinoPoseEstimator(n_coarse_poses, cache_size, save_all, cache_dir,
                                                  resolution=resolution, **extractor_kwargs)
        self.feature_extractor = self.coarse_estimator.feature_extractor  # one set of weights in HBM, not two
        self.fine_mesh_poses = np.array(self.coarse_estimator.generate_poses(n_fine_poses))
        self.renderer = self.coarse_estimator.renderer
        self.rendering_scale = 0.25
        self.device = self.coarse_estimator.device
        self._scaled_meshes = {}   # id(mesh) -> (mesh, mesh at rendering scale): uploaded / mip-mapped once, not per frame
        self._fine_rot = None      # (n_fine, 9) rotation parts, for the neighbourhood pre-filter

    def to(self, *args, **kwargs):
        return self

    @staticmethod
    def geodesic_distance(render_poses, query_pose, degrees=True):
        diffs = render_poses[:, :3, :3] @ query_pose[:3, :3].T
        d = _rotvec_angle(diffs)
        return np.rad2deg(d) if degrees else d

    def neighbourhood(self, prev_pose, neighborhood=15):
        """Indices of the fine poses within `neighborhood` degrees of prev_pose -- np.where(geodesic_distance(...) <
        neighborhood) of online_pose_estimator.py:55-56.  The 20 000 candidates are pre-filtered by the trace of
        R_i R_prev^T (cos of the angle: 9 multiply-adds per pose) with a one-degree margin; the exact rotation-vector
        angle is evaluated only on the ~25 survivors, so the selected set is the reference's."""
        Rq = prev_pose[:3, :3]
        if self._fine_rot is None:
            self._fine_rot = np.ascontiguousarray(self.fine_mesh_poses[:, :3, :3].reshape(-1, 9))
        tr = self._fine_rot @ Rq.reshape(9)
        cand = np.nonzero(tr > 1.0 + 2.0 * np.cos(np.deg2rad(min(neighborhood + 1.0, 180.0))))[0]
        if neighborhood + 1.0 >= 180.0:
            cand = np.arange(len(self.fine_mesh_poses))
        d = self.geodesic_distance(self.fine_mesh_poses[cand], prev_pose)
        return cand[d < neighborhood]

    def forward(self, proposal, proposal_mask, template_dict, mesh, K, bbox, est_scale, prev_pose=None,
                neighborhood=15, layer=22, batch_size=128, mask_scores=False):
        if prev_pose is None:
            coarse = self.coarse_estimator.forward(proposal, template_dict, K, bbox, est_scale, layer, batch_size,
                                                   return_query_feat=True)
            query_feat = coarse["query_feat"]  # NOT normalised -- reference quirk kept (online_pose_estimator.py:41,50)
            prev_pose = coarse["TCO"][0]
        else:
            query_feat = None
        return self.forward_fine(proposal, proposal_mask, template_dict, mesh, K, bbox, est_scale, prev_pose,
                                 neighborhood, layer, mask_scores, query_feat)

    def _scaled_mesh(self, mesh):
        """The caller's mesh at rendering scale (the reference scales in place and back, online_pose_estimator.py:60,87).
        Cached per mesh object so that the per-frame video loop neither copies the mesh nor re-uploads vertices,
        faces and texture (the device copies hang off the scaled Mesh)."""
        hit = self._scaled_meshes.get(id(mesh))
        if hit is not None and hit[0] is mesh:
            return hit[1]
        scaled = as_mesh(mesh).copy().apply_scale(self.rendering_scale)
        if len(self._scaled_meshes) >= 64:
            self._scaled_meshes.pop(next(iter(self._scaled_meshes)))
        self._scaled_meshes[id(mesh)] = (mesh, scaled)
        return scaled

    @torch.inference_mode()
    @on_device
    def forward_fine(self, proposal, proposal_mask, template_dict, mesh, K, bbox, est_scale, prev_pose,
                     neighborhood=15, layer=22, mask_scores=False, query_feat=None):
        normalise_query = query_feat is None
        close = self.neighbourhood(np.asarray(prev_pose), neighborhood)
        if close.size == 0:
            raise ValueError("no fine pose within the neighbourhood of prev_pose")
        selected = self.fine_mesh_poses[close]
        m = self._scaled_mesh(mesh)
        rgb, depth = self.renderer.render_device(m, selected)
        T = self.renderer.resolution
        B, P = len(close), (T // 14) ** 2
        if query_feat is None:
            # frames > 0: the query crop rides in the same batch as the ~19 fine renders (one ViT forward instead of two;
            # the reference runs it as a separate batch of one, online_pose_estimator.py:50-52 -- same arithmetic per image)
            patches = torch.empty((B + 1) * P, ops.KPAD, dtype=torch.bfloat16, device=self.device)
            _, _, masks, _ = self.renderer.proposals_device(rgb, depth, T, to_patches=True, out=patches, want_mask=mask_scores)
            q = torch.as_tensor(proposal).to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
            ops.im2col(q[None], out=patches[B * P:])
            both = self.feature_extractor.forward_patches(patches, res=T, layer=layer)
            feats, query_feat = both[:B], both[B:]
        else:
            patches, _, masks, _ = self.renderer.proposals_device(rgb, depth, T, to_patches=True, want_mask=mask_scores)
            feats = self.feature_extractor.forward_patches(patches, res=T, layer=layer)
        weights = None
        if mask_scores:
            g = T // 14
            pm = torch.as_tensor(proposal_mask).to(self.device).bool()
            mk = torch.logical_or(masks.bool(), pm[None]).float()
            weights = F.interpolate(mk[None], size=(g, g), mode="bilinear")[0].reshape(len(close), g * g).contiguous()
        scores, top_idx, top_val, _ = ops.score_topk(feats, query_feat, k=1, weights=weights,
                                                     normalise_query=normalise_query)
        top_index = int(top_idx.item())
        K_t = np.asarray(template_dict["intrinsic"]) if template_dict is not None else \
            np.array([[self.renderer.focal, 0, T / 2], [0, self.renderer.focal, T / 2], [0, 0, 1]])
        ext = ops.depth_extents(depth, K_t, view_idx=top_idx).cpu().numpy()
        dx, dy = rescaled_extents(ext[0], est_scale, recentre=False)
        TCO = tco_from_extents(bbox, dx, dy, K, selected[top_index])
        return {"TCO": [TCO], "scores": [top_val[0].float().cpu().numpy()], "proposal": proposal, "K": K,
                "bbox": bbox, "all_scores": scores, "selected_poses": selected}

    # ---------------------------------------------------------------- B200-native: all proposals of a frame at once
    @torch.inference_mode()
    @on_device
    def forward_batch(self, items, neighborhood=15, layer=22, batch_size=128, mask_scores=False):
        """``forward`` for all proposals of a frame in ONE ViT pass.  ``items``: a list of dicts with the keyword
        arguments of ``forward`` (proposal, proposal_mask, template_dict, mesh, K, bbox, est_scale, prev_pose).  Returns
        the list of ``forward`` results, identical to calling ``forward`` per proposal (the arithmetic per image does not
        depend on the batch it rides in).

        Why: the fine stage of one proposal is ~19 renders + 1 query crop = 5 220 token rows -- 84 GEMM tiles on 74 CTA
        pairs (half the second wave idle) and ~200 launches of ~10 us kernels; the reference's per-proposal loop
        (scripts/dino_inference_video.py:134-156) leaves a B200 launch- and wave-bound.  With the 8 proposals of a frame in
        one batch the same kernels run at 41 760 rows."""
        T = self.renderer.resolution
        P = (T // 14) ** 2
        plans = []
        for it in items:
            prev_pose, query_feat = it.get("prev_pose"), None
            if prev_pose is None:          # first frame: coarse stage per proposal (template features are cached per mesh)
                coarse = self.coarse_estimator.forward(it["proposal"], it["template_dict"], it["K"], it["bbox"],
                                                       it["est_scale"], layer, batch_size, return_query_feat=True)
                query_feat, prev_pose = coarse["query_feat"], coarse["TCO"][0]   # raw query: reference quirk kept
            close = self.neighbourhood(np.asarray(prev_pose), neighborhood)
            if close.size == 0:
                raise ValueError("no fine pose within the neighbourhood of prev_pose")
            plans.append({"it": it, "close": close, "selected": self.fine_mesh_poses[close], "query_feat": query_feat})
        n_img = sum(len(pl["close"]) + (pl["query_feat"] is None) for pl in plans)
        patches = torch.empty(n_img * P, ops.KPAD, dtype=torch.bfloat16, device=self.device)
        row = 0
        for pl in plans:
            B = len(pl["close"])
            rgb, depth = self.renderer.render_device(self._scaled_mesh(pl["it"]["mesh"]), pl["selected"])
            _, _, masks, _ = self.renderer.proposals_device(rgb, depth, T, to_patches=True, out=patches[row * P:],
                                                            want_mask=mask_scores)
            pl.update(depth=depth, masks=masks, lo=row, hi=row + B)
            row += B
            if pl["query_feat"] is None:
                q = torch.as_tensor(pl["it"]["proposal"]).to(self.device, dtype=torch.float32, non_blocking=True)
                ops.im2col(q.contiguous()[None], out=patches[row * P:(row + 1) * P])
                pl["q_row"] = row
                row += 1
        feats_all = self.feature_extractor.forward_patches(patches, res=T, layer=layer)
        g = T // 14
        tops = []
        for pl in plans:
            feats = feats_all[pl["lo"]:pl["hi"]]
            normalise = pl["query_feat"] is None
            qf = feats_all[pl["q_row"]:pl["q_row"] + 1] if normalise else pl["query_feat"]
            weights = None
            if mask_scores:
                pm = torch.as_tensor(pl["it"]["proposal_mask"]).to(self.device).bool()
                mk = torch.logical_or(pl["masks"].bool(), pm[None]).float()
                weights = F.interpolate(mk[None], size=(g, g), mode="bilinear")[0].reshape(len(pl["close"]), g * g).contiguous()
            scores, top_idx, top_val, _ = ops.score_topk(feats, qf, k=1, weights=weights, normalise_query=normalise)
            td = pl["it"].get("template_dict")
            K_t = np.asarray(td["intrinsic"]) if td is not None else \
                np.array([[self.renderer.focal, 0, T / 2], [0, self.renderer.focal, T / 2], [0, 0, 1]])
            ext = ops.depth_extents(pl["depth"], K_t, view_idx=top_idx)
            pl["scores"] = scores
            tops.append(torch.cat((top_idx.double(), top_val.double(), ext.reshape(-1))))
        host = torch.stack(tops).cpu().numpy()          # ONE device -> host read for the whole frame
        outs = []
        for pl, h in zip(plans, host):
            it = pl["it"]
            top_index = int(h[0])
            dx, dy = rescaled_extents(h[2:10], it["est_scale"], recentre=False)
            TCO = tco_from_extents(it["bbox"], dx, dy, it["K"], pl["selected"][top_index])
            outs.append({"TCO": [TCO], "scores": [np.float32(h[1])], "proposal": it["proposal"], "K": it["K"],
                         "bbox": it["bbox"], "all_scores": pl["scores"], "selected_poses": pl["selected"]})
        return outs

