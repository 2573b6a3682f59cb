"""Device-resident mesh-retrieval database.

B200 replacement of the inline retrieval code of the reference's proposal scripts:

* ``scripts/extract_proposals_ground.py:39-41``  -- ``np.load`` of the ``(46037, 1024)`` feature table, bf16,
  ``F.normalize``                                                 -> :class:`RetrievalDatabase` constructor
* ``:136-141`` coarse ``(db @ feature).float()`` + ``topk(100)``  -> :meth:`RetrievalDatabase.coarse`
* ``:148-160`` per-view fine re-rank over the 100 candidates      -> :meth:`RetrievalDatabase.fine`
* ``scripts/extract_proposals_ground_video.py:148-190`` the same per frame + soft vote over frames
                                                                  -> :class:`SoftVote`

Everything stays on the device: the coarse table (94 MB) is scanned once per call for all proposals of an image, and
the per-view features of the meshes (the reference re-reads 100 ``.npy`` files and copies them to the GPU for every
proposal) live in one bf16 pool -- 46 037 meshes x 600 views x 1024 x 2 B = 56.6 GB fits a B200's 180 GB.
Tie order of every top-k is defined (lowest index first; ``torch.topk`` leaves it unspecified).
"""
from __future__ import annotations

from pathlib import Path
from typing import Callable, Sequence

import numpy as np
import torch

from ... import ops
from ..._lib import on_device

COARSE_K = 100  # extract_proposals_ground.py:140


class RetrievalDatabase:
    def __init__(self, features, filelist: Sequence[str], fine_features=None,
                 fine_loader: Callable[[int], np.ndarray] | None = None, device="cuda",
                 pool_views: int | None = None):
        """features: (M, D) float array (``data/<retrieval>.npy``); filelist: M mesh ids (``data/mesh_cache.txt``).
        fine_features: optional list of M per-mesh ``(views_i, D)`` arrays, uploaded and normalised up front.
        fine_loader: optional ``mesh_index -> (views, D)`` array, used on demand (candidates not yet resident);
        pool_views: capacity of the on-demand pool in view rows (default 100 candidates x 32 proposals x 600 views)."""
        self.device = torch.device(device)
        feats = torch.as_tensor(np.asarray(features) if not torch.is_tensor(features) else features)
        assert feats.dim() == 2 and len(filelist) == feats.shape[0], "one mesh id per database row"
        self.filelist = list(filelist)
        self.M, self.D = feats.shape
        assert self.M < 2 ** 24, "mesh indices travel through an fp32 staging tensor"
        src = feats.to(self.device)
        src = src if src.dtype in (torch.float32, torch.bfloat16) else src.float()
        self.features = ops.normalize_rows(src.contiguous())           # (M, D) bf16, F.normalize(x.to(bf16))
        self._loader = fine_loader
        self._start = torch.zeros(self.M, dtype=torch.int64, device=self.device)
        self._count = torch.zeros(self.M, dtype=torch.int32, device=self.device)
        self._count_host = np.zeros(self.M, dtype=np.int64)            # 0 = not resident
        self._pool = None
        self._used = 0
        if fine_features is not None:
            assert len(fine_features) == self.M
            counts = np.array([len(f) for f in fine_features], dtype=np.int64)
            self._alloc_pool(int(counts.sum()))
            for m, f in enumerate(fine_features):
                self._append(m, f)
        elif fine_loader is not None:
            self._alloc_pool(pool_views or COARSE_K * 32 * 600)

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_files(cls, retrieval: str, filelist: str = "mesh_cache.txt", data_root="data", preload_fine=False,
                   device="cuda", **kw):
        """The reference's file layout: ``data/<retrieval>.npy``, ``data/<filelist>``, ``data/datasets/<retrieval>/<id>.npy``."""
        root = Path(data_root)
        ids = (root / filelist).read_text().splitlines()
        feats = np.load(root / f"{retrieval}.npy")
        fine_dir = root / "datasets" / retrieval

        def loader(m: int) -> np.ndarray:
            return np.load(fine_dir / f"{ids[m]}.npy")

        fine = [loader(m) for m in range(len(ids))] if preload_fine else None
        return cls(feats, ids, fine_features=fine, fine_loader=None if preload_fine else loader, device=device, **kw)

    def _alloc_pool(self, rows: int):
        self._pool = torch.empty(max(rows, 1), self.D, dtype=torch.bfloat16, device=self.device)
        self._used = 0

    def _append(self, m: int, views) -> None:
        v = torch.as_tensor(np.asarray(views) if not torch.is_tensor(views) else views)
        assert v.dim() == 2 and v.shape[1] == self.D, f"mesh {self.filelist[m]}: per-view features must be (views, {self.D})"
        n = v.shape[0]
        if self._used + n > self._pool.shape[0]:
            raise RuntimeError(f"retrieval fine-feature pool is full ({self._pool.shape[0]} view rows); "
                               "construct the database with a larger pool_views")
        src = v.to(self.device)
        src = src if src.dtype in (torch.float32, torch.bfloat16) else src.float()
        ops.normalize_rows(src.contiguous(), out=self._pool[self._used:self._used + n])
        self._start[m] = self._used
        self._count[m] = n
        self._count_host[m] = n
        self._used += n

    def _ensure_resident(self, mesh_indices: np.ndarray) -> None:
        need = [int(m) for m in np.unique(mesh_indices) if m >= 0]
        missing = [m for m in need if self._count_host[m] == 0]
        if not missing:
            return
        if self._loader is None:
            raise RuntimeError("fine retrieval needs per-view features: pass fine_features or fine_loader")
        loaded = {m: self._loader(m) for m in missing}
        if self._used + sum(len(v) for v in loaded.values()) > self._pool.shape[0]:
            # on-demand pool exhausted: drop everything resident and rebuild it from this request
            self._count.zero_()
            self._count_host[:] = 0
            self._used = 0
            for m in need:
                if m not in loaded:
                    loaded[m] = self._loader(m)
        for m, views in loaded.items():
            self._append(m, views)

    # ------------------------------------------------------------------ queries
    @on_device
    def normalize(self, feats: torch.Tensor) -> torch.Tensor:
        """``F.normalize(feats, dim=-1)`` on a bf16 (or fp32 -> bf16) feature matrix (``extract_proposals_ground.py:121,134``)."""
        f = feats.to(self.device)
        if f.dim() == 1:
            f = f[None]
        f = f if f.dtype in (torch.float32, torch.bfloat16) else f.float()
        return ops.normalize_rows(f.contiguous())

    @on_device
    def scores(self, features: torch.Tensor) -> torch.Tensor:
        """``(retrieval_features @ feature).float()`` for every row of ``features`` (Q, D) -> (Q, M) fp32."""
        return ops.retrieval_scan(self.features, features.contiguous())

    @on_device
    def coarse(self, features: torch.Tensor, k: int = COARSE_K):
        """-> (scores (Q,k) fp32 descending, mesh indices (Q,k) int32)."""
        if k > self.M:
            raise RuntimeError(f"selected index k out of range (k={k}, database rows={self.M})")  # as torch.topk
        idx, val = ops.topk_rows(self.scores(features), k)
        return val, idx

    @on_device
    def fine(self, features: torch.Tensor, cand: torch.Tensor, topk: int) -> torch.Tensor:
        """Per candidate: ``torch.topk(views @ feature, topk).values.mean()`` -> (Q, C) fp32."""
        cand_host = cand.cpu().numpy()
        self._ensure_resident(cand_host)
        counts = self._count_host[cand_host[cand_host >= 0]]
        if counts.size and counts.min() < topk:
            raise RuntimeError(f"selected index k out of range (k={topk}, a candidate mesh has {counts.min()} views)")
        return ops.retrieval_fine(self._pool, self._start, self._count, int(self._count_host.max()), cand.contiguous(),
                                  features.contiguous(), topk)

    @on_device
    def retrieve(self, features: torch.Tensor, topk: int = 0, coarse_k: int = COARSE_K, return_sparse: bool = False):
        """The per-proposal loop of ``extract_proposals_ground.py:136-160`` for all proposals at once.

        features: (Q, D) normalised bf16.  Returns (mesh ids [Q], scores [Q] python floats) and, with
        ``return_sparse``, the candidate indices (Q, C) int32 + their scores (Q, C) fp32 on the device -- the
        non-zero entries of the reference's dense ``s`` vector (video soft vote)."""
        val, idx = self.coarse(features, coarse_k)
        if topk == 0:
            best = torch.zeros(idx.shape[0], dtype=torch.int64, device=self.device)
            cand_scores = val
        else:
            cand_scores = self.fine(features, idx, topk)
            best = ops.topk_rows(cand_scores, 1)[0][:, 0].long()     # max(scores, key=scores.get): first maximum
        rows = torch.arange(idx.shape[0], device=self.device)
        packed = torch.stack([idx[rows, best].float(), cand_scores[rows, best]], dim=1).cpu().numpy()  # one D2H
        meshes = [self.filelist[int(i)] for i in packed[:, 0]]
        scores = [float(np.float32(s)) for s in packed[:, 1]]
        if return_sparse:
            return meshes, scores, idx, cand_scores
        return meshes, scores


class SoftVote:
    """Video soft vote (``extract_proposals_ground_video.py:153-190``): per frame a dense ``s`` vector per tracked
    object (zero except at the 100 candidates), mean over frames, ``topk(1)``.  Accumulated on the device in frame
    order (fp32, ``acc = acc + s``)."""

    def __init__(self, database: RetrievalDatabase, n_objects: int):
        self.db = database
        self.acc = torch.zeros(n_objects, database.M, dtype=torch.float32, device=database.device)
        self.frames = 0

    def add_frame(self, cand_idx: torch.Tensor, cand_scores: torch.Tensor) -> None:
        assert cand_idx.shape[0] == self.acc.shape[0], "one row per tracked object in every frame"
        ops.softvote_add(self.acc, cand_idx.contiguous(), cand_scores.contiguous())
        self.frames += 1

    def result(self):
        """-> (mesh ids [P], scores [P])."""
        mean = ops.softvote_mean(self.acc, self.frames)
        idx, val = ops.topk_rows(mean, 1)
        packed = torch.stack([idx[:, 0].float(), val[:, 0]], dim=1).cpu().numpy()
        return [self.db.filelist[int(i)] for i in packed[:, 0]], [float(np.float32(s)) for s in packed[:, 1]]
