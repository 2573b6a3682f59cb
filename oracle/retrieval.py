"""ORACLE (test infrastructure): the reference's mesh-retrieval arithmetic on the CPU.

* ``reference_retrieve`` / ``reference_softvote`` execute the reference's own lines
  (scripts/extract_proposals_ground.py:39-41,136-160; scripts/extract_proposals_ground_video.py:148-190) with CPU bf16
  tensors -- the loops are inline script code there, so they are restated line by line, not imported.
* ``engine_*`` restate the same rounding points in numpy fp32 with the summation order the CUDA kernels fix
  (freepose_b200/csrc/retrieval.cu, rowops.cuh), so kernel outputs can be compared BIT-EXACTLY; ties of every top-k go
  to the lowest index (torch.topk leaves the order unspecified).
tests/test_oracle_golden.py pins the second against the first on the committed fixture (tests/golden/retrieval.npz).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .score import _bf16_round, _lane_reduce, _normalise_rows


# ------------------------------------------------------------------------------------------- reference lines
def reference_database(features: np.ndarray) -> torch.Tensor:
    """extract_proposals_ground.py:39-41 (device 'cuda' -> CPU)."""
    retrieval_features = torch.from_numpy(features).to(dtype=torch.bfloat16)
    return F.normalize(retrieval_features, dim=-1)


def reference_retrieve(retrieval_features: torch.Tensor, fine_features, feature: torch.Tensor, topk: int):
    """One iteration of the proposal loop, extract_proposals_ground.py:136-160.  fine_features: list of (V_i, D) fp32
    arrays indexed by mesh.  Returns (best index, best score, coarse indices I, dense s vector of the video variant)."""
    scores = (retrieval_features @ feature).float()
    scores, I = torch.topk(scores, 100)
    s = torch.zeros(retrieval_features.shape[0])
    if topk == 0:
        s[I] = scores
        return int(I[0].item()), scores[0].item(), I.numpy(), s
    fine = {}
    for i, idx in enumerate(I.numpy()):
        finegrained_features = torch.from_numpy(fine_features[idx]).to(dtype=torch.bfloat16)
        finegrained_features = F.normalize(finegrained_features, dim=-1)
        pred_scores = (finegrained_features @ feature).float()
        topk_scores, topk_idx = torch.topk(pred_scores, topk)
        fine[int(idx)] = topk_scores.numpy().mean().item()
        s[idx] = topk_scores.numpy().mean().item()
    mesh = max(fine, key=fine.get)
    return mesh, fine[mesh], I.numpy(), s


def reference_softvote(per_frame_s):
    """extract_proposals_ground_video.py:186-188: per_frame_s = list over frames of (P, M) tensors."""
    softvote_scores = torch.mean(torch.stack(per_frame_s), axis=0)
    scores, I = torch.topk(softvote_scores, 1, dim=1)
    return I[:, 0].numpy(), scores[:, 0].numpy()


# ------------------------------------------------------------------------------------------- engine order
def engine_normalize(x: np.ndarray) -> np.ndarray:
    """F.normalize(x.to(bf16), dim=-1) with the kernel's reduction order; fp32 array holding bf16 values."""
    return _normalise_rows(_bf16_round(np.asarray(x, dtype=np.float32)))


def engine_scan(db_n: np.ndarray, q_n: np.ndarray) -> np.ndarray:
    """(M, D) x (Q, D) normalised -> (Q, M) bf16-valued fp32 scores."""
    out = np.empty((q_n.shape[0], db_n.shape[0]), dtype=np.float32)
    for q in range(q_n.shape[0]):
        out[q] = _bf16_round(_lane_reduce((db_n * q_n[q][None, :]).astype(np.float32)))
    return out


def engine_topk(scores: np.ndarray, k: int):
    """Descending, NaN first, -0 == +0, ties -> lowest index.  1-D -> (idx (k,), val (k,)); 2-D row-wise."""
    s = np.asarray(scores, dtype=np.float32)
    if s.ndim == 2:
        idx = np.empty((s.shape[0], k), dtype=np.int32)
        val = np.empty((s.shape[0], k), dtype=np.float32)
        for r in range(s.shape[0]):
            idx[r], val[r] = engine_topk(s[r], k)
        return idx, val
    nan = np.isnan(s)
    finite_key = np.where(nan, 0.0, s).astype(np.float64)            # -0.0 == +0.0 under comparison
    order = np.lexsort((np.arange(s.shape[0]), -finite_key, ~nan))   # last key is primary: NaN (False) first
    return order[:k].astype(np.int32), s[order[:k]]


def numpy_order_mean(v: np.ndarray) -> np.float32:
    """np.mean of a contiguous float32 vector with n <= 128, restated (pairwise add.reduce: < 8 sequential, else eight
    strided partial sums + ordered combine + sequential remainder); the kernel follows the same order."""
    v = np.asarray(v, dtype=np.float32)
    n = v.shape[0]
    assert 1 <= n <= 128
    if n < 8:
        acc = v[0]
        for i in range(1, n):
            acc = np.float32(acc + v[i])
    else:
        r = [np.float32(v[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = np.float32(r[j] + v[i + j])
            i += 8
        acc = np.float32(np.float32(np.float32(r[0] + r[1]) + np.float32(r[2] + r[3])) +
                         np.float32(np.float32(r[4] + r[5]) + np.float32(r[6] + r[7])))
        while i < n:
            acc = np.float32(acc + v[i])
            i += 1
    return np.float32(acc / np.float32(n))


def engine_fine(fine_features, cand: np.ndarray, q_n: np.ndarray, k: int) -> np.ndarray:
    """cand (Q, C) mesh indices -> (Q, C) float32 means of the top-k per-view scores."""
    out = np.empty(cand.shape, dtype=np.float32)
    cache = {}
    for q in range(cand.shape[0]):
        for c in range(cand.shape[1]):
            m = int(cand[q, c])
            if m not in cache:
                cache[m] = engine_normalize(fine_features[m])
            p = _bf16_round(_lane_reduce((cache[m] * q_n[q][None, :]).astype(np.float32)))
            _, top = engine_topk(p, k)
            out[q, c] = numpy_order_mean(top)
    return out


def engine_retrieve(db_n: np.ndarray, fine_features, q_n: np.ndarray, topk: int, coarse_k: int = 100):
    """-> (best mesh index (Q,), best score (Q,), candidate indices (Q, C), candidate scores (Q, C))."""
    idx, val = engine_topk(engine_scan(db_n, q_n), coarse_k)
    if topk == 0:
        return idx[:, 0].copy(), val[:, 0].copy(), idx, val
    fine = engine_fine(fine_features, idx, q_n, topk)
    best = np.array([engine_topk(fine[q], 1)[0][0] for q in range(fine.shape[0])])
    rows = np.arange(fine.shape[0])
    return idx[rows, best], fine[rows, best], idx, fine


def engine_softvote_dense(per_frame, M: int):
    P = per_frame[0][0].shape[0]
    acc = np.zeros((P, M), dtype=np.float32)
    for idx, val in per_frame:
        for p in range(P):
            acc[p, idx[p]] = (acc[p, idx[p]] + val[p]).astype(np.float32)
    mean = (acc / np.float32(len(per_frame))).astype(np.float32)
    best, score = engine_topk(mean, 1)
    return best[:, 0], score[:, 0], mean


# ------------------------------------------------------------------------------------------- seeded test case
def synthetic_case(seed: int = 0, M: int = 1500, D: int = 1024, Q: int = 5, frames: int = 6):
    """Seeded stand-in for data/<retrieval>.npy + per-mesh view features + proposal features (no dataset offline).
    Mesh m has 12..40 views scattered around a mesh centre; the coarse row is their mean (as the FFA/cls tables are
    built); queries are noisy views of chosen meshes; rows 3 and 7 are exact duplicates (ties)."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((M, D)).astype(np.float32)
    centres += 0.8 * rng.standard_normal((1, D)).astype(np.float32)          # common component: similar scores
    nviews = rng.integers(12, 41, size=M)
    fine = [(centres[m] + 0.9 * rng.standard_normal((int(nviews[m]), D))).astype(np.float32) for m in range(M)]
    fine[7] = fine[3].copy()
    db = np.stack([f.mean(axis=0) for f in fine]).astype(np.float32)
    targets = rng.integers(0, M, size=Q)
    queries = np.stack([fine[t][0] + 0.7 * rng.standard_normal(D) for t in targets]).astype(np.float32)
    queries[1] = fine[3][2] + 0.5 * rng.standard_normal(D).astype(np.float32)  # lands on the duplicated pair
    video = [(queries[:3] + 0.6 * rng.standard_normal((3, D))).astype(np.float32) for _ in range(frames)]
    return {"db": db, "fine": fine, "queries": queries, "video": video, "ids": [f"mesh_{m:05d}" for m in range(M)]}
