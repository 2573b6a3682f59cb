"""ORACLE (test infrastructure): the reference's scoring arithmetic on the CPU.

* ``reference_scores`` executes the reference's own lines (pose_estimator.py:85-90,
  online_pose_estimator.py:68-79) with CPU bf16 tensors -- F.normalize, einops-style einsum, mean, topk.
* ``engine_order_scores`` restates the same rounding points in numpy fp32 with the summation order the CUDA
  kernel fixes (freepose_b200/csrc/score.cu header), so kernel output can be compared BIT-EXACTLY.
  tests/test_oracle_score.py pins the second against the first (identical up to rare 1-ulp bf16 flips caused by
  ATen's different fp32 summation order; identical top-k on the seeded cases).
* ``stable_topk``: descending, ties -> lowest index (torch.topk's tie order is unspecified).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def reference_scores(feats_t: torch.Tensor, feat_q: torch.Tensor, weights=None, normalise_query=True):
    """feats_t (B,P,D) bf16, feat_q (1,P,D) bf16 -> scores (B,) (bf16, or fp32 when weighted)."""
    q = F.normalize(feat_q, dim=-1) if normalise_query else feat_q
    s = torch.einsum("bnd,bnd->bn", F.normalize(feats_t, dim=-1), q.expand_as(feats_t))
    if weights is None:
        return s.mean(dim=-1)
    return (s * weights).sum(dim=-1) / weights.sum(dim=-1)


def _bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 -> nearest bf16 (ties to even) -> fp32, on raw bits."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(np.float32).copy()
    nan = np.isnan(x)
    if nan.any():
        out[nan] = np.nan
    return out


def _lane_reduce(prod: np.ndarray) -> np.ndarray:
    """prod (..., D) fp32 -> (...,) fp32 in kernel order (csrc/rowops.cuh): lane l owns elements c*256 + l*8 + j and
    keeps two running sums -- over its even j and over its odd j, chunks c outer -- adds them (even + odd), then the
    xor butterfly 16,8,4,2,1 across the 32 lanes."""
    D = prod.shape[-1]
    chunks = D // 256
    p = prod.reshape(prod.shape[:-1] + (chunks, 32, 4, 2))
    acc = np.zeros(prod.shape[:-1] + (32, 2), dtype=np.float32)
    for c in range(chunks):
        for w in range(4):
            acc = (acc + p[..., c, :, w, :]).astype(np.float32)
    acc = (acc[..., 0] + acc[..., 1]).astype(np.float32)
    lanes = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        acc = (acc + acc[..., lanes ^ o]).astype(np.float32)
    return acc[..., 0]


def _strided_reduce(vals: np.ndarray) -> np.ndarray:
    """vals (..., P) fp32 -> (...,): lane l sums n = l, l+32, ... ascending, then the butterfly."""
    P = vals.shape[-1]
    acc = np.zeros(vals.shape[:-1] + (32,), dtype=np.float32)
    for n0 in range(0, P, 32):
        blk = vals[..., n0:n0 + 32]
        w = blk.shape[-1]
        acc[..., :w] = (acc[..., :w] + blk).astype(np.float32)
    lanes = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        acc = (acc + acc[..., lanes ^ o]).astype(np.float32)
    return acc[..., 0]


def _normalise_rows(x: np.ndarray) -> np.ndarray:
    ss = _lane_reduce((x * x).astype(np.float32))
    nrm = _bf16_round(np.sqrt(ss, dtype=np.float32))
    nrm = np.maximum(nrm, _bf16_round(np.array([1e-12], dtype=np.float32))[0])
    return _bf16_round((x / nrm[..., None]).astype(np.float32))


def engine_order_scores(feats_t: torch.Tensor, feat_q: torch.Tensor, weights=None, normalise_query=True,
                        return_patch=False):
    t = feats_t.float().numpy()
    q = feat_q.float().numpy().reshape(t.shape[1], t.shape[2])
    qn = _normalise_rows(q) if normalise_query else q
    B = t.shape[0]
    s = np.empty(t.shape[:2], dtype=np.float32)
    for b in range(B):  # one hypothesis at a time keeps the temporary at P x D
        tn = _normalise_rows(t[b])
        s[b] = _bf16_round(_lane_reduce((tn * qn).astype(np.float32)))
    if weights is None:
        tot = _strided_reduce(s)
        scores = _bf16_round((tot / np.float32(t.shape[1])).astype(np.float32))
    else:
        w = weights.float().numpy()
        num = _strided_reduce((s * w).astype(np.float32))
        den = _strided_reduce(w)
        scores = (num / den).astype(np.float32)
    return (scores, s) if return_patch else scores


def stable_topk(scores, k: int):
    s = torch.as_tensor(np.asarray(scores, dtype=np.float32))
    vals, idx = torch.sort(s, descending=True, stable=True)
    return idx[:k].numpy().astype(np.int64), vals[:k].numpy()


def ffa_reference(feats: torch.Tensor, masks: np.ndarray):
    """Reference lines extract_retrieval_features.py:51-57 (cv2 INTER_AREA resize > 0, masked bf16 mean)."""
    import cv2
    out = []
    g = int(round(feats.shape[1] ** 0.5))
    for feat, mask in zip(feats, masks):
        m = cv2.resize(mask.astype(np.float32), (g, g), interpolation=cv2.INTER_AREA) > 0
        out.append(feat[torch.from_numpy(m.flatten())].mean(dim=0).float().numpy())
    return np.stack(out)


def ffa_engine_order(feats: torch.Tensor, masks: np.ndarray):
    """14x14 max-pool of the mask, sequential fp32 sum over selected patches, bf16-rounded mean."""
    V, P, D = feats.shape
    g = int(round(P ** 0.5))
    f = feats.float().numpy()
    out = np.empty((V, D), dtype=np.float32)
    counts = np.empty(V, dtype=np.int32)
    for v in range(V):
        cell = masks[v].reshape(g, 14, g, 14).any(axis=(1, 3)).reshape(-1)
        acc = np.zeros(D, dtype=np.float32)
        for c in np.nonzero(cell)[0]:
            acc = (acc + f[v, c]).astype(np.float32)
        cnt = int(cell.sum())
        counts[v] = cnt
        with np.errstate(invalid="ignore", divide="ignore"):
            out[v] = _bf16_round((acc / np.float32(cnt)).astype(np.float32))
    return out, counts
