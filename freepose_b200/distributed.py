"""Multi-GPU plumbing: one process per GPU, hypotheses of a proposal sharded contiguously across ranks, ONE
all-gather of the per-hypothesis fp32 scores (<= a few KB: latency bound over NVLink/NVSwitch), then every rank
runs the identical deterministic top-k (ties -> lowest global index).  SURVEY.md section 8e.

The reference has no distributed code (SLURM array jobs + CSV concatenation, scripts/dino_inference.py:37-40).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split: rank r owns [r*ceil(n/W), min(n, (r+1)*ceil(n/W)))."""
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


class ScoreGather:
    """Owns the (world * per) fp32 gather buffer.  ``local_view(rank)`` is where this rank's score kernel writes
    its scores directly, so no copy kernel runs between scoring and the collective."""

    def __init__(self, n_total: int, world: int, device):
        self.n = n_total
        self.world = world
        self.per = -(-n_total // world)
        # padding slots (when world does not divide n) stay at -inf and never win the top-k
        self.buf = torch.full((world * self.per,), float("-inf"), dtype=torch.float32, device=device)

    def local_view(self, rank: int) -> torch.Tensor:
        return self.buf[rank * self.per:(rank + 1) * self.per]

    def gather(self, rank: int) -> torch.Tensor:
        """All ranks end with all scores; returns the first n_total entries (global hypothesis order)."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.buf, self.local_view(rank).clone() if self.buf.device.type == "cpu"
                                        else self.local_view(rank))
        return self.buf[:self.n]


def stable_topk_host(scores: torch.Tensor, k: int):
    """Deterministic CPU top-k (descending, ties -> lowest index): used on gloo/CPU ranks in tests; GPUs use fp_topk."""
    vals, idx = torch.sort(scores, descending=True, stable=True)
    return idx[:k], vals[:k]
