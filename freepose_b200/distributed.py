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
    its scores directly, so no copy kernel runs between scoring and the collective.

    On CUDA the collective is ``fp_allgather_scores`` of the C ABI (an NCCL communicator created by the library from
    an id that rank 0 makes and torch.distributed hands round), enqueued on the current compute stream right behind the
    score kernel.  On CPU tensors (gloo tests of the host logic) it is ``torch.distributed.all_gather_into_tensor``."""

    def __init__(self, n_total: int, world: int, device, rank: int | None = None):
        self.n = n_total
        self.world = world
        self.per = -(-n_total // world)
        self.device = torch.device(device)
        # padding slots (when world does not divide n) stay at -inf and never win the top-k
        self.buf = torch.full((world * self.per,), float("-inf"), dtype=torch.float32, device=device)
        self._comm = None
        if world > 1 and self.device.type == "cuda":
            import ctypes as C
            from . import _lib
            lib = _lib.load()
            rank = dist.get_rank() if rank is None else rank
            ident = (C.c_char * 128)()
            if rank == 0:
                _lib.check(lib.fp_comm_unique_id(ident), "fp_comm_unique_id")
            box = [bytes(ident)]
            dist.broadcast_object_list(box, src=0)
            ident = (C.c_char * 128).from_buffer_copy(box[0])
            comm = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(lib.fp_comm_create(ident, rank, world, C.byref(comm)), "fp_comm_create")
            self._comm, self._lib = comm, lib

    def local_view(self, rank: int) -> torch.Tensor:
        return self.buf[rank * self.per:(rank + 1) * self.per]

    def gather(self, rank: int) -> torch.Tensor:
        """All ranks end with all scores; returns the first n_total entries (global hypothesis order)."""
        if self.world > 1:
            if self._comm is not None:
                from . import _lib
                _lib.check(self._lib.fp_allgather_scores(self._comm, _lib.ptr(self.buf), self.per, _lib.stream_ptr()),
                           "fp_allgather_scores")
            else:
                dist.all_gather_into_tensor(self.buf, self.local_view(rank).clone())
        return self.buf[:self.n]

    def close(self):
        if self._comm is not None:
            self._lib.fp_comm_destroy(self._comm)
            self._comm = None


def stable_topk_host(scores: torch.Tensor, k: int):
    """Deterministic CPU top-k (descending, ties -> lowest index): used on gloo/CPU ranks in tests; GPUs use fp_topk."""
    vals, idx = torch.sort(scores, descending=True, stable=True)
    return idx[:k], vals[:k]
