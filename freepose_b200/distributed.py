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

    # -- the two calls the estimators make (overridden by PeerScoreGather) ----------------------------------------
    def score_into(self, rank: int, feats_t: torch.Tensor, feat_q: torch.Tensor) -> None:
        """Scores of this rank's hypotheses, written straight into its slice of the gather buffer."""
        from . import ops
        ops.score_topk(feats_t, feat_q, k=0, scores_out=self.local_view(rank))

    def gather_topk(self, rank: int, k: int):
        """-> (all n scores, top-k indices, top-k values), identical on every rank."""
        from . import ops
        scores = self.gather(rank)
        idx, val = ops.topk(scores, k)
        return scores, idx, val


class _DeviceMemory:
    """Raw device memory as a CUDA array (for torch.as_tensor)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class PeerScoreGather(ScoreGather):
    """The same exchange over PEER MEMORY instead of a collective call (one node, NVLink / NVSwitch): every rank's
    exchange buffer is mapped into every process through CUDA IPC; ``fp_score_publish`` -- the score kernel -- stores each
    score into slot (rank, b) of every rank's buffer as it is produced and its last CTA raises this rank's flag in every
    buffer; ``fp_topk_after_exchange`` waits on the device for the world flags of the own buffer and runs the top-k.
    Two launches per proposal on the exchange path (score + top-k), no NCCL call, nothing on the host."""

    def __init__(self, n_total: int, world: int, device, rank: int | None = None):
        import ctypes as C
        from . import _lib
        self.n, self.world = n_total, world
        self.per = -(-n_total // world)
        self.device = torch.device(device)
        self.rank = dist.get_rank() if rank is None else rank
        self._lib = lib = _lib.load()
        self._comm = None
        self.epoch = 0
        nbytes = lib.fp_exchange_bytes(world, self.per)
        own, handle = C.c_void_p(), (C.c_char * 64)()
        with torch.cuda.device(self.device):
            _lib.check(lib.fp_p2p_alloc(nbytes, C.byref(own), handle), "fp_p2p_alloc")
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle))
            self._peer_ptrs = []
            for r in range(world):
                if r == self.rank:
                    self._peer_ptrs.append(own.value)
                else:
                    p = C.c_void_p()
                    _lib.check(lib.fp_p2p_open((C.c_char * 64).from_buffer_copy(handles[r]), C.byref(p)), "fp_p2p_open")
                    self._peer_ptrs.append(p.value)
        self._own = own.value
        self._peers_dev = torch.tensor(self._peer_ptrs, dtype=torch.int64, device=self.device)
        self._half = (nbytes - 256) // 8                       # floats per parity half
        self._view = torch.as_tensor(_DeviceMemory(self._own, 2 * self._half), device=self.device)
        self._ws = {}
        dist.barrier()                                         # every buffer is mapped (and zeroed) before the first publish

    def _publish_ws(self, B, P, D):
        key = (B, P, D)
        if key not in self._ws:
            n = self._lib.fp_score_workspace_bytes(B, P, D) + 512
            self._ws[key] = torch.zeros(n, dtype=torch.uint8, device=self.device)
        return self._ws[key]

    def score_into(self, rank, feats_t, feat_q):
        from . import _lib
        B, P, D = feats_t.shape
        feat_q = feat_q.reshape(P, D)
        ws = self._publish_ws(B, P, D)
        self.epoch += 1
        _lib.check(self._lib.fp_score_publish(_lib.ptr(feats_t), _lib.ptr(feat_q), None, B, P, D, 1, _lib.ptr(self._peers_dev),
                                              self._own, self.rank, self.world, self.per, self.epoch, _lib.ptr(ws),
                                              ws.numel(), _lib.stream_ptr()), "fp_score_publish")

    def gather_topk(self, rank, k):
        from . import _lib
        idx = torch.empty(k, dtype=torch.int32, device=self.device)
        val = torch.empty(k, dtype=torch.float32, device=self.device)
        ws = torch.empty(max(self.n, 256), dtype=torch.uint8, device=self.device)
        _lib.check(self._lib.fp_topk_after_exchange(self._own, self.world, self.per, self.n, self.epoch, k, _lib.ptr(idx),
                                                    _lib.ptr(val), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                   "fp_topk_after_exchange")
        half = self._view[(self.epoch & 1) * self._half:]
        return half[:self.n].clone(), idx, val                 # a copy: the half is rewritten two exchanges later

    def local_view(self, rank):
        raise RuntimeError("PeerScoreGather has no host-visible local slice: use score_into / gather_topk")

    def gather(self, rank):
        raise RuntimeError("PeerScoreGather exchanges inside score_into / gather_topk")

    def close(self):
        if getattr(self, "_own", None):
            torch.cuda.synchronize(self.device)
            dist.barrier()                                     # nobody still stores into a buffer that is about to go
            for r, p in enumerate(self._peer_ptrs):
                if r != self.rank:
                    self._lib.fp_p2p_close(p)
            self._lib.fp_p2p_free(self._own)
            self._own = None


def stable_topk_host(scores: torch.Tensor, k: int):
    """Deterministic CPU top-k (descending, ties -> lowest index): used on gloo/CPU ranks in tests; GPUs use fp_topk."""
    vals, idx = torch.sort(scores, descending=True, stable=True)
    return idx[:k], vals[:k]
