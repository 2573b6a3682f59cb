#!/usr/bin/env python
"""Benchmark of the render-and-compare hot path (BASELINE.json metric: pose hypotheses / s).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 engine
    python bench.py --impl reference --steps K --warmup W    # the reference path on the host cores

Workload (BASELINE.json configs[1]): one proposal x one retrieved mesh x 520 pose hypotheses at 224^2.  One
step = rasterise 520 views -> mask/bbox/CropResizePad -> DINOv2 ViT-L/14-reg to layer 22 on the 520 renders and
on the query crop -> per-patch cosine score -> top-3 (+ depth extents for the translation).  With N > 1 GPUs every
rank runs one such proposal per step (weak scaling) and the per-hypothesis scores are exchanged with ONE
all-gather.  Prints one JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GFLOP_PER_HYP_224 = 150.949  # 2*588*D*g^2 + 22*(24*D^2*N + 4*N^2*D), N = 261
METRIC = "pose hypotheses/sec (raster+ViT-L22+score) @224^2"


def vit_gflop(res: int, layers: int) -> float:
    g = res // 14
    n = g * g + 5
    d = 1024
    return (2 * 588 * d * g * g + layers * (24 * d * d * n + 4 * n * n * d)) / 1e9


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"bf16_burst": j["bf16_tflops"], "bf16_sustained": j["bf16_tflops_sustained"], "hbm": j["hbm_gbs"],
                "source": "MEASURED_PEAKS.json (measured)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "B200_PROFILING.md fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- reference arm
class CpuReference:
    """The CPU restatement of the reference path (oracle/pipeline.py), set up once and timed per sample.

    The reference keeps its model in bf16; PyTorch-eager bf16 on x86 cores without AMX is several times SLOWER than
    fp32, with AMX / AVX512-BF16 it is faster: one sample of each decides, the baseline `value` is the faster one (the
    more favourable number for the reference) and the other rate is reported beside it.

    Accounting: the query crop's forward is timed SEPARATELY from the hypotheses and charged at n/520 of its cost, the
    share it has in the 520-hypothesis workload the B200 arm runs (a sample of n hypotheses must not pay a whole query
    forward: with n = 8 that would under-report the CPU by 1/9)."""

    def __init__(self, n_hyp: int, res: int, layers: int, total_hyp: int = 520):
        from freepose_b200.pipeline.utils import generate_poses
        from freepose_b200.synthetic import synthetic_mesh
        from freepose_b200.vit_weights import synthetic_state_dict
        from oracle.pipeline import OraclePipeline, synthetic_query
        torch.set_num_threads(os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        self.n = n_hyp
        self.total = total_hyp
        self.sd = synthetic_state_dict(seed=0, depth=layers)
        self.mesh = synthetic_mesh(0, subdivisions=5)
        self.res, self.layers = res, layers
        self.query, _ = synthetic_query(self.mesh, res, seed=1)
        self.poses = generate_poses(total_hyp)[:n_hyp]
        self.K = np.array([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
        self.bbox = np.array([200.0, 150.0, 330.0, 290.0])
        self._pipes = {}
        self._mk = OraclePipeline

    def pipe(self, mode):
        if mode not in self._pipes:
            self._pipes[mode] = self._mk(self.sd, self.res, mode=mode, layer=self.layers)
        return self._pipes[mode]

    def run(self, mode: str = "fp32"):
        """-> (hyp/s with the query amortised over the full workload, seconds of this sample, stage seconds)."""
        pipe = self.pipe(mode)
        t0 = time.perf_counter()
        rgb, depth = pipe.render(self.mesh, self.poses)
        templates, _, _ = pipe.proposals(rgb, depth)
        feats_t = pipe.features(torch.from_numpy(templates))
        t1 = time.perf_counter()
        feat_q = pipe.features(self.query[None])
        t2 = time.perf_counter()
        _, idx, _ = pipe.score(feats_t, feat_q, min(3, self.n))
        K_t = np.array([[pipe.focal, 0, self.res / 2], [0, pipe.focal, self.res / 2], [0, 0, 1]])
        for i in idx:
            pipe.translation(depth[i], K_t, self.bbox, self.K, np.asarray(self.poses[i]), 0.3)
        t3 = time.perf_counter()
        hyp_s, query_s = (t1 - t0) + (t3 - t2), t2 - t1
        charged = hyp_s + query_s * self.n / self.total
        return self.n / charged, t3 - t0, {"hypotheses_s": hyp_s, "query_s": query_s}

    def sample_text(self, dt, stages, other_rate=None, mode="fp32"):
        name = {"fp32": "fp32", "eager": "bf16 (the reference's dtype)"}
        s = (f"{self.n} hypotheses ({stages['hypotheses_s']:.2f} s) + 1 query forward ({stages['query_s']:.2f} s, charged "
             f"at {self.n}/{self.total} as in the {self.total}-hypothesis workload) per sample: oracle/pipeline.py = "
             f"C raster restatement + CropResizePad + PyTorch-eager {name[mode]} ViT-L/14-reg to layer {self.layers} + "
             "reference scoring lines, all host threads (the faster of fp32 / bf16 on these cores is the one timed); the "
             "true pyrender/EGL renderer is not installable offline")
        if other_rate is not None:
            s += f"; the same path in {name['eager' if mode == 'fp32' else 'fp32']} runs at {other_rate:.3f} hyp/s on these cores"
        return s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(args.ref_hyp, args.res, args.layer, args.hyp)
    # the reference keeps its model in bf16 (eager); on cores without fast bf16 kernels fp32 is the quicker way to run the
    # same path -- one sample of each decides, the faster one is timed (the more favourable number for the reference)
    mode = "fp32" if ref.run("fp32")[0] >= ref.run("eager")[0] else "eager"
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt, st = ref.run(mode)
        if i >= args.warmup:
            rates.append((r, dt, st))
    value = statistics.mean(r for r, _, _ in rates)
    dt_mean = statistics.mean(dt for _, dt, _ in rates)
    stages = {k: statistics.mean(st[k] for _, _, st in rates) for k in rates[0][2]}
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "hyp/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_mean, "higher_is_better": True,
           "scaling": args.scaling, "vs_baseline": None, "dtype": "fp32" if mode == "fp32" else "bf16", "data": "synthetic",
           "config": workload_config(args),
           "cpu_baseline": {"value": value, "unit": "hyp/s", "cores": ref.cores, "kind": "port",
                            "sample": ref.sample_text(dt_mean, stages, mode=mode)},
           "e2e": {"value": value, "unit": "hyp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def workload_config(args):
    per_gpu = "1 proposal per GPU per step" if args.scaling == "weak" else \
        "ONE proposal per step, its hypotheses split contiguously over the GPUs"
    return {"workload": f"dino_inference (BASELINE configs[1]): 1 proposal x 1 mesh x {args.hyp} pose hypotheses, "
                        f"raster + DINOv2 ViT-L/14-reg layer {args.layer} + per-patch cosine score/top-3",
            "hypotheses": args.hyp, "crop": args.res, "layer": args.layer, "mesh_faces": 20480, "msaa": 4,
            "sharding": per_gpu,
            "l2": "per-step working set ~2.9 GB (activations) >> 126 MB L2; no explicit flush needed"}


# ----------------------------------------------------------------------------------------------- parity (same run)
def parity_block(est, mesh, args, n: int):
    """SURVEY.md section 8d, last row: the B200 engine against the CPU contract oracle on a small sample of this very
    workload (the first `n` of the hypotheses + the query, all `layer` blocks), in the same process as the timing.
    The oracle is the CHECKER here, nothing timed runs through it."""
    from oracle.pipeline import OraclePipeline, synthetic_query
    from freepose_b200.vit_weights import synthetic_state_dict
    sd = synthetic_state_dict(seed=0, depth=args.layer)
    query, _ = synthetic_query(mesh, args.res, seed=1)
    poses = est.mesh_poses[:n]
    K = np.array([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
    bbox = np.array([200.0, 150.0, 330.0, 290.0])
    want = OraclePipeline(sd, args.res, mode="contract", layer=args.layer).forward(query, mesh, K, bbox, 0.3, poses,
                                                                                   k=min(3, n))
    got = est.forward_mesh(query, mesh, K, bbox, 0.3, layer=args.layer, poses=poses, k=min(3, n))
    rgb, depth = est.renderer.render_device(mesh, poses)
    feats, _, _ = est.render_features(mesh, poses, layer=args.layer)
    a, b = feats.double().cpu(), want["feats_t"].double()
    s_g, s_o = got["all_scores"].cpu().numpy().astype(np.float64), np.asarray(want["all_scores"], dtype=np.float64)
    ulp = 2.0 ** (np.floor(np.log2(np.abs(s_o))) - 7)
    rgb_h = rgb.cpu().numpy()
    return {"sample": f"first {n} of the {args.hyp} hypotheses + the query, {args.layer} blocks, vs oracle/pipeline.py "
                      "(contract mode) on the host",
            "rgb_equal": bool(np.array_equal(rgb_h, want["rgb"])),
            "rgb_max_abs_diff": int(np.abs(rgb_h.astype(np.int16) - want["rgb"].astype(np.int16)).max()),
            "depth_equal": bool(np.array_equal(depth.cpu().numpy(), want["depth"])),
            "token_rel_l2": float((a - b).norm() / b.norm()),
            "token_rel_inf": float((a - b).abs().max() / b.abs().max()),
            "token_max_abs": float((a - b).abs().max()),
            "scores_max_bf16_ulp": float((np.abs(s_g - s_o) / ulp).max()),
            "argmax_equal": bool(int(got["top_indices"][0]) == int(want["top_indices"][0])),
            "top3_equal": bool([int(i) for i in got["top_indices"]] == [int(i) for i in want["top_indices"]]),
            "tco_max_abs_diff": float(max(np.abs(x - y).max() for x, y in zip(got["TCO"], want["TCO"]))
                                      if [int(i) for i in got["top_indices"]] == [int(i) for i in want["top_indices"]]
                                      else float("nan"))}


# ----------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    from freepose_b200 import _lib
    from freepose_b200.distributed import PeerScoreGather, ScoreGather, init_from_env
    from freepose_b200.pipeline.estimators.pose_estimator import DinoPoseEstimator
    from freepose_b200.synthetic import synthetic_mesh
    from freepose_b200.vit_weights import synthetic_state_dict
    import torch.distributed as dist

    rank, local, world = init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    dev = torch.device("cuda", local)
    lib = _lib.load()
    torch.manual_seed(0)
    strong = args.scaling == "strong"

    sd = synthetic_state_dict(seed=0, depth=args.layer)
    mesh = synthetic_mesh(0, subdivisions=5)
    est = DinoPoseEstimator(n_poses=args.hyp, cache_size=0, cache_dir=f"/tmp/fp_bench_cache_{rank}", weights=sd,
                            resolution=args.res, chunk=args.chunk)
    # query: render of the mesh at a held-out rotation + noise, cropped like a proposal.  Weak scaling: every rank has
    # its own proposal (seed 1 + rank); strong scaling: all ranks work on the SAME proposal (seed 1).
    qseed = 1 if strong else 1 + rank
    rng = np.random.default_rng(qseed)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    qpose = np.eye(4); qpose[:3, :3] = q; qpose[2, 3] = 1.1
    rgb, depth = est.renderer.render_device(mesh, [qpose])
    crop, _, _, _ = est.renderer.proposals_device(rgb, depth, args.res, to_patches=False)
    noise = torch.randn(crop.shape, generator=torch.Generator().manual_seed(qseed - 1)).to(dev) * 0.02
    query_dev = (crop[0] + noise[0]).clamp(0, 1).contiguous()
    query_host = query_dev.cpu().pin_memory()
    K = np.array([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
    bbox = np.array([200.0, 150.0, 330.0, 290.0])
    # strong: ONE gather buffer of args.hyp scores, each rank fills its shard.  weak: the ranks' proposals are
    # independent -- nothing is exchanged per step; one gather of every rank's scores closes the timed region.
    peer = strong and world > 1 and args.exchange == "p2p"
    sg = (PeerScoreGather if peer else ScoreGather)(args.hyp if strong else args.hyp * world, world, dev, rank=rank)
    shard = (rank, world, sg) if (strong and world > 1) else None

    def step_device():
        """One proposal, device resident: raster -> crop -> ViT -> score -> [all-gather] -> top-k -> extents."""
        return est.forward_mesh_device(query_dev, mesh, layer=args.layer, k=3, shard=shard)

    def step_e2e():
        """The public call (DinoPoseEstimator.forward_mesh) with HOST buffers: pinned query crop in, TCO / scores out."""
        return est.forward_mesh(query_host, mesh, K, bbox, 0.3, layer=args.layer, k=3, shard=shard)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up, then the device-resident timed region (value) with live per-kernel events
    for _ in range(max(args.warmup, 3)):
        last = step_device()
    if world > 1 and not peer:
        sg.gather(rank)                                    # NCCL sets its transports up lazily at the first collective
    barrier()
    lib.fp_profile_reset()
    lib.fp_profile_enable(1)
    launches0 = lib.fp_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        last = step_device()
    if not strong and world > 1:
        sg.local_view(rank).copy_(last[0])
        sg.gather(rank)                                    # the one exchange of the weak mode: after the K steps
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    lib.fp_profile_enable(0)
    launches = lib.fp_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    units = args.hyp * args.steps * (1 if strong else world)
    value = units / (ms_total / 1e3)

    kinds = {}
    for k in range(lib.fp_profile_num_kinds()):
        ms, work, n = C.c_double(), C.c_double(), C.c_longlong()
        _lib.check(lib.fp_profile_collect(k, C.byref(ms), C.byref(work), C.byref(n)), "fp_profile_collect")
        if n.value:
            kinds[lib.fp_profile_kind_name(k).decode()] = (ms.value, work.value, n.value)
    lib.fp_profile_reset()

    # ---- strong scaling: every rank must hold the identical result, and it must equal the single-GPU result
    strong_check = None
    if strong:
        sc_s, idx_s, val_s, _ = step_device()
        sc_1, idx_1, val_1, _ = est.forward_mesh_device(query_dev, mesh, layer=args.layer, k=3, shard=None)
        same_as_n1 = bool(torch.equal(sc_s, sc_1) and torch.equal(idx_s, idx_1) and torch.equal(val_s, val_1))
        mine = [int(i) for i in idx_s.cpu()] + [float(v) for v in val_s.cpu()]
        everyone = [None] * world
        if world > 1:
            dist.all_gather_object(everyone, (mine, same_as_n1))
        else:
            everyone = [(mine, same_as_n1)]
        strong_check = {"identical_on_all_ranks": all(e[0] == everyone[0][0] for e in everyone),
                        "equal_to_single_gpu": all(e[1] for e in everyone), "top3": mine[:3],
                        "hypotheses_per_rank": -(-args.hyp // world)}

    # ---- end-to-end through the public estimator call with HOST buffers (pinned query in, results out)
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_e2e = step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = units / e2e_s
    h2d = query_host.numel() * 4 + 9 * 8                          # query crop + K^-1 of the template camera
    d2h = 3 * 4 + 3 * 4 + 3 * 8 * 8                               # top-3 idx, top-3 scores, 3 extents rows
    mesh_bytes = sum(t.numel() * t.element_size() for t in mesh._device_cache.get(str(dev), ()) if t is not None)

    if rank != 0:
        sg.close()
        return
    pk = peaks()
    tensor_kinds = {k for k in kinds if k.startswith("gemm") or k == "attention"}
    total_kernel_ms = sum(v[0] for v in kinds.values())
    breakdown = {}
    for name, (ms, work, n) in sorted(kinds.items(), key=lambda kv: -kv[1][0]):
        ent = {"ms_per_step": ms / args.steps, "share": ms / total_kernel_ms, "launches_per_step": n / args.steps}
        if name in tensor_kinds:
            ent["tflops"] = work / ms / 1e9
        else:
            ent["gbs"] = work / ms / 1e6
        breakdown[name] = ent
    dom = max((k for k in kinds if k.startswith("gemm")), key=lambda k: kinds[k][0])
    dms, dwork, dn = kinds[dom]
    achieved = dwork / dms / 1e9
    traffic = None
    tfile = ROOT / "profiles" / "kernel_traffic.json"
    if tfile.exists() and not strong:
        traffic = json.loads(tfile.read_text()).get(dom)
    roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"],
                "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"], "traffic": traffic,
                "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step)",
                "avg_launch_ms": dms / dn, "algorithmic_flops_per_launch": dwork / dn}
    per_gpu_hyp_s = value / world
    hyp_local = -(-args.hyp // world) if strong else args.hyp
    vit_tflops = per_gpu_hyp_s * vit_gflop(args.res, args.layer) * (hyp_local + 1) / hyp_local / 1e3
    out = {"metric": METRIC, "value": value, "unit": "hyp/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": workload_config(args), "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": "hyp/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": 1e3 * e2e_s / args.steps,
                   "api": "DinoPoseEstimator.forward_mesh(query_host_pinned, mesh, K, bbox, est_scale, layer) -> TCO, scores",
                   "resident": f"ViT weights and the retrieved mesh ({mesh_bytes} B, uploaded once per mesh) stay in HBM "
                               "across steps, like the reference's model weights and its cached template features"},
           "gpu_launches": int(launches),
           "roofline": roofline,
           "vit_flop_roofline": {"achieved_tflops_per_gpu": vit_tflops,
                                 "frac_of_sustained": vit_tflops / pk["bf16_sustained"],
                                 "frac_of_burst": vit_tflops / pk["bf16_burst"],
                                 "gflop_per_hypothesis": vit_gflop(args.res, args.layer)},
           "kernels": breakdown,
           "kernel_time_share_of_step": total_kernel_ms / args.steps / ms_per_step,
           "best_hypothesis": int(res_e2e["top_indices"][0])}
    if strong:
        strong_check["exchange"] = ("peer memory: fp_score_publish stores into every rank's buffer over NVLink, "
                                    "fp_topk_after_exchange waits on the device" if peer else
                                    "fp_allgather_scores (NCCL from the C ABI)") if world > 1 else "none (one GPU)"
        out["strong_scaling"] = strong_check
        out["latency_ms_per_proposal"] = ms_per_step
    if world == 1 and not args.no_cpu_baseline:
        out["parity"] = parity_block(est, mesh, args, args.ref_hyp)
        # ~15 s of CPU work: one warm-up sample (thread pools, page faults) of ref_hyp hypotheses, then cpu_samples timed
        # samples of cpu_hyp hypotheses each
        CpuReference(args.ref_hyp, args.res, args.layer, args.hyp).run("fp32")
        ref = CpuReference(min(args.cpu_hyp, args.hyp), args.res, args.layer, args.hyp)
        mode = "fp32" if ref.run("fp32")[0] >= ref.run("eager")[0] else "eager"      # the faster arithmetic on these cores
        other = ref.run("eager" if mode == "fp32" else "fp32")[0]
        runs = [ref.run(mode) for _ in range(max(1, args.cpu_samples))]
        v = statistics.mean(r[0] for r in runs)
        dt = sum(r[1] for r in runs)
        st = {k: statistics.mean(r[2][k] for r in runs) for k in runs[0][2]}
        out["cpu_baseline"] = {"value": v, "unit": "hyp/s", "cores": ref.cores, "kind": "port",
                               "sample": f"mean of {len(runs)} samples ({dt:.1f} s in total), each: "
                                         + ref.sample_text(dt, st, other, mode=mode)}
    sg.close()
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------- secondary configs
# BASELINE.json configs 3, 4 and 5 through the same harness (the headline line stays configs[1] = --config pose).
def setup_ffa(args, rank, world, dev):
    """configs[2]: extract_retrieval_features --feature ffa --layer 22 --batch_size 256 over 50k synthetic template
    renders (reference scripts/extract_retrieval_features.py:40-70)."""
    from freepose_b200 import ops
    from freepose_b200.pipeline.estimators.pose_estimator import DinoPoseEstimator
    from freepose_b200.synthetic import synthetic_mesh
    from freepose_b200.vit_weights import synthetic_state_dict
    B = args.batch
    est = DinoPoseEstimator(n_poses=B, cache_size=0, cache_dir=f"/tmp/fp_bench_ffa_{rank}",
                            weights=synthetic_state_dict(seed=0, depth=args.layer), resolution=args.res, chunk=B)
    meshes = [synthetic_mesh(i + 4 * rank, subdivisions=5) for i in range(4)]
    state = {"i": 0}
    fe = est.feature_extractor

    def step_device():
        mesh = meshes[state["i"] % len(meshes)]
        state["i"] += 1
        rgb, depth = est.renderer.render_device(mesh)
        patches, _, mask, _ = est.renderer.proposals_device(rgb, depth, args.res, to_patches=True)
        feats = fe.forward_patches(patches, res=args.res, layer=args.layer)
        return ops.ffa_pool(feats, mask)

    # e2e: what the reference script does per mesh -- host templates + masks in, (views, 1024) fp32 out
    rgb, depth = est.renderer.render_device(meshes[0])
    templates = (rgb.float() / 255).permute(0, 3, 1, 2).contiguous().cpu().pin_memory()
    masks = (depth > 0).cpu().pin_memory()

    def step_e2e():
        feats = fe(templates, layer=args.layer, feature_type="patch")
        pooled, valid = ops.ffa_pool(feats, masks.to(dev, non_blocking=True))
        return pooled.cpu().numpy()[(valid > 0).cpu().numpy()]

    steps_for_50k = -(-50000 // (B * world))
    return dict(metric=f"template images/sec (raster+ViT-L{args.layer}+FFA pool) @{args.res}^2", unit="img/s",
                units_per_step=B, step_device=step_device, step_e2e=step_e2e,
                h2d=templates.numel() * 4 + masks.numel(), d2h=B * 1024 * 4 + B * 4,
                api="DINOv2FeatureExtractor.forward(templates_host, layer, 'patch') + ops.ffa_pool -> .npy rows",
                flops_per_unit=vit_gflop(args.res, args.layer) * 1e9,
                workload={"workload": f"extract_retrieval_features --feature ffa --layer {args.layer} --batch_size {B} "
                                      f"(BASELINE configs[2]): {steps_for_50k * B * world} synthetic template renders at "
                                      f"--steps {steps_for_50k}", "batch": B, "crop": args.res, "layer": args.layer,
                          "mesh_faces": 20480, "l2": "activations of a 256-image batch ~1.4 GB >> 126 MB L2"})


def setup_refiner(args, rank, world, dev):
    """configs[4], the part of smooth_poses_video that is the hot-path pattern: TrackingRefiner.pose_confidence for 64
    refinement renders per frame (reference tracking_refiner.py:45-100): roi_align 518^2 + render at the cropped K +
    2 x ViT-B/14-reg (1374 tokens) + masked per-patch cosine."""
    from freepose_b200.pipeline.estimators.tracking_refiner import TrackingRefiner
    from freepose_b200.pipeline.utils import generate_poses
    from freepose_b200.synthetic import synthetic_mesh
    from freepose_b200.vit_weights import VITB14_REG, synthetic_state_dict
    n = args.batch
    ref = TrackingRefiner(weights=synthetic_state_dict(VITB14_REG, seed=0), chunk=n)
    mesh = synthetic_mesh(4 + rank, subdivisions=5, scale=0.1)
    rng = np.random.default_rng(rank)
    K = np.array([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
    frame_dev = torch.rand(3, 480, 640, device=dev)
    frame_host = (frame_dev * 255).to(torch.uint8).permute(1, 2, 0).contiguous().cpu().pin_memory()
    Ts = []
    for p in generate_poses(n):
        T = np.array(p)
        T[:3, 3] = [rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(0.5, 0.7)]
        Ts.append(T)
    g = 518 // 14
    tokens = g * g + 5
    d, mlp, L = VITB14_REG.embed_dim, VITB14_REG.mlp_dim, VITB14_REG.depth
    flops = 2 * (2 * 588 * d * g * g + L * ((8 * d * d + 4 * d * mlp) * tokens + 4 * tokens * tokens * d))   # 2 forwards per render
    return dict(metric="refinement renders/sec (roi_align+raster+2xViT-B/14@518^2+masked cosine)", unit="renders/s",
                units_per_step=n, step_device=lambda: ref.pose_confidences(mesh, [frame_dev] * n, K, Ts),
                step_e2e=lambda: ref.pose_confidences(mesh, [frame_host.numpy()] * n, K, Ts).cpu().numpy(),
                h2d=n * frame_host.numel() + n * (16 + 9) * 4, d2h=n * g * g * 4,
                api="TrackingRefiner.pose_confidences(mesh, frames_host_u8, K, transforms) -> (n,37,37) confidences",
                flops_per_unit=float(flops),
                workload={"workload": f"smooth_poses_video refiner confidence pass (BASELINE configs[4]): {n} refinement "
                                      "renders per frame, ViT-B/14-reg at 518^2 on the photo crop and on the render",
                          "renders_per_frame": n, "crop": 518, "tokens": tokens, "mesh_faces": 20480,
                          "l2": "activations of 2 x 64 x 1374 tokens ~1.6 GB >> 126 MB L2"})


def setup_video(args, rank, world, dev):
    """configs[3]: dino_inference_video on a 640x480 synthetic video with 8 proposals per frame: Proposals -> crop ->
    DinoOnlinePoseEstimator.forward (coarse on the first frame, fine around prev_pose afterwards, reference
    scripts/dino_inference_video.py:122-182).  Object tracks are dealt round-robin to the GPUs (one track per GPU at 8)."""
    from freepose_b200 import cli, ops
    from freepose_b200.pipeline.estimators.online_pose_estimator import DinoOnlinePoseEstimator
    from freepose_b200.pipeline.proposals import Proposals
    from freepose_b200.vit_weights import synthetic_state_dict
    n_obj = 8
    model = DinoOnlinePoseEstimator(n_coarse_poses=args.hyp, n_fine_poses=20000, cache_size=50,
                                    cache_dir=f"/tmp/fp_bench_video_{rank}", resolution=args.res, chunk=args.hyp + 1,
                                    weights=synthetic_state_dict(seed=0, depth=args.layer))
    templates = cli.SyntheticTemplates(n_obj, args.hyp, args.res, crop=True, subdivisions=4)
    mine = [j for j in range(n_obj) if j % world == rank]
    meshes_r = {j: templates.mesh(j) for j in mine}
    meshes_full = {j: meshes_r[j].copy().apply_scale(4.0) for j in mine}
    entries = {j: templates[j] for j in mine}
    # scene: the 8 objects on a 4 x 2 grid, 3 m away (K from the image diagonal, dino_inference_video.py:116-118)
    h, w = 480, 640
    f = float(np.sqrt(h ** 2 + w ** 2))
    K = np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1]])
    rng = np.random.default_rng(0)
    img = rng.integers(0, 60, (h, w, 3), dtype=np.uint8)
    boxes, masks = [], []
    for j in range(n_obj):
        qm, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(qm) < 0:
            qm[:, 0] = -qm[:, 0]
        pose = np.eye(4); pose[:3, :3] = qm
        z = 3.0
        pose[:3, 3] = [((80 + 160 * (j % 4)) - w / 2) * z / f, ((120 + 240 * (j // 4)) - h / 2) * z / f, z]
        metric = templates.mesh(j).copy().apply_scale(4.0 * 0.3)
        rgb, depth = ops.rasterize_mesh(metric, torch.from_numpy(pose[None]).float().to(dev), f, f, w / 2, h / 2, 640)
        rgb, depth = rgb[0, :h, :w].cpu().numpy(), depth[0, :h, :w].cpu().numpy()
        m = depth > 0
        img[m] = rgb[m]
        ys, xs = np.nonzero(m)
        boxes.append([xs.min(), ys.min(), xs.max(), ys.max()])
        masks.append(m)
    boxes, masks = np.array(boxes), np.array(masks)
    prev = {j: None for j in mine}
    count = {"hyp": 0}

    def step_e2e():
        props = Proposals(img, {"boxes": torch.from_numpy(boxes), "masks": torch.from_numpy(masks)}, args.res,
                          bbox_extend=0.05)
        items = [dict(proposal=props.proposals[j], proposal_mask=props.proposals_masks[j], template_dict=entries[j],
                      mesh=meshes_full[j], K=K, bbox=boxes[j].astype(np.float64), est_scale=0.3, prev_pose=prev[j])
                 for j in mine]
        first = [prev[j] is None for j in mine]
        if args.per_proposal:                      # the reference's loop: one estimator call per proposal
            outs = [model(it["proposal"], it["proposal_mask"], it["template_dict"], it["mesh"], it["K"], it["bbox"],
                          it["est_scale"], prev_pose=it["prev_pose"], neighborhood=15, layer=args.layer, batch_size=128)
                    for it in items]
        else:                                      # all proposals of the frame in one ViT pass
            outs = model.forward_batch(items, neighborhood=15, layer=args.layer, batch_size=128)
        for j, out, f in zip(mine, outs, first):
            prev[j] = out["TCO"][0]
            count["hyp"] += len(out["selected_poses"]) + (args.hyp if f else 0)
        return outs

    def reset():
        for j in mine:
            prev[j] = None
        model.coarse_estimator.feature_cache.clear()
        model.coarse_estimator._device_cache.clear()

    return dict(metric=f"video frames/sec (8 proposals/frame, coarse->fine, crops @{args.res}^2)", unit="frames/s",
                units_per_step=1.0 / world, step_device=None, step_e2e=step_e2e, counters=count, reset=reset,
                h2d=img.size + masks.size + boxes.size * 4, d2h=len(mine) * (8 * 8 + 4 + 4),
                api="Proposals(frame_host_u8, masks, boxes) + DinoOnlinePoseEstimator." +
                    ("forward(..., prev_pose) per proposal" if args.per_proposal else "forward_batch(proposals of the frame)"),
                flops_per_unit=None,
                workload={"workload": "dino_inference_video (BASELINE configs[3]): 640x480 synthetic frames, 8 proposals per "
                                      f"frame, {args.hyp} coarse hypotheses on the first frame, then the fine poses within 15 "
                                      "degrees of prev_pose out of 20000 (~19 per proposal), static scene",
                          "proposals_per_frame": n_obj, "tracks_per_gpu": len(mine), "crop": args.res, "layer": args.layer,
                          "l2": "per-call working set ~60 MB of activations (20 crops): fits L2; weights (554 MB) do not"})


def run_secondary(args):
    from freepose_b200 import _lib
    from freepose_b200.distributed import init_from_env
    import torch.distributed as dist
    rank, local, world = init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU"
    dev = torch.device("cuda", local)
    lib = _lib.load()
    cfg = {"ffa": setup_ffa, "video": setup_video, "refiner": setup_refiner}[args.config](args, rank, world, dev)
    step = cfg["step_device"] or cfg["step_e2e"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    lib.fp_profile_reset(); lib.fp_profile_enable(1)
    launches0 = lib.fp_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if "counters" in cfg:
        cfg["counters"]["hyp"] = 0
    if "reset" in cfg:
        cfg["reset"]()      # video: the timed frames start a NEW video (no previous poses, no cached template features)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    lib.fp_profile_enable(0)
    launches = lib.fp_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    units = cfg["units_per_step"] * world * args.steps
    value = units / (ms_total / 1e3)
    kinds = {}
    for k in range(lib.fp_profile_num_kinds()):
        ms, work, n = C.c_double(), C.c_double(), C.c_longlong()
        _lib.check(lib.fp_profile_collect(k, C.byref(ms), C.byref(work), C.byref(n)), "fp_profile_collect")
        if n.value:
            kinds[lib.fp_profile_kind_name(k).decode()] = (ms.value, work.value, n.value)
    lib.fp_profile_reset()
    hyp_timed = cfg.get("counters", {}).get("hyp")
    # e2e with host buffers
    for _ in range(2):
        cfg["step_e2e"]()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cfg["step_e2e"]()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    if rank != 0:
        return
    pk = peaks()
    tensor_kinds = {k for k in kinds if k.startswith("gemm") or k == "attention"}
    total_kernel_ms = sum(v[0] for v in kinds.values())
    breakdown = {}
    for name, (ms, work, n) in sorted(kinds.items(), key=lambda kv: -kv[1][0]):
        ent = {"ms_per_step": ms / args.steps, "share": ms / total_kernel_ms, "launches_per_step": n / args.steps}
        ent["tflops" if name in tensor_kinds else "gbs"] = work / ms / (1e9 if name in tensor_kinds else 1e6)
        breakdown[name] = ent
    dom = max(kinds, key=lambda k: kinds[k][0])
    dms, dwork, dn = kinds[dom]
    is_tensor = dom in tensor_kinds
    achieved = dwork / dms / (1e9 if is_tensor else 1e6)
    peak = pk["bf16_sustained"] if is_tensor else pk["hbm"]
    out = {"metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg["workload"], "clocks": clocks,
           "value_includes_host_io": cfg["step_device"] is None,
           "e2e": {"value": units / e2e_s, "unit": cfg["unit"], "h2d_bytes_per_step": cfg["h2d"],
                   "d2h_bytes_per_step": cfg["d2h"], "ms_per_step": 1e3 * e2e_s / args.steps, "api": cfg["api"]},
           "gpu_launches": int(launches),
           "roofline": {"kernel": dom, "bound": "tensor" if is_tensor else "hbm", "achieved": achieved, "peak": peak,
                        "unit": "TFLOP/s" if is_tensor else "GB/s", "frac": achieved / peak, "traffic": None,
                        "peak_source": pk["source"], "avg_launch_ms": dms / dn},
           "kernels": breakdown, "kernel_time_share_of_step": total_kernel_ms / args.steps / (ms_total / args.steps),
           "cpu_baseline": None}
    if cfg["flops_per_unit"]:
        tf = value / world * cfg["flops_per_unit"] / 1e12
        out["vit_flop_roofline"] = {"achieved_tflops_per_gpu": tf, "frac_of_sustained": tf / pk["bf16_sustained"]}
    if hyp_timed is not None:
        out["hypotheses_per_sec"] = hyp_timed * world / (ms_total / 1e3)
        out["hypotheses_per_frame"] = hyp_timed * world / args.steps
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hyp", type=int, default=520)
    ap.add_argument("--res", type=int, default=224)
    ap.add_argument("--layer", type=int, default=22)
    ap.add_argument("--chunk", type=int, default=521)
    ap.add_argument("--ref-hyp", type=int, default=8, help="hypotheses per sample of the reference arm and of the parity block")
    ap.add_argument("--cpu-hyp", type=int, default=48, help="hypotheses per timed cpu_baseline sample of the B200 arm")
    ap.add_argument("--cpu-samples", type=int, default=3, help="timed cpu_baseline samples (their mean is reported)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline and parity legs")
    ap.add_argument("--config", default="pose", choices=["pose", "ffa", "video", "refiner"],
                    help="pose = BASELINE configs[1] (the headline line); ffa / video / refiner = configs[2] / [3] / [4]")
    ap.add_argument("--per-proposal", action="store_true",
                    help="video: one estimator call per proposal (the reference's loop) instead of forward_batch per frame")
    ap.add_argument("--batch", type=int, default=None, help="ffa: renders per batch (256); refiner: renders per frame (64)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="strong scaling: how the scores travel (peer-memory stores fused into the score kernel, or NCCL)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: one proposal per GPU per step; strong: one proposal, hypotheses sharded over the GPUs")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 64 if args.config == "refiner" else 256
    if args.impl == "reference":
        run_reference(args)
    elif args.config != "pose":
        run_secondary(args)
    else:
        run_b200(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
