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
    fp32, so the baseline `value` is the fp32 run (the more favourable number for the reference) and the bf16-eager
    rate is reported beside it."""

    def __init__(self, n_hyp: int, res: int, layers: int):
        from freepose_b200.pipeline.utils import generate_poses
        from freepose_b200.synthetic import synthetic_mesh
        from freepose_b200.vit_weights import synthetic_state_dict
        from oracle.pipeline import OraclePipeline, synthetic_query
        torch.set_num_threads(os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        self.n = n_hyp
        self.sd = synthetic_state_dict(seed=0, depth=layers)
        self.mesh = synthetic_mesh(0, subdivisions=5)
        self.res, self.layers = res, layers
        self.query, _ = synthetic_query(self.mesh, res, seed=1)
        self.poses = generate_poses(520)[:n_hyp]
        self.K = np.array([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
        self.bbox = np.array([200.0, 150.0, 330.0, 290.0])
        self._pipes = {}
        self._mk = OraclePipeline

    def run(self, mode: str = "fp32"):
        if mode not in self._pipes:
            self._pipes[mode] = self._mk(self.sd, self.res, mode=mode, layer=self.layers)
        t0 = time.perf_counter()
        self._pipes[mode].forward(self.query, self.mesh, self.K, self.bbox, 0.3, self.poses, k=min(3, self.n))
        dt = time.perf_counter() - t0
        return self.n / dt, dt

    def sample_text(self, dt, bf16_rate=None):
        s = (f"{self.n} hypotheses + 1 query of the 520-hypothesis workload per sample ({dt:.1f} s): oracle/pipeline.py = "
             f"C raster restatement + CropResizePad + PyTorch-eager fp32 ViT-L/14-reg to layer {self.layers} + reference "
             "scoring lines, all host threads; the true pyrender/EGL renderer is not installable offline")
        if bf16_rate is not None:
            s += f"; the same path in the reference's bf16 dtype runs at {bf16_rate:.3f} hyp/s on these cores"
        return s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(args.ref_hyp, args.res, args.layer)
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt = ref.run("fp32")
        if i >= args.warmup:
            rates.append((r, dt))
    value = statistics.mean(r for r, _ in rates)
    dt_mean = statistics.mean(dt for _, dt in rates)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "hyp/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_mean, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
           "config": workload_config(args),
           "cpu_baseline": {"value": value, "unit": "hyp/s", "cores": ref.cores, "kind": "port",
                            "sample": ref.sample_text(dt_mean)},
           "e2e": {"value": value, "unit": "hyp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def workload_config(args):
    return {"workload": f"dino_inference (BASELINE configs[1]): 1 proposal x 1 mesh x {args.hyp} pose hypotheses, "
                        f"raster + DINOv2 ViT-L/14-reg layer {args.layer} + per-patch cosine score/top-3",
            "hypotheses": args.hyp, "crop": args.res, "layer": args.layer, "mesh_faces": 20480, "msaa": 4,
            "proposals_per_step_per_gpu": 1,
            "l2": "per-step working set ~2.9 GB (activations) >> 126 MB L2; no explicit flush needed"}


# ----------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    from freepose_b200 import _lib, ops
    from freepose_b200.distributed import ScoreGather, init_from_env
    from freepose_b200.pipeline.estimators.pose_estimator import DinoPoseEstimator
    from freepose_b200.pipeline.utils import rescaled_extents, tco_from_extents
    from freepose_b200.synthetic import synthetic_mesh
    from freepose_b200.vit_weights import synthetic_state_dict
    import torch.distributed as dist

    rank, local, world = init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    dev = torch.device("cuda", local)
    lib = _lib.load()
    torch.manual_seed(0)

    sd = synthetic_state_dict(seed=0, depth=args.layer)
    mesh = synthetic_mesh(0, subdivisions=5)
    est = DinoPoseEstimator(n_poses=args.hyp, cache_size=0, cache_dir=f"/tmp/fp_bench_cache_{rank}", weights=sd,
                            resolution=args.res, chunk=args.chunk)
    # query: render of the mesh at a held-out rotation (seed 1 + rank) + noise, cropped like a proposal
    rng = np.random.default_rng(1 + rank)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    qpose = np.eye(4); qpose[:3, :3] = q; qpose[2, 3] = 1.1
    rgb, depth = est.renderer.render_device(mesh, [qpose])
    crop, _, _, _ = est.renderer.proposals_device(rgb, depth, args.res, to_patches=False)
    noise = torch.randn(crop.shape, generator=torch.Generator().manual_seed(rank)).to(dev) * 0.02
    query_dev = (crop[0] + noise[0]).clamp(0, 1).contiguous()
    query_host = query_dev.cpu().pin_memory()
    K = np.array([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
    bbox = np.array([200.0, 150.0, 330.0, 290.0])
    r = args.res
    K_t = np.array([[est.renderer.focal, 0, r / 2], [0, est.renderer.focal, r / 2], [0, 0, 1]])
    sg = ScoreGather(args.hyp * world, world, dev)

    def step(query, host_io: bool):
        """One proposal: everything on the device; host_io adds the H2D of the query and the D2H of the result."""
        if host_io:
            query = query.to(dev, non_blocking=True)
        feats, depth, _, qf = est.render_features(mesh, None, layer=args.layer, query=query)  # 520 renders + query
        _, idx, vals, _ = ops.score_topk(feats, qf, k=3, scores_out=sg.local_view(rank))
        all_scores = sg.gather(rank)                       # ONE all-gather of per-hypothesis scores (no-op at N=1)
        ext = ops.depth_extents(depth, K_t, view_idx=idx)
        if not host_io:
            return idx, vals, ext, all_scores
        idx_h, vals_h, ext_h = idx.cpu().numpy(), vals.cpu().numpy(), ext.cpu().numpy()
        tco = []
        for j, i in enumerate(idx_h):
            dx, dy = rescaled_extents(ext_h[j], 0.3, recentre=True)
            tco.append(tco_from_extents(bbox, dx, dy, K, est.mesh_poses[int(i)]))
        return idx_h, vals_h, tco

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
        step(query_dev, False)
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
        step(query_dev, False)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    lib.fp_profile_enable(0)
    launches = lib.fp_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = world * args.hyp * args.steps / (ms_total / 1e3)

    kinds = {}
    for k in range(lib.fp_profile_num_kinds()):
        ms, work, n = C.c_double(), C.c_double(), C.c_longlong()
        _lib.check(lib.fp_profile_collect(k, C.byref(ms), C.byref(work), C.byref(n)), "fp_profile_collect")
        if n.value:
            kinds[lib.fp_profile_kind_name(k).decode()] = (ms.value, work.value, n.value)
    lib.fp_profile_reset()

    # ---- end-to-end through the public estimator call path with HOST buffers (pinned query in, results out)
    for _ in range(2):
        step(query_host, True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_e2e = step(query_host, True)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * args.hyp * args.steps / e2e_s
    h2d = query_host.numel() * 4 + 9 * 8                          # query crop + K^-1
    d2h = 3 * 4 + 3 * 4 + 3 * 8 * 8                               # top-3 idx, top-3 scores, 3 extents rows

    if rank != 0:
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
    if tfile.exists():
        traffic = json.loads(tfile.read_text()).get(dom)
    roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"],
                "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"], "traffic": traffic,
                "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step)",
                "avg_launch_ms": dms / dn, "algorithmic_flops_per_launch": dwork / dn}
    vit_tflops = value * vit_gflop(args.res, args.layer) * (args.hyp + 1) / args.hyp / 1e3
    out = {"metric": METRIC, "value": value, "unit": "hyp/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
           "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": "hyp/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": 1e3 * e2e_s / args.steps, "api": "DinoPoseEstimator.render_features + "
                   "feature_extractor + ops.score_topk + depth_extents + tco_from_extents (= forward_mesh)"},
           "gpu_launches": int(launches),
           "roofline": roofline,
           "vit_flop_roofline": {"achieved_tflops_per_gpu": vit_tflops / world,
                                 "frac_of_sustained": vit_tflops / world / pk["bf16_sustained"],
                                 "frac_of_burst": vit_tflops / world / pk["bf16_burst"],
                                 "gflop_per_hypothesis": vit_gflop(args.res, args.layer)},
           "kernels": breakdown,
           "kernel_time_share_of_step": total_kernel_ms / args.steps / ms_per_step,
           "best_hypothesis": int(res_e2e[0][0])}
    if world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(args.ref_hyp, args.res, args.layer)
        ref.run("fp32")                                    # warm-up (thread pools, page faults)
        v, dt = ref.run("fp32")
        small = CpuReference(2, args.res, args.layer)
        vb, _ = small.run("eager")
        out["cpu_baseline"] = {"value": v, "unit": "hyp/s", "cores": ref.cores, "kind": "port",
                               "sample": ref.sample_text(dt, vb)}
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
    ap.add_argument("--ref-hyp", type=int, default=8, help="hypotheses per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
