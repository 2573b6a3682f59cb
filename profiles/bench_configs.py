"""Secondary measurements for the record (NOT bench.py's headline line): BASELINE.json configs 3 and 5 and the mesh
retrieval scan on one B200, device-resident, CUDA events.  Run:  gpurun -- python profiles/bench_configs.py"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from freepose_b200 import ops  # noqa: E402
from freepose_b200.pipeline.estimators.pose_estimator import DinoPoseEstimator  # noqa: E402
from freepose_b200.pipeline.estimators.tracking_refiner import TrackingRefiner  # noqa: E402
from freepose_b200.pipeline.utils import generate_poses  # noqa: E402
from freepose_b200.synthetic import synthetic_mesh  # noqa: E402
from freepose_b200.vit_weights import VITB14_REG, synthetic_state_dict  # noqa: E402


def timed(fn, n, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
# ---- config 3: extract_retrieval_features --feature ffa --layer 22 --batch_size 256 over 50k synthetic template renders
B = 256
est = DinoPoseEstimator(n_poses=B, cache_size=0, cache_dir="/tmp/fp_cfg3", weights=synthetic_state_dict(seed=0, depth=22),
                        resolution=224, chunk=B)
meshes = [synthetic_mesh(i, subdivisions=5) for i in range(4)]
state = {"i": 0}


def ffa_batch():
    mesh = meshes[state["i"] % len(meshes)]
    state["i"] += 1
    rgb, depth = est.renderer.render_device(mesh)
    patches, bbox, mask, _ = est.renderer.proposals_device(rgb, depth, 224, to_patches=True)
    feats = est.feature_extractor.engine.forward(patches, layer=22, feature_type="patch", res=224)
    return ops.ffa_pool(feats, mask)


n_batches = 50000 // B + 1
ms = timed(ffa_batch, n_batches)
out["config3_ffa_extraction"] = {"images_per_s": B / ms * 1e3, "ms_per_batch": ms, "batch": B, "images": n_batches * B,
                                 "crop": 224, "what": "raster + mask/bbox + crop + ViT-L/14-reg layer 22 + FFA pooling"}

# ---- config 5: refiner confidence pass, 64 renders per frame (ViT-B/14-reg at 518^2 on photo crop and render)
ref = TrackingRefiner(weights=synthetic_state_dict(VITB14_REG, seed=0), chunk=64)
mesh = synthetic_mesh(4, subdivisions=5, scale=0.1)
rng = np.random.default_rng(0)
K = np.array([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
frame = torch.rand(3, 480, 640, device="cuda")
Ts = []
for p in generate_poses(64):
    T = np.array(p)
    T[:3, 3] = [rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(0.5, 0.7)]
    Ts.append(T)
ms = timed(lambda: ref.pose_confidences(mesh, [frame] * 64, K, Ts), 5)
out["config5_refiner_confidence"] = {"renders_per_s": 64 / ms * 1e3, "ms_per_frame_of_64": ms,
                                     "what": "roi_align 518^2 + render at cropped K + 2 x ViT-B/14-reg (1374 tokens) + masked cosine"}

# ---- mesh retrieval: coarse scan of the reference's table size + top-100
from freepose_b200.pipeline.retrieval.database import RetrievalDatabase  # noqa: E402
table = torch.randn(46037, 1024, device="cuda")
q = torch.randn(8, 1024, device="cuda")
tn, qn = ops.normalize_rows(table), ops.normalize_rows(q)
ms = timed(lambda: ops.retrieval_scan(tn, qn), 50)
out["retrieval_scan"] = {"ms": ms, "gb_per_s": 46037 * 1024 * 2 / ms / 1e6, "queries": 8, "rows": 46037}
print(json.dumps(out, indent=1))
