"""Profiling driver: a few device-resident steps of the benchmark workload (used under ncu via gpurun).
    python profiles/prof_step.py [steps] [hyp]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from freepose_b200 import ops  # noqa: E402
from freepose_b200.pipeline.estimators.pose_estimator import DinoPoseEstimator  # noqa: E402
from freepose_b200.synthetic import synthetic_mesh  # noqa: E402
from freepose_b200.vit_weights import synthetic_state_dict  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
hyp = int(sys.argv[2]) if len(sys.argv) > 2 else 520
layers = int(sys.argv[3]) if len(sys.argv) > 3 else 22
sd = synthetic_state_dict(seed=0, depth=layers)
mesh = synthetic_mesh(0, subdivisions=5)
est = DinoPoseEstimator(n_poses=hyp, cache_size=0, cache_dir="/tmp/fp_prof_cache", weights=sd, resolution=224, chunk=hyp + 1)
query = torch.rand(3, 224, 224, device="cuda")
for _ in range(steps):
    feats, depth, _, qf = est.render_features(mesh, None, layer=layers, query=query)  # as bench.py does
    ops.score_topk(feats, qf, k=3)
torch.cuda.synchronize()
print("done")
