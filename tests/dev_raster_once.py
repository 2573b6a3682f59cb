"""Dev aid for ncu launch lists: two rasterize calls at the benchmark shape (the first warms up)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from freepose_b200 import ops
from freepose_b200.pipeline.utils import generate_poses
from freepose_b200.synthetic import synthetic_mesh
mesh = synthetic_mesh(0, subdivisions=5)
poses = torch.from_numpy(np.array(generate_poses(521))).float().cuda()
for _ in range(2):
    ops.rasterize_mesh(mesh, poses, 320.0, 320.0, 112.0, 112.0, 224)
torch.cuda.synchronize()
