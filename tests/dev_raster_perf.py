"""Developer timing script (not a pytest file): the rasteriser at the benchmark shape (521 views x 20 480 faces, 224^2).
Run:  gpurun -- python tests/dev_raster_perf.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from freepose_b200 import ops  # noqa: E402
from freepose_b200.pipeline.utils import generate_poses  # noqa: E402
from freepose_b200.synthetic import synthetic_mesh  # noqa: E402

mesh = synthetic_mesh(0, subdivisions=5)
poses = torch.from_numpy(np.array(generate_poses(521))).float().cuda()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = timeit(lambda: ops.rasterize_mesh(mesh, poses, 320.0, 320.0, 112.0, 112.0, 224))
print(f"rasterize 521 views: {ms:.3f} ms  ({521 * 224 * 224 * 7 / ms / 1e6:.1f} GB/s of output)")
