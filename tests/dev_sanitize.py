"""Developer aid (not a pytest file): one small launch of every kernel family, for compute-sanitizer.

    compute-sanitizer --tool memcheck   python tests/dev_sanitize.py
    compute-sanitizer --tool racecheck  python tests/dev_sanitize.py attention score raster

Sizes are small (memcheck slows kernels 20-50x) but ragged on purpose: row counts that are not tile multiples, the
5-row attention tail, a clipped triangle, B not a multiple of the score stage's item size."""
import os
import sys

import numpy as np
import torch

os.environ.setdefault("PYTORCH_NO_CUDA_MEMORY_CACHING", "1")   # every tensor its own allocation: overruns are caught
sys.path.insert(0, ".")
from freepose_b200 import ops  # noqa: E402
from freepose_b200._lib import FP_EPI_BIAS, FP_EPI_BIAS_GELU, FP_EPI_BIAS_LS_RES, FP_EPI_PATCH_EMBED  # noqa: E402

dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)


def gemm():
    for M in (77, 2500):                       # 1-CTA tiles / 2-CTA clusters, both ragged
        a = torch.randn(M, 1024, device=dev).to(bf)
        w = (torch.randn(3072, 1024, device=dev) / 32).to(bf)
        b = torch.randn(3072, device=dev).to(bf)
        ops.gemm(a, w, b, FP_EPI_BIAS)
        w1 = (torch.randn(4096, 1024, device=dev) / 32).to(bf)
        h = ops.gemm(a, w1, torch.randn(4096, device=dev).to(bf), FP_EPI_BIAS_GELU)
        w2 = (torch.randn(1024, 4096, device=dev) / 64).to(bf)
        x = torch.randn(M, 1024, device=dev).to(bf)
        ops.gemm(h, w2, torch.randn(1024, device=dev).to(bf), FP_EPI_BIAS_LS_RES, gamma=torch.rand(1024, device=dev).to(bf),
                 residual=x)
    B, P, T = 3, 16, 21
    pa = torch.randn(B * P, 640, device=dev).to(bf)
    pw = (torch.randn(1024, 640, device=dev) / 24).to(bf)
    tok = torch.zeros(B * T, 1024, dtype=bf, device=dev)
    ops.gemm(pa, pw, torch.randn(1024, device=dev).to(bf), FP_EPI_PATCH_EMBED, out=tok,
             pos=torch.randn(1 + P, 1024, device=dev).to(bf), patches_per_img=P, tokens_per_img=T, token_offset=5)


def attention():
    for B, T, H in ((3, 261, 16), (2, 905, 16), (1, 1374, 12), (2, 21, 16)):   # split / pair / pair (ViT-B) / single stream
        qkv = torch.randn(B * T, 3 * H * 64, device=dev).to(bf)
        ops.attention(qkv, B, T, heads=H)


def layernorm():
    for rows, D in ((1000, 1024), (333, 768)):
        x = torch.randn(rows, D, device=dev).to(bf)
        ops.layernorm(x, torch.rand(D, device=dev).to(bf), torch.randn(D, device=dev).to(bf))


def score():
    for B, P, D in ((13, 256, 1024), (5, 16, 256)):
        f = torch.randn(B, P, D, device=dev).to(bf)
        q = torch.randn(P, D, device=dev).to(bf)
        ops.score_topk(f, q, k=3)
        ops.score_topk(f, q, k=1, weights=torch.rand(B, P, device=dev), return_patch_scores=True)
    f = torch.randn(7, 256, 1024, device=dev).to(bf)
    m = (torch.rand(7, 224, 224, device=dev) > 0.5)
    ops.ffa_pool(f, m)


def retrieval():
    table = ops.normalize_rows(torch.randn(3001, 1024, device=dev))
    for Q in (1, 8, 33):
        q = ops.normalize_rows(torch.randn(Q, 1024, device=dev))
        s = ops.retrieval_scan(table, q)
        ops.topk_rows(s, 100)


def raster():
    from freepose_b200.pipeline.utils import generate_poses
    from freepose_b200.synthetic import synthetic_mesh
    mesh = synthetic_mesh(0, subdivisions=3)
    poses = torch.from_numpy(np.array(generate_poses(9))).float().cuda()
    poses[4, 2, 3] = 0.1                         # the camera inside the mesh: clipped triangles (hard path)
    poses[5, 0, 3] = 40.0                        # the object far off screen: empty screen box
    for msaa in (4, 1):
        rgb, depth = ops.rasterize_mesh(mesh, poses, 320.0, 320.0, 112.0, 112.0, 224, msaa=msaa)
    bbox, count, mask = ops.mask_bbox(depth, fallback=(56, 168), min_count=100, return_mask=True)
    ops.crop_resize_pad(rgb, bbox, 224, to_patches=True)
    ops.crop_resize_pad(rgb, bbox, 224, to_patches=False)
    K = np.array([[320.0, 0, 112], [0, 320.0, 112], [0, 0, 1]])
    ops.depth_extents(depth, K, view_idx=torch.tensor([0, 3], device=dev, dtype=torch.int32))


ALL = {"gemm": gemm, "attention": attention, "layernorm": layernorm, "score": score, "retrieval": retrieval, "raster": raster}

if __name__ == "__main__":
    names = sys.argv[1:] or list(ALL)
    for n in names:
        ALL[n]()
        torch.cuda.synchronize()
        print("ran", n, flush=True)
    print("dev_sanitize: done")
