"""Developer timing script (not a pytest file): per-kernel timings at the benchmark shapes.
Run:  gpurun -- python tests/dev_gpu_perf.py"""
import sys

import torch

sys.path.insert(0, ".")
from freepose_b200 import ops  # noqa: E402
from freepose_b200._lib import FP_EPI_BIAS, FP_EPI_BIAS_GELU, FP_EPI_BIAS_LS_RES  # noqa: E402

dev, bf = "cuda", torch.bfloat16
B, T = 521, 261
M = B * T
torch.manual_seed(0)


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


x = torch.randn(M, 1024, device=dev).to(bf)
h4 = torch.randn(M, 4096, device=dev).to(bf)
res = torch.randn(M, 1024, device=dev).to(bf)
w_qkv = (torch.randn(3072, 1024, device=dev) / 32).to(bf)
w_proj = (torch.randn(1024, 1024, device=dev) / 32).to(bf)
w_fc1 = (torch.randn(4096, 1024, device=dev) / 32).to(bf)
w_fc2 = (torch.randn(1024, 4096, device=dev) / 64).to(bf)
b3, b1, b4 = (torch.randn(n, device=dev).to(bf) for n in (3072, 1024, 4096))
g = torch.rand(1024, device=dev).to(bf)
o3 = torch.empty(M, 3072, device=dev, dtype=bf)
o4 = torch.empty(M, 4096, device=dev, dtype=bf)

cases = [
    ("qkv  ", lambda: ops.gemm(x, w_qkv, b3, FP_EPI_BIAS, out=o3), 2 * M * 1024 * 3072),
    ("proj ", lambda: ops.gemm(x, w_proj, b1, FP_EPI_BIAS_LS_RES, gamma=g, residual=res), 2 * M * 1024 * 1024),
    ("fc1  ", lambda: ops.gemm(x, w_fc1, b4, FP_EPI_BIAS_GELU, out=o4), 2 * M * 1024 * 4096),
    ("fc2  ", lambda: ops.gemm(h4, w_fc2, b1, FP_EPI_BIAS_LS_RES, gamma=g, residual=res), 2 * M * 4096 * 1024),
]
total = 0
for name, fn, flops in cases:
    ms = timeit(fn)
    total += ms
    print(f"{name} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)
qkv = torch.randn(M, 3072, device=dev).to(bf)
ms = timeit(lambda: ops.attention(qkv, B, T))
total += ms
print(f"attn  {ms:7.3f} ms  {4 * B * 16 * T * T * 64 / ms / 1e9:7.1f} TFLOP/s")
lnw, lnb = torch.ones(1024, device=dev, dtype=bf), torch.zeros(1024, device=dev, dtype=bf)
ms = timeit(lambda: ops.layernorm(x, lnw, lnb))
total += 2 * ms
print(f"ln    {ms:7.3f} ms  {2 * M * 2048 / ms / 1e6:7.1f} GB/s")
print(f"per-block total {total:.3f} ms -> 22 blocks {22 * total:.1f} ms -> {520 / (22 * total / 1e3):.0f} hyp/s (ViT only)")
# cuBLAS reference points (library GEMM without the fused epilogue)
for name, a, w in (("qkv", x, w_qkv), ("fc1", x, w_fc1), ("fc2", h4, w_fc2), ("proj", x, w_proj)):
    ms = timeit(lambda: torch.nn.functional.linear(a, w))
    print(f"cublas {name} {ms:7.3f} ms {2 * M * a.shape[1] * w.shape[0] / ms / 1e9:7.1f} TFLOP/s")
