"""Developer timing script (not a pytest file): the tiled-key attention kernel at the reference-native crop (420^2 -> 905
tokens) and at the refiner's 518^2 (1374 tokens, 12 heads).  Run:  gpurun -- python tests/dev_attn_long_perf.py"""
import sys

import torch

sys.path.insert(0, ".")
from freepose_b200 import ops  # noqa: E402


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


for B, T, H in ((150, 905, 16), (128, 1374, 12), (521, 261, 16)):
    qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(torch.bfloat16)
    ms = timeit(lambda: ops.attention(qkv, B, T, heads=H))
    print(f"attention B={B} T={T} H={H}: {ms:.3f} ms  {4 * B * H * T * T * 64 / ms / 1e9:.1f} TFLOP/s")
