"""Developer aid (not a pytest file): per-phase cycle counters of the attention kernel's TIMING instance.
Run:  gpurun -- python tests/dev_attn_phases.py"""
import os
import sys

import torch

sys.path.insert(0, ".")
from freepose_b200 import ops  # noqa: E402

B, T = 521, 261
torch.manual_seed(0)
qkv = torch.randn(B * T, 3072, device="cuda").to(torch.bfloat16)
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
for _ in range(2):
    ops.attention(qkv, B, T)
torch.cuda.synchronize()
os.environ["FP_ATTN_DBG"] = str(dbg.data_ptr())
ops.attention(qkv, B, T)
torch.cuda.synchronize()
del os.environ["FP_ATTN_DBG"]
d = dbg.cpu().tolist()
n = max(d[7], 1)
names = {0: "wait s_full", 1: "pass1 max", 2: "max exchange", 3: "pass2 exp", 6: "deferred epilogue"}
print("softmax warp 4 of CTA 0, cycles per tile over", n, "tiles")
for k, v in names.items():
    print(f"  {v:18s} {d[k] / n:8.1f}")
print("  total              %8.1f" % (sum(d[k] for k in names) / n))


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


ms = timeit(lambda: ops.attention(qkv, B, T), n=30)
print(f"attention B={B} T={T}: {ms:.3f} ms  {4 * B * 16 * T * T * 64 / ms / 1e9:.1f} TFLOP/s")
