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
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
for _ in range(2):
    ops.attention(qkv, B, T)
torch.cuda.synchronize()
os.environ["FP_ATTN_DBG"] = str(dbg.data_ptr())
ops.attention(qkv, B, T)
torch.cuda.synchronize()
del os.environ["FP_ATTN_DBG"]
d = dbg.cpu().tolist()
if os.environ.get("FP_ATTN_SPLIT", "1") != "0":
    n0, n1 = max(d[3], 1), max(d[9], 1)
    print("two-stream kernel, CTA 0, cycles per tile")
    print("  stream 0 (warp 4):  wait s_full %8.1f | ld+max %8.1f | exp+publish %8.1f | total %8.1f  (%d tiles)" % (
        d[0] / n0, d[1] / n0, d[2] / n0, (d[0] + d[1] + d[2]) / n0, n0))
    print("  stream 1 (warp 8):  wait s_full %8.1f | ld+max %8.1f | epilogue(g-1) %8.1f | exp+publish %8.1f | total %8.1f" % (
        d[4] / n1, d[5] / n1, d[8] / n1, d[6] / n1, sum(d[4:9]) / n1))
    print("  MMA issuer 0:  wait o0_empty %8.1f | PV0 (waits on P) %8.1f | next S0 (waits on Q/KV) %8.1f" % (d[10] / n1, d[11] / n1, d[12] / n1))
    print("  MMA issuer 1:  wait o1_empty %8.1f | PV1 (waits on P) %8.1f | next S1 (waits on Q/KV) %8.1f" % (d[13] / n1, d[14] / n1, d[15] / n1))
else:
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
