"""Dev aid: timing of the per-hypothesis score kernel at the benchmark size (520 x 256 x 1024 bf16 = 273 MB)."""
import sys, torch
sys.path.insert(0, ".")
from freepose_b200 import ops

B, P, D = 520, 256, 1024
feats = torch.randn(B, P, D, device="cuda").to(torch.bfloat16)
q = torch.randn(P, D, device="cuda").to(torch.bfloat16)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


t = timeit(lambda: ops.score_topk(feats, q, k=3))
print("score_topk (prep_query + score + topk) %.3f ms  %.0f GB/s" % (t, B * P * D * 2 / t / 1e6))
t = timeit(lambda: ops.score_topk(feats, q, k=0))
print("score only (prep_query + score)        %.3f ms  %.0f GB/s" % (t, B * P * D * 2 / t / 1e6))
