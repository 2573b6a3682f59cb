"""Dev aid: timing of the mesh-retrieval kernels at the reference's table size (46 037 x 1024)."""
import sys, torch
sys.path.insert(0, ".")
from freepose_b200 import ops

M, D = 46037, 1024
dev = "cuda"
table = ops.normalize_rows(torch.randn(M, D, device=dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()                      # L2 flush: the 94 MB table would otherwise sit in the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for Q in (1, 8, 32):
    q = ops.normalize_rows(torch.randn(Q, D, device=dev))
    t = timeit(lambda: ops.retrieval_scan(table, q))
    print("scan Q=%2d  %.3f ms  %.0f GB/s (table bytes / time)" % (Q, t, M * D * 2 / t / 1e6))
    s = ops.retrieval_scan(table, q)
    t = timeit(lambda: ops.topk_rows(s, 100))
    print("top-100 Q=%2d  %.3f ms" % (Q, t))
views = ops.normalize_rows(torch.randn(100 * 600, D, device=dev))
start = (torch.arange(100, device=dev) * 600).long()
count = torch.full((100,), 600, dtype=torch.int32, device=dev)
for Q in (1, 8):
    cand = torch.arange(100, device=dev, dtype=torch.int32).repeat(Q, 1).contiguous()
    q = ops.normalize_rows(torch.randn(Q, D, device=dev))
    t = timeit(lambda: ops.retrieval_fine(views, start, count, 600, cand, q, 5))
    print("fine Q=%d (100 candidates x 600 views)  %.3f ms  %.0f GB/s" % (Q, t, Q * 100 * 600 * D * 2 / t / 1e6))
t = timeit(lambda: ops.normalize_rows(torch.empty(M, D, device=dev)))
