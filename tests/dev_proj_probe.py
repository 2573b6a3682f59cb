"""Developer timing script (not a pytest file): the proj-shaped GEMM (M=135 981, N=K=1024) with each epilogue.
Finding: bias-only 1350 TFLOP/s, bias+GELU 1299, bias+LayerScale+residual 1147 -- the residual read makes the launch move
834 MB (x in, residual in, x out) for 285 GFLOP, i.e. 130 us of HBM time next to 180 us of tensor time."""
import sys, torch
sys.path.insert(0, ".")
from freepose_b200 import ops
from freepose_b200._lib import FP_EPI_BIAS, FP_EPI_BIAS_GELU, FP_EPI_BIAS_LS_RES
dev, bf = "cuda", torch.bfloat16
B, T = 521, 261
M = B * T
torch.manual_seed(0)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x = torch.randn(M, 1024, device=dev).to(bf)
res = torch.randn(M, 1024, device=dev).to(bf)
w = (torch.randn(1024, 1024, device=dev) / 32).to(bf)
b1 = torch.randn(1024, device=dev).to(bf); g = torch.rand(1024, device=dev).to(bf)
o = torch.empty(M, 1024, device=dev, dtype=bf)
fl = 2 * M * 1024 * 1024
for name, fn in (("bias only      ", lambda: ops.gemm(x, w, b1, FP_EPI_BIAS, out=o)),
                 ("bias+LS+res    ", lambda: ops.gemm(x, w, b1, FP_EPI_BIAS_LS_RES, gamma=g, residual=res)),
                 ("bias+gelu      ", lambda: ops.gemm(x, w, b1, FP_EPI_BIAS_GELU, out=o))):
    ms = timeit(fn); print(f"proj-shaped {name} {ms:.3f} ms {fl/ms/1e9:7.1f} TF")
