"""Dev aid: LayerNorm at the benchmark shape (135 981 x 1024 bf16: 557 MB moved)."""
import sys, torch
sys.path.insert(0, ".")
from freepose_b200 import ops
M = 521 * 261
x = torch.randn(M, 1024, device="cuda").to(torch.bfloat16)
w = torch.ones(1024, device="cuda", dtype=torch.bfloat16); b = torch.zeros(1024, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.layernorm(x, w, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): ops.layernorm(x, w, b)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print("layernorm %.4f ms  %.0f GB/s" % (ms, 2 * M * 2048 / ms / 1e6))
