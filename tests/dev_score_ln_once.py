"""Dev aid for ncu: a few launches of the score stage and of LayerNorm at the benchmark shapes."""
import sys, torch
sys.path.insert(0, ".")
from freepose_b200 import ops
B, P, D = 520, 256, 1024
feats = torch.randn(B, P, D, device="cuda").to(torch.bfloat16)
q = torch.randn(P, D, device="cuda").to(torch.bfloat16)
x = torch.randn(521 * 261, 1024, device="cuda").to(torch.bfloat16)
w = torch.ones(1024, device="cuda", dtype=torch.bfloat16); b = torch.zeros(1024, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.score_topk(feats, q, k=3)
    ops.layernorm(x, w, b)
torch.cuda.synchronize()
