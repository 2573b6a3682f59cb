"""Dev aid: per-phase cycle counters of the attention kernel (softmax warp 4 and the MMA warp of CTA 0)."""
import os, sys, torch
sys.path.insert(0, ".")
from freepose_b200 import ops
B = int(os.environ.get("ATTN_B", "521"))
T = int(sys.argv[1]) if len(sys.argv) > 1 else 261
qkv = torch.randn(B * T, 3072, device="cuda").to(torch.bfloat16)
for _ in range(3): ops.attention(qkv, B, T)
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["FP_ATTN_DBG"] = str(dbg.data_ptr())
ops.attention(qkv, B, T)
torch.cuda.synchronize()
del os.environ["FP_ATTN_DBG"]
d = dbg.cpu().tolist()
n = max(d[7], 1)
names = ["sm wait s_full", "sm pass1", "sm max xchg", "sm pass2", "sm sum xchg", "sm wait o_full", "sm epilogue", "tiles",
         "mma wait q_full", "mma wait s_empty", "mma S issue", "mma wait o_empty", "mma PV (p_full waits)"]
for i, nm in enumerate(names):
    print("%-24s %10.0f cycles/tile" % (nm, d[i] / n if i != 7 else d[i]))
