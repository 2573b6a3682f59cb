import sys, torch
sys.path.insert(0, ".")
from freepose_b200 import ops
B = int(__import__("os").environ.get("ATTN_B", "521"))
for T in ([int(a) for a in sys.argv[1:]] or [261]):
    qkv = torch.randn(B*T, 3072, device="cuda").to(torch.bfloat16)
    for _ in range(3): ops.attention(qkv,B,T)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.attention(qkv,B,T)
    e1.record(); torch.cuda.synchronize()
    print("attn T=%d ms %.4f" % (T, e0.elapsed_time(e1)/20))
