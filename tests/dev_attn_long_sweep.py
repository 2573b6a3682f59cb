import sys, torch
sys.path.insert(0, ".")
from freepose_b200 import ops
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B, T, H in ((148, 1024, 16), (148, 768, 16), (148, 512, 16), (148, 896, 16), (148, 905, 16), (148, 1025, 16), (148, 1280, 16)):
    qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(torch.bfloat16)
    ms = timeit(lambda: ops.attention(qkv, B, T, heads=H))
    nq = (T + 127) // 128; nkb = ((T + 15) // 16 * 16 + 255) // 256
    items = B * H * nq / 148
    print(f"T={T:5d} nq={nq} nkb={nkb}: {ms:.3f} ms  {4*B*H*T*T*64/ms/1e9:6.1f} TF  per item {ms*1e3/items:6.2f} us  per block {ms*1e3/items/nkb:5.2f} us")
