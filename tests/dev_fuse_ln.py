"""Dev aid: LayerNorm fused into the residual GEMMs (FP_FUSE_LN, default on) against the standalone launches (=0):
bit-equality of the tokens, repeated, and timing of the whole forward."""
import os
import sys

import torch

sys.path.insert(0, ".")
from freepose_b200.vit_engine import ViTEngine  # noqa: E402
from freepose_b200.vit_weights import synthetic_state_dict  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 4
batches = [int(a) for a in sys.argv[2:]] or [9, 64]
sd = synthetic_state_dict(seed=0, depth=depth)
eng = ViTEngine(sd, device="cuda")
torch.manual_seed(1)


def run(x, fuse):
    os.environ["FP_FUSE_LN"] = str(fuse)
    return eng.forward(x, layer=depth, feature_type="all").clone()


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if os.environ.get("FP_GEMM_DEBUG"):     # perf experiments: timing only (results are wrong by construction)
    x = torch.rand(batches[-1], 3, 224, 224, device="cuda")
    print("FP_GEMM_DEBUG=%s: fc2-fused %.3f ms, unfused %.3f ms" % (os.environ["FP_GEMM_DEBUG"], timeit(lambda: run(x, 1)),
                                                                  timeit(lambda: run(x, 0))))
    sys.exit(0)
ok = True
for B in batches:
    x = torch.rand(B, 3, 224, 224, device="cuda")
    ref = run(x, 0)
    bad = 0
    for rep in range(6):
        got = run(x, 1 + 2 * (rep % 2))
        torch.cuda.synchronize()
        bad += int((got.view(torch.int16) != ref.view(torch.int16)).sum())
    t0 = timeit(lambda: run(x, 0))
    t1 = timeit(lambda: run(x, 1))
    t3 = timeit(lambda: run(x, 3))
    print(f"B={B:4d} rows={B * 261:7d} depth={depth}: mismatching elements over 6 runs = {bad};  "
          f"standalone LN {t0:.3f} ms, fc2-fused {t1:.3f} ms, fc2+proj-fused {t3:.3f} ms", flush=True)
    ok &= bad == 0
print("dev_fuse_ln:", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
