"""Developer smoke script (not a pytest file): first-contact checks of every kernel on a B200.
Run:  gpurun -- python tests/dev_gpu_check.py
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from freepose_b200 import ops  # noqa: E402
from freepose_b200._lib import FP_EPI_BIAS, FP_EPI_BIAS_GELU, FP_EPI_BIAS_LS_RES, FP_EPI_PATCH_EMBED  # noqa: E402
from oracle.vit import contract_attention, contract_layernorm, rb  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
bf = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm()).item(), (a - b).abs().max().item()


def run(name, fn):
    try:
        t0 = time.time()
        fn()
        torch.cuda.synchronize()
        print(f"[ok] {name} ({time.time() - t0:.2f}s)", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"[FAIL] {name}: {type(e).__name__}: {e}", flush=True)


def t_gemm_bias():
    for M in (128, 1000, 5000):
        a = torch.randn(M, 1024, device=dev).to(bf)
        w = (torch.randn(3072, 1024, device=dev) / 32).to(bf)
        b = torch.randn(3072, device=dev).to(bf)
        out = ops.gemm(a, w, b, FP_EPI_BIAS)
        ref = rb(a.float() @ w.float().t() + b.float())
        print("  gemm bias M=%d rel/max" % M, rel(out, ref), "mismatch frac", (out.float() != ref).float().mean().item())


def t_gemm_gelu():
    M = 777
    a = torch.randn(M, 1024, device=dev).to(bf)
    w = (torch.randn(4096, 1024, device=dev) / 32).to(bf)
    b = torch.randn(4096, device=dev).to(bf)
    out = ops.gemm(a, w, b, FP_EPI_BIAS_GELU)
    ref = rb(torch.nn.functional.gelu(rb(a.float() @ w.float().t() + b.float())))
    print("  gemm gelu rel/max", rel(out, ref), "mismatch frac", (out.float() != ref).float().mean().item())


def t_gemm_res():
    M = 900
    a = torch.randn(M, 4096, device=dev).to(bf)
    w = (torch.randn(1024, 4096, device=dev) / 64).to(bf)
    b = torch.randn(1024, device=dev).to(bf)
    g = torch.rand(1024, device=dev).to(bf)
    x = torch.randn(M, 1024, device=dev).to(bf)
    ref = rb(x.float() + rb(rb(a.float() @ w.float().t() + b.float()) * g.float()))
    out = ops.gemm(a, w, b, FP_EPI_BIAS_LS_RES, gamma=g, residual=x.clone())
    print("  gemm ls+res rel/max", rel(out, ref), "mismatch frac", (out.float() != ref).float().mean().item())


def t_gemm_patch():
    B, P, T = 3, 256, 261
    a = torch.zeros(B * P, 640, device=dev, dtype=bf)
    a[:, :588] = torch.randn(B * P, 588, device=dev).to(bf)
    w = torch.zeros(1024, 640, device=dev, dtype=bf)
    w[:, :588] = (torch.randn(1024, 588, device=dev) / 24).to(bf)
    b = torch.randn(1024, device=dev).to(bf)
    pos = torch.randn(1 + P, 1024, device=dev).to(bf)
    out = torch.zeros(B * T, 1024, device=dev, dtype=bf)
    ops.gemm(a, w, b, FP_EPI_PATCH_EMBED, out=out, pos=pos, patches_per_img=P, tokens_per_img=T, token_offset=5)
    ref = rb(rb(a.float() @ w.float().t() + b.float()).view(B, P, 1024) + pos[1:].float())
    got = out.view(B, T, 1024)[:, 5:]
    print("  gemm patch rel/max", rel(got, ref), "special rows untouched:", out.view(B, T, 1024)[:, :5].abs().max().item())


def t_layernorm():
    x = (torch.randn(1000, 1024, device=dev) * 2 + 0.3).to(bf)
    w = (1 + 0.1 * torch.randn(1024, device=dev)).to(bf)
    b = (0.1 * torch.randn(1024, device=dev)).to(bf)
    out = ops.layernorm(x, w, b)
    ref = contract_layernorm(x.float(), w.float(), b.float(), 1e-6)
    print("  layernorm rel/max", rel(out, ref), "mismatch frac", (out.float() != ref).float().mean().item())


def t_attention():
    for B, T in ((2, 261), (1, 100), (3, 272), (2, 256)):
        qkv = torch.randn(B * T, 3072, device=dev).to(bf)
        out = ops.attention(qkv, B, T)
        q, k, v = qkv.float().view(B, T, 3, 16, 64).permute(2, 0, 3, 1, 4)
        ref = contract_attention(q, k, v, 0.125).transpose(1, 2).reshape(B * T, 1024)
        print("  attention B=%d T=%d rel/max" % (B, T), rel(out, ref), "mismatch frac",
              (out.float() != ref).float().mean().item())


def t_score():
    B, P, D = 37, 256, 1024
    ft = torch.randn(B, P, D, device=dev).to(bf)
    fq = torch.randn(P, D, device=dev).to(bf)
    s, idx, vals, _ = ops.score_topk(ft, fq, 3)
    tn = torch.nn.functional.normalize(ft, dim=-1)
    qn = torch.nn.functional.normalize(fq[None], dim=-1)
    ref = (tn * qn).sum(-1, dtype=torch.float32).to(bf).float().mean(-1).to(bf).float()
    print("  score max abs diff", (s - ref).abs().max().item(), "idx", idx.tolist(), torch.topk(ref, 3).indices.tolist())


def t_big_gemm():
    M = 135720
    a = torch.randn(M, 1024, device=dev).to(bf)
    w = (torch.randn(3072, 1024, device=dev) / 32).to(bf)
    b = torch.randn(3072, device=dev).to(bf)
    out = torch.empty(M, 3072, device=dev, dtype=bf)
    for _ in range(2):
        ops.gemm(a, w, b, FP_EPI_BIAS, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.gemm(a, w, b, FP_EPI_BIAS, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("  qkv gemm %.3f ms  %.1f TFLOP/s" % (ms, 2 * M * 1024 * 3072 / ms / 1e9))
    e0.record()
    for _ in range(5):
        ref = torch.nn.functional.linear(a, w, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("  cublas    %.3f ms  %.1f TFLOP/s" % (ms, 2 * M * 1024 * 3072 / ms / 1e9))
    print("  vs cublas rel", rel(out[:4096], ref[:4096]))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    run("gemm bias", t_gemm_bias)
    run("gemm gelu", t_gemm_gelu)
    run("gemm ls+res", t_gemm_res)
    run("gemm patch", t_gemm_patch)
    run("layernorm", t_layernorm)
    run("attention", t_attention)
    run("score", t_score)
    run("big gemm", t_big_gemm)
