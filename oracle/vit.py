"""ORACLE (test infrastructure, never shipped / never on the product path).

CPU restatement of the DINOv2 ViT-L/14-reg forward that the reference reaches through
``torch.hub.load('facebookresearch/dinov2', 'dinov2_vitl14_reg')``
(reference ``src/pipeline/retrieval/dino.py:10``).  The hub source is NOT under /root/reference
(third-party, unpinned); the published architecture is restated here and pinned against the
independent implementation that *is* installed: ``transformers.Dinov2WithRegistersModel``
(tests/test_oracle_vit.py).  PARITY NOTE: no golden vectors exist in the reference for this path
 -- the ViT oracle is pinned by agreement of two independent
implementations, not by reference fixtures.

The module is *hub shaped*: it exposes exactly the attributes the reference touches
(``prepare_tokens_with_masks``, ``blocks``, ``norm``, ``num_register_tokens`` --
``dino.py:16-30``), so the reference's own ``DINOv2FeatureExtractor.forward`` runs on it unmodified
(oracle/refimport.py).

Three arithmetic modes:

* ``OracleViT(sd).float()``                  plain fp32 (architecture check vs transformers);
* ``OracleViT(sd).to(torch.bfloat16)``       native PyTorch-eager bf16 (what the reference runs:
                                              every ATen op rounds its output to bf16);
* ``OracleViT(sd, contract=True)``           the explicit *rounding contract* the CUDA engine
                                              implements: fp32 math on bf16-valued tensors with a
                                              round-to-bf16 at exactly the points where eager bf16
                                              rounds, and flash/xformers-style attention (logits and
                                              softmax statistics in fp32, un-normalised P rounded to
                                              bf16 before P.V, division by the fp32 row sum last) --
                                              the reference environment installs xformers
                                              (``environment_cuda.yaml:33``), whose fused kernel
                                              never materialises bf16 logits.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from freepose_b200.vit_weights import VITL14_REG, VitConfig, interpolated_pos_embed, state_dict_depth


# Accumulation dtype of the contract mode.  float32 is the contract; float64 exists only so that a test can measure how
# far two realisations of the SAME rounding contract drift apart through accumulation order alone
# (tests/test_oracle_vit.py::test_contract_drift_fp32_vs_fp64_accumulation).
ACC_DTYPE = torch.float32


class accumulate_in:
    """``with accumulate_in(torch.float64): ...`` -- contract-mode math in that dtype (rounding points unchanged)."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global ACC_DTYPE
        self.prev, ACC_DTYPE = ACC_DTYPE, self.dtype

    def __exit__(self, *exc):
        global ACC_DTYPE
        ACC_DTYPE = self.prev


def rb(x: torch.Tensor) -> torch.Tensor:
    """Round to the nearest bf16 value (ties to even) and return it in the accumulation dtype (fp32)."""
    return x.to(torch.bfloat16).to(ACC_DTYPE)


class _LayerScale(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class _Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.scale = (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        o = F.scaled_dot_product_attention(q, k, v, scale=self.scale)
        return self.proj(o.transpose(1, 2).reshape(B, N, C))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, cfg: VitConfig, contract: bool):
        super().__init__()
        D = cfg.embed_dim
        self.cfg = cfg
        self.contract = contract
        self.norm1 = nn.LayerNorm(D, eps=cfg.ln_eps)
        self.attn = _Attention(D, cfg.num_heads)
        self.ls1 = _LayerScale(D)
        self.norm2 = nn.LayerNorm(D, eps=cfg.ln_eps)
        self.mlp = _Mlp(D, cfg.mlp_dim)
        self.ls2 = _LayerScale(D)

    def forward(self, x):
        if self.contract:
            return self._forward_contract(x)
        x = x + self.ls1(self.attn(self.norm1(x)))
        x = x + self.ls2(self.mlp(self.norm2(x)))
        return x

    # ---- explicit rounding contract (fp32 tensors holding bf16 values) -------------------
    def _forward_contract(self, x):
        cfg = self.cfg
        f = lambda p: p.detach().to(ACC_DTYPE)
        B, N, D = x.shape
        Hn, hd = cfg.num_heads, cfg.head_dim
        h = contract_layernorm(x, f(self.norm1.weight), f(self.norm1.bias), cfg.ln_eps)
        qkv = rb(h @ f(self.attn.qkv.weight).t() + f(self.attn.qkv.bias))
        qkv = qkv.reshape(B, N, 3, Hn, hd).permute(2, 0, 3, 1, 4)
        o = contract_attention(qkv[0], qkv[1], qkv[2], self.attn.scale)
        o = o.transpose(1, 2).reshape(B, N, D)
        y = rb(o @ f(self.attn.proj.weight).t() + f(self.attn.proj.bias))
        x = rb(x + rb(y * f(self.ls1.gamma)))
        h = contract_layernorm(x, f(self.norm2.weight), f(self.norm2.bias), cfg.ln_eps)
        u = rb(h @ f(self.mlp.fc1.weight).t() + f(self.mlp.fc1.bias))
        a = rb(F.gelu(u))
        y = rb(a @ f(self.mlp.fc2.weight).t() + f(self.mlp.fc2.bias))
        x = rb(x + rb(y * f(self.ls2.gamma)))
        return x


def contract_layernorm(x, w, b, eps):
    """fp32 two-pass statistics, y = ((x - mean) * rstd) * w + b, one rounding to bf16."""
    mean = x.mean(dim=-1, keepdim=True)
    xc = x - mean
    var = (xc * xc).mean(dim=-1, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    return rb((xc * rstd) * w + b)


SINGLE_PASS_KEYS = 272   # padded keys the single-pass kernel holds in TMEM (freepose_b200/csrc/attention.cu)
KEY_BLOCK = 96           # key block of the tiled-key kernel (freepose_b200/csrc/attention_pair.cu)


SPLIT_TOKENS = 261       # 224^2 crops: the two-stream kernel (freepose_b200/csrc/attention_split.cu)
SPLIT_KEYS = 144         # stream 0 = keys [0, 144), stream 1 = keys [144, N)


def contract_attention_split(q, k, v, scale, split=SPLIT_KEYS, tile=128):
    """Split-key flash attention, the arithmetic of attention_split.cu: the keys of every full 128-row query tile are
    dealt to two independent streams, each with its own row max m_h, un-normalised bf16 P_h = bf16(exp(s - m_h)), fp32
    row sum l_h of the unrounded p and accumulator O_h = P_h V_h; the tile ends with
    O = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1), a_h = exp(m_h - max(m_0, m_1)), rounded once.  (The same as two key
    blocks of an online softmax.)  The N % 128 leftover query rows (cls + registers come first, so these are the last
    patch rows) are computed by CUDA-core warps in ONE pass over all keys."""
    N = k.shape[-2]
    full = N // tile * tile
    s = (q @ k.transpose(-2, -1)) * scale
    out = torch.empty_like(q)
    # full tiles: two streams
    parts = []
    for lo, hi in ((0, split), (split, N)):
        sh = s[..., :full, lo:hi]
        m = sh.amax(dim=-1, keepdim=True)
        p = torch.exp(sh - m)
        parts.append((m, p.sum(dim=-1, keepdim=True), rb(p) @ v[..., lo:hi, :]))
    (m0, l0, o0), (m1, l1, o1) = parts
    mm = torch.maximum(m0, m1)
    a0, a1 = torch.exp(m0 - mm), torch.exp(m1 - mm)
    out[..., :full, :] = (o0 * a0 + o1 * a1) / (l0 * a0 + l1 * a1)
    # leftover rows: single pass
    st = s[..., full:, :]
    mt = st.amax(dim=-1, keepdim=True)
    pt = torch.exp(st - mt)
    out[..., full:, :] = (rb(pt) @ v) / pt.sum(dim=-1, keepdim=True)
    return rb(out)


LAZY_TAU = 8.0           # attention_pair.cu: log2 of the largest un-normalised P before a row's reference max moves


def contract_attention(q, k, v, scale, key_block=None, lazy_tau=LAZY_TAU):
    """Flash/xformers-style attention on bf16-valued fp32 tensors (B, H, N, hd).

    261 tokens (224^2 crops, the headline shape) run the two-stream form (contract_attention_split).  Otherwise, up to
    SINGLE_PASS_KEYS keys the softmax is a single pass; above that (crops larger than 224^2) it is the
    block-wise online softmax of flash attention with KEY_BLOCK keys per block: the bf16 P of block b is taken
    against the row's reference max after block b (the running max, moved lazily: FlashAttention-4's rule), and O / l
    are rescaled by exp(m_old - m_new) when it moves.
    """
    N = k.shape[-2]
    if key_block is None and N == SPLIT_TOKENS:
        return contract_attention_split(q, k, v, scale)
    if key_block is None:
        key_block = KEY_BLOCK if (N + 15) // 16 * 16 > SINGLE_PASS_KEYS else 0
    if not key_block or N <= key_block:
        s = (q @ k.transpose(-2, -1)) * scale            # fp32 logits, never rounded
        m = s.amax(dim=-1, keepdim=True)
        p = torch.exp(s - m)                              # fp32
        l = p.sum(dim=-1, keepdim=True)                   # row sum of the UNROUNDED p
        o = (rb(p) @ v) / l                               # P rounded to bf16 for the tensor-core P.V
        return rb(o)
    m = torch.full(q.shape[:-1] + (1,), float("-inf"), dtype=q.dtype)
    l = torch.zeros_like(m)
    o = torch.zeros_like(q)
    log2e = 1.4426950408889634
    for k0 in range(0, N, key_block):
        s = (q @ k[..., k0:k0 + key_block, :].transpose(-2, -1)) * scale
        blk = s.amax(dim=-1, keepdim=True)
        # lazy reference maximum (attention_pair.cu LAZY_TAU): it moves only when the block maximum exceeds it by more
        # than 2^lazy_tau in the exponent; lazy_tau = 0 is the classic running maximum
        grow = (blk - m) * log2e > lazy_tau
        m_new = torch.where(grow, blk, m)
        alpha = torch.where(grow, torch.exp(m - m_new), torch.ones_like(m))   # 0 for the first block
        p = torch.exp(s - m_new)
        l = l * alpha + p.sum(dim=-1, keepdim=True)
        o = o * alpha + rb(p) @ v[..., k0:k0 + key_block, :]
        m = m_new
    return rb(o / l)


class _PatchEmbed(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.proj = nn.Conv2d(3, cfg.embed_dim, kernel_size=cfg.patch_size, stride=cfg.patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class OracleViT(nn.Module):
    """Hub-shaped DINOv2-with-registers ViT (see module docstring)."""

    def __init__(self, state_dict: dict, cfg: VitConfig = VITL14_REG, contract: bool = False):
        super().__init__()
        self.cfg = cfg
        self.contract = contract
        depth = state_dict_depth(state_dict)
        D = cfg.embed_dim
        self.patch_size = cfg.patch_size
        self.num_register_tokens = cfg.num_register_tokens
        self.patch_embed = _PatchEmbed(cfg)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, D))
        self.pos_embed = nn.Parameter(torch.zeros(1, cfg.pos_grid ** 2 + 1, D))
        self.register_tokens = nn.Parameter(torch.zeros(1, cfg.num_register_tokens, D))
        self.mask_token = nn.Parameter(torch.zeros(1, D))
        self.blocks = nn.ModuleList([_Block(cfg, contract) for _ in range(depth)])
        self.norm = _ContractableLayerNorm(D, cfg.ln_eps, contract)
        self.load_state_dict({k: v.float() for k, v in state_dict.items()}, strict=True)
        self.eval()
        self._sd_dtype = state_dict["cls_token"].dtype

    def interpolate_pos_encoding(self, x, w, h):
        assert w == h, "square crops only on this path (reference crops are square)"
        sd = {"pos_embed": self.pos_embed.detach()}
        return interpolated_pos_embed(sd, self.cfg, w).unsqueeze(0)

    def prepare_tokens_with_masks(self, x, masks=None):
        assert masks is None
        B, nc, w, h = x.shape
        if self.contract:
            return self._prepare_contract(x)
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
        x = x + self.interpolate_pos_encoding(x, w, h).to(x.dtype)
        x = torch.cat((x[:, :1], self.register_tokens.expand(B, -1, -1), x[:, 1:]), dim=1)
        return x

    def _prepare_contract(self, x):
        """x: normalised image, bf16 values (any float dtype).  Returns fp32 holding bf16 values."""
        B, nc, w, h = x.shape
        f = lambda p: p.detach().to(ACC_DTYPE)
        x = rb(x.to(ACC_DTYPE))
        t = rb(F.conv2d(x, f(self.patch_embed.proj.weight), f(self.patch_embed.proj.bias),
                        stride=self.cfg.patch_size).flatten(2).transpose(1, 2))
        t = torch.cat((f(self.cls_token).expand(B, -1, -1), t), dim=1)
        # pos-embed: fp32 bicubic of the (bf16-valued) parameter, rounded to the model dtype (bf16)
        pos = rb(interpolated_pos_embed({"pos_embed": f(self.pos_embed)}, self.cfg, w)).unsqueeze(0)
        t = rb(t + pos)
        t = torch.cat((t[:, :1], f(self.register_tokens).expand(B, -1, -1), t[:, 1:]), dim=1)
        return t

    def forward_features(self, x, layer: int):
        """Tokens after `layer` blocks and the final norm -- restates reference dino.py:16-23."""
        x = self.prepare_tokens_with_masks(x, None)
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if i + 1 == layer:
                break
        return self.norm(x)


class _ContractableLayerNorm(nn.LayerNorm):
    def __init__(self, dim, eps, contract):
        super().__init__(dim, eps=eps)
        self.contract = contract

    def forward(self, x):
        if self.contract:
            return contract_layernorm(x, self.weight.detach().to(ACC_DTYPE), self.bias.detach().to(ACC_DTYPE), self.eps)
        return super().forward(x)


def to_hf_state_dict(sd: dict, cfg: VitConfig = VITL14_REG) -> dict:
    """Map hub-format keys onto transformers' ``Dinov2WithRegistersModel`` (for the cross-check)."""
    D = cfg.embed_dim
    out = {
        "embeddings.cls_token": sd["cls_token"],
        "embeddings.mask_token": sd["mask_token"],
        "embeddings.register_tokens": sd["register_tokens"],
        "embeddings.position_embeddings": sd["pos_embed"],
        "embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
        "embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
        "layernorm.weight": sd["norm.weight"],
        "layernorm.bias": sd["norm.bias"],
    }
    for i in range(state_dict_depth(sd)):
        p, q = f"blocks.{i}.", f"encoder.layer.{i}."
        w, b = sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]
        for j, name in enumerate(("query", "key", "value")):
            out[q + f"attention.attention.{name}.weight"] = w[j * D:(j + 1) * D]
            out[q + f"attention.attention.{name}.bias"] = b[j * D:(j + 1) * D]
        out[q + "attention.output.dense.weight"] = sd[p + "attn.proj.weight"]
        out[q + "attention.output.dense.bias"] = sd[p + "attn.proj.bias"]
        out[q + "layer_scale1.lambda1"] = sd[p + "ls1.gamma"]
        out[q + "layer_scale2.lambda1"] = sd[p + "ls2.gamma"]
        out[q + "norm1.weight"] = sd[p + "norm1.weight"]
        out[q + "norm1.bias"] = sd[p + "norm1.bias"]
        out[q + "norm2.weight"] = sd[p + "norm2.weight"]
        out[q + "norm2.bias"] = sd[p + "norm2.bias"]
        out[q + "mlp.fc1.weight"] = sd[p + "mlp.fc1.weight"]
        out[q + "mlp.fc1.bias"] = sd[p + "mlp.fc1.bias"]
        out[q + "mlp.fc2.weight"] = sd[p + "mlp.fc2.weight"]
        out[q + "mlp.fc2.bias"] = sd[p + "mlp.fc2.bias"]
    return out
