"""DINOv2 ViT-L/14-reg weight containers.

The reference loads ``dinov2_vitl14_reg`` through ``torch.hub`` (reference
``src/pipeline/retrieval/dino.py:10``).  That needs network access, so this module offers

* :func:`synthetic_state_dict` -- a seeded state dict with exactly the hub checkpoint's key names
  and shapes (``dinov2_vitl14_reg4_pretrain.pth``), scaled so that every residual branch
  contributes O(1) signal (parity tests against such weights are not vacuous);
* :func:`load_state_dict_file` -- reads a real hub checkpoint when the user has one.

Both produce the same ``dict[str, torch.Tensor]`` which the CUDA engine packs
(:mod:`freepose_b200.vit_engine`) and which ``oracle/vit.py`` consumes unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class VitConfig:
    """Architecture constants of ``dinov2_vitl14_reg``."""

    embed_dim: int = 1024
    depth: int = 24
    num_heads: int = 16
    mlp_dim: int = 4096
    patch_size: int = 14
    num_register_tokens: int = 4
    pos_grid: int = 37  # 518 / 14: the checkpoint's native position-embedding grid
    ln_eps: float = 1e-6

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    def num_tokens(self, res: int) -> int:
        g = res // self.patch_size
        return g * g + 1 + self.num_register_tokens


VITL14_REG = VitConfig()
# dinov2_vitb14_reg: the refiner's confidence model (reference tracking_refiner.py:20-23)
VITB14_REG = VitConfig(embed_dim=768, depth=12, num_heads=12, mlp_dim=3072)


def synthetic_state_dict(cfg: VitConfig = VITL14_REG, seed: int = 0, depth: int | None = None,
                         dtype: torch.dtype = torch.bfloat16) -> dict:
    """Seeded stand-in for the hub checkpoint (same keys / shapes).

    Scales: every Linear is N(0, 1/fan_in) ("unit gain"), fc2 gets an extra 0.5, LayerScale
    gamma is uniform in [0.15, 0.45], LayerNorm affine is 1 +- 0.1 / 0 +- 0.05, biases are
    N(0, 0.02^2), patch-embed conv is N(0, 1/588), pos-embed / cls / registers N(0, 0.3^2).
    With these, attention logits have ~unit variance (non-trivial softmax) and the residual
    stream grows from ~1 to ~3 std over 22 blocks -- the regime of a trained network.
    Values are rounded to ``dtype`` (bf16, the reference's model dtype:
    ``pose_estimator.py:21``) so the oracle and the engine read bit-identical parameters.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    D, H = cfg.embed_dim, cfg.mlp_dim
    L = cfg.depth if depth is None else depth

    def normal(shape, std):
        return (torch.randn(shape, generator=g, dtype=torch.float32) * std).to(dtype)

    def uniform(shape, lo, hi):
        return (torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo).to(dtype)

    sd = {}
    k = cfg.patch_size
    sd["cls_token"] = normal((1, 1, D), 0.3)
    sd["pos_embed"] = normal((1, cfg.pos_grid * cfg.pos_grid + 1, D), 0.3)
    sd["register_tokens"] = normal((1, cfg.num_register_tokens, D), 0.3)
    sd["mask_token"] = torch.zeros(1, D, dtype=dtype)
    sd["patch_embed.proj.weight"] = normal((D, 3, k, k), 1.0 / math.sqrt(3 * k * k))
    sd["patch_embed.proj.bias"] = normal((D,), 0.02)
    for i in range(L):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = uniform((D,), 0.9, 1.1)
        sd[p + "norm1.bias"] = normal((D,), 0.05)
        sd[p + "attn.qkv.weight"] = normal((3 * D, D), 1.0 / math.sqrt(D))
        sd[p + "attn.qkv.bias"] = normal((3 * D,), 0.02)
        sd[p + "attn.proj.weight"] = normal((D, D), 1.0 / math.sqrt(D))
        sd[p + "attn.proj.bias"] = normal((D,), 0.02)
        sd[p + "ls1.gamma"] = uniform((D,), 0.15, 0.45)
        sd[p + "norm2.weight"] = uniform((D,), 0.9, 1.1)
        sd[p + "norm2.bias"] = normal((D,), 0.05)
        sd[p + "mlp.fc1.weight"] = normal((H, D), 1.0 / math.sqrt(D))
        sd[p + "mlp.fc1.bias"] = normal((H,), 0.02)
        sd[p + "mlp.fc2.weight"] = normal((D, H), 0.5 / math.sqrt(H))
        sd[p + "mlp.fc2.bias"] = normal((D,), 0.02)
        sd[p + "ls2.gamma"] = uniform((D,), 0.15, 0.45)
    sd["norm.weight"] = uniform((D,), 0.9, 1.1)
    sd["norm.bias"] = normal((D,), 0.05)
    return sd


def load_state_dict_file(path: str, dtype: torch.dtype = torch.bfloat16) -> dict:
    """Read a real ``dinov2_vitl14_reg4_pretrain.pth`` (hub format) into the same dict layout."""
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if "model" in sd and "cls_token" not in sd:
        sd = sd["model"]
    return {k: v.to(dtype) for k, v in sd.items()}


def state_dict_depth(sd: dict) -> int:
    n = 0
    while f"blocks.{n}.norm1.weight" in sd:
        n += 1
    return n


def interpolated_pos_embed(sd: dict, cfg: VitConfig, res: int) -> torch.Tensor:
    """Patch position embedding resampled to the (res/14)^2 grid, shape (1 + g*g, D), in the
    parameter dtype.

    Restates hub ``DinoVisionTransformer.interpolate_pos_encoding`` for ``_reg`` models
    (interpolate_offset = 0, antialias = True, fp32 bicubic) --
    the same arithmetic as transformers' ``Dinov2WithRegistersEmbeddings.interpolate_pos_encoding``.
    This is once-per-resolution host set-up, not hot-path work.
    """
    pos = sd["pos_embed"]
    dt = pos.dtype
    g = res // cfg.patch_size
    M = cfg.pos_grid
    if g == M:
        return pos[0].clone()
    p = pos.float()
    cls_pos = p[:, 0]
    patch = p[:, 1:].reshape(1, M, M, cfg.embed_dim).permute(0, 3, 1, 2)
    patch = torch.nn.functional.interpolate(patch, size=(g, g), mode="bicubic",
                                            align_corners=False, antialias=True)
    patch = patch.permute(0, 2, 3, 1).reshape(1, g * g, cfg.embed_dim)
    return torch.cat((cls_pos.unsqueeze(0), patch), dim=1)[0].to(dt)
