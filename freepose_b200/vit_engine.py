"""Device-resident DINOv2 ViT-L/14-reg (or ViT-B/14-reg) engine: packs a hub-format state dict for the C ABI and drives
``fp_vit_forward`` in image chunks over one reusable workspace."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (FP_FEATURE_ALL, FP_FEATURE_CLS, FP_FEATURE_PATCH, FP_FEATURE_REG, FP_INPUT_IMAGE_BF16,
                   FP_INPUT_IMAGE_F32, FP_INPUT_PATCHES, KPAD, check, load, on_device, ptr, stream_ptr)
from .vit_weights import VITL14_REG, VitConfig, interpolated_pos_embed, state_dict_depth

bf16 = torch.bfloat16
FEATURE_TYPES = {"all": FP_FEATURE_ALL, "cls": FP_FEATURE_CLS, "reg": FP_FEATURE_REG, "patch": FP_FEATURE_PATCH}


class ViTEngine:
    """Holds the bf16 parameters in HBM (554 MB for 22 blocks) and runs the forward on the current stream.

    ``chunk`` images are processed per ``fp_vit_forward`` call; the workspace (residual stream, qkv, MLP
    hidden, patch matrix: ~5.2 MB per image at 224^2) is allocated once for ``chunk`` and reused.
    """

    def __init__(self, state_dict: dict, cfg: VitConfig = VITL14_REG, device="cuda", chunk: int = 256):
        if not torch.cuda.is_available():
            raise RuntimeError("ViTEngine needs a CUDA device (no CPU fallback)")
        assert cfg.embed_dim in (1024, 768) and cfg.head_dim == 64 and cfg.mlp_dim % 256 == 0 and cfg.patch_size == 14, \
            "the sm_100a kernels cover ViT-L/14 and ViT-B/14 (head dim 64)"
        self.cfg = cfg
        self.device = torch.device(device)
        self.depth = state_dict_depth(state_dict)
        self.chunk = chunk
        self._lib = load()
        sd = {k: v.detach().to(bf16) for k, v in state_dict.items()}
        self._host_sd = {"pos_embed": sd["pos_embed"], "cls_token": sd["cls_token"],
                         "register_tokens": sd["register_tokens"]}
        dev = self.device
        keep = []  # device tensors referenced by raw pointers

        def put(t):
            t = t.to(dev).contiguous()
            keep.append(t)
            return ptr(t)

        pw = torch.zeros(cfg.embed_dim, KPAD, dtype=bf16)
        pw[:, :588] = sd["patch_embed.proj.weight"].reshape(cfg.embed_dim, 588)
        self._layers = (_lib.VitLayer * self.depth)()
        for i in range(self.depth):
            p = f"blocks.{i}."
            L = self._layers[i]
            L.ln1_w, L.ln1_b = put(sd[p + "norm1.weight"]), put(sd[p + "norm1.bias"])
            L.qkv_w, L.qkv_b = put(sd[p + "attn.qkv.weight"]), put(sd[p + "attn.qkv.bias"])
            L.proj_w, L.proj_b = put(sd[p + "attn.proj.weight"]), put(sd[p + "attn.proj.bias"])
            L.ls1 = put(sd[p + "ls1.gamma"])
            L.ln2_w, L.ln2_b = put(sd[p + "norm2.weight"]), put(sd[p + "norm2.bias"])
            L.fc1_w, L.fc1_b = put(sd[p + "mlp.fc1.weight"]), put(sd[p + "mlp.fc1.bias"])
            L.fc2_w, L.fc2_b = put(sd[p + "mlp.fc2.weight"]), put(sd[p + "mlp.fc2.bias"])
            L.ls2 = put(sd[p + "ls2.gamma"])
        self._w = _lib.VitWeights()
        self._w.depth = self.depth
        self._w.dim, self._w.heads, self._w.mlp_dim = cfg.embed_dim, cfg.num_heads, cfg.mlp_dim
        self._w.layers = C.cast(self._layers, C.POINTER(_lib.VitLayer))
        self._w.patch_w, self._w.patch_b = put(pw), put(sd["patch_embed.proj.bias"])
        self._w.norm_w, self._w.norm_b = put(sd["norm.weight"]), put(sd["norm.bias"])
        self._keep = keep
        self._pos_cache = {}
        self._ws = None
        self._ws_key = None

    # -- per-resolution constants: resampled pos-embed and the 5 special token rows ---------------
    def _prepare_res(self, res: int):
        if res not in self._pos_cache:
            cfg = self.cfg
            pos = interpolated_pos_embed(self._host_sd, cfg, res)                     # (1+g*g, D) bf16
            cls = (self._host_sd["cls_token"][0, 0].float() + pos[0].float()).to(bf16)  # bf16(cls + pos[0])
            special = torch.cat((cls[None], self._host_sd["register_tokens"][0]), dim=0)
            self._pos_cache[res] = (pos.to(self.device).contiguous(), special.to(self.device).contiguous())
        pos, special = self._pos_cache[res]
        self._w.pos_res = res
        self._w.pos_embed = ptr(pos)
        self._w.special_tokens = ptr(special)

    def _workspace(self, batch: int, res: int) -> torch.Tensor:
        key = (batch, res)
        need = self._lib.fp_vit_workspace_bytes(self.cfg.embed_dim, self.cfg.mlp_dim, batch, res)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        self._ws_key = key
        return self._ws

    def tokens_per_image(self, res: int, feature_type: str) -> int:
        g = res // self.cfg.patch_size
        return {"all": g * g + 5, "cls": 1, "reg": 4, "patch": g * g}[feature_type]

    @on_device
    def forward(self, x: torch.Tensor, layer: int = 22, feature_type: str = "patch", res: int | None = None,
                out: torch.Tensor | None = None) -> torch.Tensor:
        """x: (B,3,res,res) fp32 in [0,1] (Normalize applied on device), (B,3,res,res) bf16 already normalised,
        or a (B*g*g, 640) bf16 patch matrix (then ``res`` is required).  Returns bf16 tokens."""
        if x.dim() == 4:
            B, res = x.shape[0], x.shape[-1]
            kind = FP_INPUT_IMAGE_F32 if x.dtype == torch.float32 else FP_INPUT_IMAGE_BF16
            assert x.dtype in (torch.float32, bf16) and x.shape[1] == 3 and x.shape[2] == x.shape[3]
            per_img = 3 * res * res
        else:
            assert res is not None and x.dtype == bf16 and x.shape[1] == KPAD
            g = res // 14
            B = x.shape[0] // (g * g)
            kind = FP_INPUT_PATCHES
            per_img = g * g * KPAD
        if layer > self.depth:
            raise ValueError(f"layer {layer} requested but only {self.depth} blocks are loaded")
        x = x.to(self.device).contiguous()
        ft = FEATURE_TYPES[feature_type]
        n_tok = self.tokens_per_image(res, feature_type)
        D = self.cfg.embed_dim
        if out is None:
            out = torch.empty(B, n_tok, D, dtype=bf16, device=self.device)
        self._prepare_res(res)
        chunk = min(self.chunk, max(B, 1))
        ws = self._workspace(chunk, res)
        flat = x.reshape(-1)
        for b0 in range(0, B, chunk):
            nb = min(chunk, B - b0)
            src = flat[b0 * per_img:(b0 + nb) * per_img]
            dst = out[b0:b0 + nb]
            check(self._lib.fp_vit_forward(C.byref(self._w), src.data_ptr(), kind, nb, res, layer, ft, dst.data_ptr(),
                                           ws.data_ptr(), ws.numel(), stream_ptr()), "fp_vit_forward")
        if feature_type == "cls":
            return out[:, 0]
        return out
