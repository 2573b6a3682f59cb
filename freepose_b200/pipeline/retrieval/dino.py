"""Drop-in for reference ``src/pipeline/retrieval/dino.py:7-32`` backed by the sm_100a ViT engine."""
from __future__ import annotations

import os
import warnings

import torch
import torch.nn as nn

from ...vit_engine import ViTEngine
from ...vit_weights import VITL14_REG, load_state_dict_file, synthetic_state_dict

WEIGHTS_ENV = "FREEPOSE_DINOV2_WEIGHTS"


class DINOv2FeatureExtractor(nn.Module):
    """Same constructor / ``forward(images, layer, feature_type)`` contract as the reference class.

    The reference fetches ``dinov2_vitl14_reg`` through ``torch.hub`` (dino.py:10).  Here the parameters come
    from ``weights`` (a hub-format ``.pth`` path or state dict), else from ``$FREEPOSE_DINOV2_WEIGHTS``, else --
    with a warning -- from the seeded synthetic stand-in (no network in the build environment).
    """

    def __init__(self, model_name: str = "dinov2_vitl14_reg", weights=None, depth: int | None = None,
                 seed: int = 0, chunk: int = 256, device="cuda"):
        super().__init__()
        if model_name != "dinov2_vitl14_reg":
            raise ValueError(f"only dinov2_vitl14_reg is built for sm_100a (got {model_name!r})")
        self.model_name = model_name
        if weights is None and os.environ.get(WEIGHTS_ENV):
            weights = os.environ[WEIGHTS_ENV]
        if isinstance(weights, (str, os.PathLike)):
            sd = load_state_dict_file(os.fspath(weights))
        elif isinstance(weights, dict):
            sd = weights
        else:
            warnings.warn("DINOv2 checkpoint not provided: using seeded SYNTHETIC ViT-L/14-reg weights "
                          f"(set ${WEIGHTS_ENV} or pass weights=...)", stacklevel=2)
            sd = synthetic_state_dict(VITL14_REG, seed=seed, depth=depth)
        self.engine = ViTEngine(sd, VITL14_REG, device=device, chunk=chunk)
        self.num_register_tokens = VITL14_REG.num_register_tokens

    # the reference calls .to('cuda', dtype=torch.bfloat16); parameters already live on the device in bf16
    def to(self, *args, **kwargs):
        return self

    def cuda(self, device=None):
        return self

    @torch.inference_mode()
    def forward(self, images, layer: int = 22, feature_type: str = "cls"):
        """images: (B,3,H,W) float in [0,1] (any float dtype, CPU or CUDA) -> bf16 CUDA features:
        'cls' (B,1024) | 'reg' (B,4,1024) | 'patch' (B,(H/14)^2,1024)   (dino.py:25-30)."""
        if feature_type not in ("cls", "reg", "patch"):
            raise ValueError(f"unknown feature_type {feature_type!r}")
        x = images
        if x.dtype != torch.float32:
            x = x.float()  # exact for bf16/fp16 inputs; the kernel re-rounds to bf16 as the reference's cast does
        x = x.to(self.engine.device, non_blocking=True)
        return self.engine.forward(x, layer=layer, feature_type=feature_type)

    @torch.inference_mode()
    def forward_patches(self, patches, res: int, layer: int = 22, feature_type: str = "patch"):
        """Hot-path entry: normalised bf16 patch matrix straight from the rasteriser/crop kernel."""
        return self.engine.forward(patches, layer=layer, feature_type=feature_type, res=res)
