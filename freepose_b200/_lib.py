"""ctypes binding of ``libfreepose_b200.so`` (include/freepose_b200.h).

The product path has NO fallback: if the library is missing or a call fails, a ``RuntimeError`` is
raised.  Tensors stay owned by PyTorch; raw ``data_ptr()`` values are borrowed for the call.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libfreepose_b200.so"

FP_INPUT_IMAGE_F32, FP_INPUT_IMAGE_BF16, FP_INPUT_PATCHES = 0, 1, 2
FP_FEATURE_ALL, FP_FEATURE_CLS, FP_FEATURE_REG, FP_FEATURE_PATCH = 0, 1, 2, 3
FP_EPI_BIAS, FP_EPI_BIAS_GELU, FP_EPI_BIAS_LS_RES, FP_EPI_PATCH_EMBED = 0, 1, 2, 3
KPAD = 640


class VitLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ls1",
        "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ls2")]


class VitWeights(C.Structure):
    _fields_ = [("depth", C.c_int), ("layers", C.POINTER(VitLayer)), ("patch_w", C.c_void_p),
                ("patch_b", C.c_void_p), ("norm_w", C.c_void_p), ("norm_b", C.c_void_p),
                ("pos_res", C.c_int), ("pos_embed", C.c_void_p), ("special_tokens", C.c_void_p),
                ("dim", C.c_int), ("heads", C.c_int), ("mlp_dim", C.c_int)]


class RasterArgs(C.Structure):
    _fields_ = [("verts", C.c_void_p), ("faces", C.c_void_p), ("colors", C.c_void_p), ("V", C.c_int),
                ("F", C.c_int), ("poses", C.c_void_p), ("B", C.c_int), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("res", C.c_int), ("msaa", C.c_int),
                ("cull_backfaces", C.c_int), ("gamma_lut", C.c_void_p), ("rgb", C.c_void_p),
                ("depth", C.c_void_p), ("primitive", C.c_int), ("uv", C.c_void_p), ("texture", C.c_void_p),
                ("tex_w", C.c_int), ("tex_h", C.c_int), ("tex_levels", C.c_int), ("srgb_lut", C.c_void_p),
                ("ambient", C.c_float), ("znear", C.c_float), ("zfar", C.c_float), ("view_k", C.c_void_p)]


_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "fp_abi_version": (_i, []),
    "fp_last_error": (C.c_char_p, []),
    "fp_device_sm_count": (_i, []),
    "fp_launch_count": (C.c_longlong, []),
    "fp_profile_enable": (None, [_i]),
    "fp_profile_reset": (None, []),
    "fp_profile_num_kinds": (_i, []),
    "fp_profile_kind_name": (C.c_char_p, [_i]),
    "fp_profile_collect": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "fp_vit_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "fp_vit_forward": (_i, [C.POINTER(VitWeights), _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "fp_gemm_bf16": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "fp_layernorm_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _i, _i, _i, _vp]),
    "fp_attention_bf16": (_i, [_vp, _vp, _i, _i, _i, _f, _vp]),
    "fp_im2col_patches": (_i, [_vp, _i, _vp, _i, _i, _i, _vp]),
    "fp_normalize_image": (_i, [_vp, _vp, _i, _i, _vp]),
    "fp_score_workspace_bytes": (_sz, [_i, _i, _i]),
    "fp_score_topk": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "fp_topk": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "fp_ffa_pool": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "fp_normalize_rows": (_i, [_vp, _i, C.c_int64, _i, _vp, _vp]),
    "fp_retrieval_scan": (_i, [_vp, _vp, C.c_int64, _i, _i, _vp, _vp]),
    "fp_topk_rows": (_i, [_vp, _i, C.c_int64, _i, _vp, _vp, _vp]),
    "fp_retrieval_fine": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "fp_softvote_add": (_i, [_vp, _vp, _vp, _i, _i, C.c_int64, _vp]),
    "fp_softvote_mean": (_i, [_vp, _vp, C.c_int64, _i, _vp]),
    "fp_raster_workspace_bytes": (_i, [_i, _i, _i, _i, _i, C.POINTER(_sz)]),
    "fp_rasterize": (_i, [C.POINTER(RasterArgs), _vp, _sz, _vp]),
    "fp_mask_bbox": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "fp_crop_resize_pad": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "fp_depth_extents": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "fp_roi_align": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "fp_depth_mask_cubic": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "fp_patch_cosine": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "fp_comm_unique_id": (_i, [_vp]),
    "fp_comm_create": (_i, [_vp, _i, _i, C.POINTER(_vp)]),
    "fp_allgather_scores": (_i, [_vp, _vp, _i, _vp]),
    "fp_comm_destroy": (_i, [_vp]),
    "fp_exchange_bytes": (_sz, [_i, _i]),
    "fp_p2p_alloc": (_i, [_sz, C.POINTER(_vp), _vp]),
    "fp_p2p_open": (_i, [_vp, C.POINTER(_vp)]),
    "fp_p2p_close": (_i, [_vp]),
    "fp_p2p_free": (_i, [_vp]),
    "fp_score_publish": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, C.c_uint, _vp, _sz, _vp]),
    "fp_topk_after_exchange": (_i, [_vp, _i, _i, _i, C.c_uint, _i, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the C-ABI library (building is an explicit step: ``python -m freepose_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("FREEPOSE_B200_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m freepose_b200.build`.")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.fp_abi_version() != 3:
        raise RuntimeError("libfreepose_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().fp_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def ptr(t: torch.Tensor | None):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("freepose_b200 kernels take CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("freepose_b200 kernels take contiguous tensors")
    if t.device.index != torch.cuda.current_device():
        # kernels launch on torch.cuda.current_stream(): a tensor of another GPU would be dereferenced on the wrong device
        raise RuntimeError(f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                           "wrap the call in `with torch.cuda.device(tensor.device)`")
    return t.data_ptr()


def on_device(method):
    """Decorator for methods of objects with a ``.device``: runs the call with that GPU current, so streams, launches
    and per-device kernel attributes all refer to the device the object's tensors live on."""
    import functools

    @functools.wraps(method)
    def wrapped(self, *args, **kwargs):
        # the common case -- the object's GPU is already current -- must cost nothing: torch.cuda.device(torch.device)
        # re-resolves the device type (a driver query of ~0.1 ms) on every entry, 70 times per video frame
        idx = self.device.index
        if idx is None or idx == torch.cuda.current_device():
            return method(self, *args, **kwargs)
        with torch.cuda.device(idx):
            return method(self, *args, **kwargs)
    return wrapped


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the current device."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream
