"""Torch-tensor shims over the C ABI (one function per ``fp_*`` entry point).

Every function enqueues on ``torch.cuda.current_stream()`` and returns device tensors; none of them
synchronises and none has a CPU path.
"""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np
import torch

from . import _lib
from ._lib import (FP_EPI_BIAS, FP_EPI_BIAS_GELU, FP_EPI_BIAS_LS_RES, FP_EPI_PATCH_EMBED, KPAD, check, load, ptr,
                   stream_ptr)

bf16 = torch.bfloat16


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------- ViT stages
def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, mode: int = FP_EPI_BIAS, *, gamma=None,
         residual=None, out=None, pos=None, patches_per_img=0, tokens_per_img=0, token_offset=0):
    """out = epilogue(a @ w.T); a (M,K) bf16, w (N,K) bf16.  See fp_gemm_bf16."""
    M, K = a.shape
    N = w.shape[0]
    assert a.dtype == bf16 and w.dtype == bf16 and w.shape[1] == K
    if out is None:
        out = residual if mode == FP_EPI_BIAS_LS_RES else torch.empty(M, N, dtype=bf16, device=a.device)
    extra = residual if mode == FP_EPI_BIAS_LS_RES else pos
    check(load().fp_gemm_bf16(ptr(a), a.stride(0), ptr(w), ptr(out), out.stride(0), M, N, K, mode, ptr(bias),
                              ptr(gamma), ptr(extra), patches_per_img, tokens_per_img, token_offset,
                              stream_ptr()), "fp_gemm_bf16")
    return out


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-6):
    D = x.shape[-1]
    rows = x.numel() // D
    out = torch.empty_like(x)
    check(load().fp_layernorm_bf16(ptr(x), ptr(w), ptr(b), ptr(out), rows, D, eps, rows, 0, rows, stream_ptr()),
          "fp_layernorm_bf16")
    return out


def attention(qkv: torch.Tensor, batch: int, tokens: int, heads: int = 16, scale: float = 0.125):
    assert qkv.dtype == bf16 and qkv.shape == (batch * tokens, 3 * heads * 64)
    out = torch.empty(batch * tokens, heads * 64, dtype=bf16, device=qkv.device)
    check(load().fp_attention_bf16(ptr(qkv), ptr(out), batch, tokens, heads, scale, stream_ptr()),
          "fp_attention_bf16")
    return out


def im2col(image: torch.Tensor, out: torch.Tensor | None = None):
    B, _, res, _ = image.shape
    g = res // 14
    if out is None:
        out = torch.empty(B * g * g, KPAD, dtype=bf16, device=image.device)
    assert out.dtype == bf16 and out.is_contiguous() and out.shape == (B * g * g, KPAD)
    is_f32 = image.dtype == torch.float32
    assert is_f32 or image.dtype == bf16
    check(load().fp_im2col_patches(ptr(image), int(is_f32), ptr(out), B, res, KPAD, stream_ptr()),
          "fp_im2col_patches")
    return out


def normalize_image(image: torch.Tensor):
    B, _, res, _ = image.shape
    assert image.dtype == torch.float32
    out = torch.empty(image.shape, dtype=bf16, device=image.device)
    check(load().fp_normalize_image(ptr(image), ptr(out), B, res, stream_ptr()), "fp_normalize_image")
    return out


# ------------------------------------------------------------------------------------------- score
def score_topk(feats_t: torch.Tensor, feat_q: torch.Tensor, k: int = 3, weights=None, normalise_query=True,
               return_patch_scores=False, scores_out: torch.Tensor | None = None):
    """Position-aligned per-patch cosine + mean + top-k (fp_score_topk).

    feats_t (B,P,D) bf16, feat_q (P,D) or (1,P,D) bf16.  Returns (scores fp32 (B,), idx int32 (k,),
    vals fp32 (k,), patch_scores or None).
    """
    B, P, D = feats_t.shape
    feat_q = feat_q.reshape(P, D)
    assert feats_t.dtype == bf16 and feat_q.dtype == bf16
    dev = feats_t.device
    lib = load()
    ws = _ws(lib.fp_score_workspace_bytes(B, P, D), dev)
    if scores_out is not None:  # e.g. this rank's slice of the all-gather buffer: no copy between score and NCCL
        assert scores_out.dtype == torch.float32 and scores_out.numel() >= B and scores_out.is_contiguous()
        scores = scores_out[:B]
    else:
        scores = torch.empty(B, dtype=torch.float32, device=dev)
    idx = torch.empty(max(k, 1), dtype=torch.int32, device=dev)
    vals = torch.empty(max(k, 1), dtype=torch.float32, device=dev)
    patch = torch.empty(B, P, dtype=torch.float32, device=dev) if return_patch_scores else None
    if weights is not None:
        assert weights.dtype == torch.float32 and weights.shape == (B, P)
    check(lib.fp_score_topk(ptr(feats_t), ptr(feat_q), ptr(weights), B, P, D, int(bool(normalise_query)),
                            ptr(scores), ptr(patch), k, ptr(idx), ptr(vals), ptr(ws), ws.numel(), stream_ptr()),
          "fp_score_topk")
    return scores, idx[:k], vals[:k], patch


def topk(scores: torch.Tensor, k: int):
    B = scores.numel()
    dev = scores.device
    ws = _ws(B, dev)
    idx = torch.empty(max(k, 1), dtype=torch.int32, device=dev)
    vals = torch.empty(max(k, 1), dtype=torch.float32, device=dev)
    check(load().fp_topk(ptr(scores), B, k, ptr(idx), ptr(vals), ptr(ws), ws.numel(), stream_ptr()), "fp_topk")
    return idx[:k], vals[:k]


def ffa_pool(feats: torch.Tensor, masks: torch.Tensor):
    """feats (V,P,D) bf16, masks (V,res,res) bool/u8 -> (V,D) fp32, valid-count (V,) int32."""
    V, P, D = feats.shape
    res = masks.shape[-1]
    m = masks.to(torch.uint8).contiguous()
    out = torch.empty(V, D, dtype=torch.float32, device=feats.device)
    valid = torch.empty(V, dtype=torch.int32, device=feats.device)
    check(load().fp_ffa_pool(ptr(feats), ptr(m), V, res, D, ptr(out), ptr(valid), stream_ptr()), "fp_ffa_pool")
    return out, valid


# ------------------------------------------------------------------------------------------- retrieval
def normalize_rows(x: torch.Tensor, out: torch.Tensor | None = None):
    """F.normalize(x.to(bfloat16), dim=-1) for x (rows, D) fp32 or bf16 on the device -> (rows, D) bf16."""
    assert x.dim() == 2 and x.dtype in (torch.float32, torch.bfloat16)
    rows, D = x.shape
    if out is None:
        out = torch.empty(rows, D, dtype=torch.bfloat16, device=x.device)
    check(load().fp_normalize_rows(ptr(x), int(x.dtype == torch.float32), rows, D, ptr(out), stream_ptr()),
          "fp_normalize_rows")
    return out


def retrieval_scan(db: torch.Tensor, queries: torch.Tensor):
    """db (M,D), queries (Q,D): normalised bf16 -> (Q,M) fp32 scores (bf16-valued, like ``(db @ q).float()``)."""
    assert db.dtype == torch.bfloat16 and queries.dtype == torch.bfloat16 and db.shape[1] == queries.shape[1]
    M, D = db.shape
    Q = queries.shape[0]
    scores = torch.empty(Q, M, dtype=torch.float32, device=db.device)
    check(load().fp_retrieval_scan(ptr(db), ptr(queries), M, D, Q, ptr(scores), stream_ptr()), "fp_retrieval_scan")
    return scores


def topk_rows(scores: torch.Tensor, k: int):
    """torch.topk(scores, k, dim=1) with a defined tie order (lowest index): (Q,M) fp32 -> idx (Q,k) int32, val (Q,k)."""
    assert scores.dim() == 2 and scores.dtype == torch.float32
    Q, M = scores.shape
    idx = torch.empty(Q, k, dtype=torch.int32, device=scores.device)
    val = torch.empty(Q, k, dtype=torch.float32, device=scores.device)
    check(load().fp_topk_rows(ptr(scores), Q, M, k, ptr(idx), ptr(val), stream_ptr()), "fp_topk_rows")
    return idx, val


def retrieval_fine(views: torch.Tensor, view_start: torch.Tensor, view_count: torch.Tensor, max_views: int,
                   cand: torch.Tensor, queries: torch.Tensor, k: int):
    """Per candidate mesh: float32 mean of the top-k per-view scores.  views (N,D) normalised bf16, view_start (M,)
    int64 / view_count (M,) int32 rows of each mesh, cand (Q,C) int32, queries (Q,D) bf16 -> (Q,C) fp32."""
    assert views.dtype == torch.bfloat16 and view_start.dtype == torch.int64 and view_count.dtype == torch.int32
    assert cand.dtype == torch.int32 and queries.dtype == torch.bfloat16
    Q, C = cand.shape
    out = torch.empty(Q, C, dtype=torch.float32, device=views.device)
    check(load().fp_retrieval_fine(ptr(views), ptr(view_start), ptr(view_count), int(max_views), ptr(cand),
                                   ptr(queries), Q, C, views.shape[1], k, ptr(out), stream_ptr()), "fp_retrieval_fine")
    return out


def softvote_add(acc: torch.Tensor, idx: torch.Tensor, val: torch.Tensor):
    """acc (P,M) fp32 += one frame's sparse scores: idx (P,C) int32 (unique per row, <0 skipped), val (P,C) fp32."""
    assert acc.dtype == torch.float32 and idx.dtype == torch.int32 and val.dtype == torch.float32
    P, C = idx.shape
    check(load().fp_softvote_add(ptr(acc), ptr(idx), ptr(val), P, C, acc.shape[1], stream_ptr()), "fp_softvote_add")
    return acc


def softvote_mean(acc: torch.Tensor, frames: int):
    out = torch.empty_like(acc)
    check(load().fp_softvote_mean(ptr(acc), ptr(out), acc.numel(), frames, stream_ptr()), "fp_softvote_mean")
    return out


# ------------------------------------------------------------------------------------------- raster
@functools.lru_cache(maxsize=None)
def _gamma_lut_host() -> np.ndarray:
    i = np.arange(65536, dtype=np.float64) / 65535.0
    return np.floor(255.0 * np.power(i, 1.0 / 2.2) + 0.5).astype(np.uint8)


_gamma_lut_dev = {}


def gamma_lut(device) -> torch.Tensor:
    key = str(device)
    if key not in _gamma_lut_dev:
        _gamma_lut_dev[key] = torch.from_numpy(_gamma_lut_host()).to(device)
    return _gamma_lut_dev[key]


_srgb_lut_dev = {}


def srgb_lut(device) -> torch.Tensor:
    key = str(device)
    if key not in _srgb_lut_dev:
        from .pipeline.utils import srgb_to_linear_lut
        _srgb_lut_dev[key] = torch.from_numpy(srgb_to_linear_lut()).to(device)
    return _srgb_lut_dev[key]


def rasterize(verts: torch.Tensor, faces: torch.Tensor, colors, poses: torch.Tensor, fx, fy, cx, cy,
              res: int, msaa: int = 4, cull_backfaces: bool = False, uv=None, texture=None, points: bool = False,
              ambient: float = 0.0, znear: float = 0.0, zfar: float = 0.0, view_k=None):
    """verts (V,3) fp32, faces (F,3) int32, colors (V,3) u8 | None, poses (B,4,4)|(B,3,4) fp32 -> rgb u8 (B,res,res,3),
    depth fp32 (B,res,res).  ``texture`` = (RGBA8 mip chain u8 tensor, w, h, levels) with ``uv`` (V,2) fp32;
    ``points`` renders the vertices as 1-pixel point sprites (faces ignored).  ``ambient`` / ``znear`` / ``zfar``: 0 = the
    renderer.py defaults (2, 0.05, 100); ``view_k`` (B,4) fp32 = per-view fx,fy,cx,cy."""
    dev = verts.device
    B = poses.shape[0]
    p34 = poses[:, :3, :4].to(torch.float32).contiguous()
    V, F = verts.shape[0], faces.shape[0]
    assert verts.dtype == torch.float32 and faces.dtype == torch.int32
    assert colors is None or colors.dtype == torch.uint8
    lib = load()
    nbytes = C.c_size_t(0)
    check(lib.fp_raster_workspace_bytes(B, V, 0 if points else F, res, msaa, C.byref(nbytes)), "fp_raster_workspace_bytes")
    ws = _ws(nbytes.value, dev)
    rgb = torch.empty(B, res, res, 3, dtype=torch.uint8, device=dev)
    depth = torch.empty(B, res, res, dtype=torch.float32, device=dev)
    chain, tw, th, tl = texture if texture is not None else (None, 0, 0, 0)
    if texture is not None:
        assert uv is not None and uv.dtype == torch.float32 and uv.shape == (V, 2) and chain.dtype == torch.uint8
    args = _lib.RasterArgs(ptr(verts), ptr(faces), ptr(colors), V, F, ptr(p34), B, float(fx), float(fy), float(cx),
                           float(cy), res, msaa, int(cull_backfaces), ptr(gamma_lut(dev)), ptr(rgb), ptr(depth),
                           int(points), ptr(uv) if texture is not None else None, ptr(chain), int(tw), int(th), int(tl),
                           ptr(srgb_lut(dev)) if texture is not None else None, float(ambient), float(znear), float(zfar),
                           ptr(view_k))
    if view_k is not None:
        assert view_k.dtype == torch.float32 and view_k.shape == (B, 4)
    check(lib.fp_rasterize(C.byref(args), ptr(ws), ws.numel(), stream_ptr()), "fp_rasterize")
    return rgb, depth


def rasterize_mesh(mesh, poses: torch.Tensor, fx, fy, cx, cy, res: int, msaa: int = 4, cull_backfaces: bool = False, **kw):
    """Any :class:`pipeline.utils.Mesh` (vertex-coloured, textured or point cloud) -> rgb, depth."""
    from .pipeline.utils import mesh_texture_to_device, mesh_to_device
    dev = poses.device
    v, f, c = mesh_to_device(mesh, dev)
    tex = mesh_texture_to_device(mesh, dev)
    if tex is not None:
        return rasterize(v, f, c, poses, fx, fy, cx, cy, res, msaa, cull_backfaces, uv=tex["uv"],
                         texture=(tex["chain"], tex["w"], tex["h"], tex["levels"]), **kw)
    return rasterize(v, f, c, poses, fx, fy, cx, cy, res, msaa, cull_backfaces, points=mesh.is_point_cloud, **kw)


# ------------------------------------------------------------------------------------------- geometry
def mask_bbox(depth: torch.Tensor, fallback=(105, 315), min_count: int = 100, return_mask: bool = False):
    B, res, _ = depth.shape
    dev = depth.device
    bbox = torch.empty(B, 4, dtype=torch.int32, device=dev)
    count = torch.empty(B, dtype=torch.int32, device=dev)
    mask = torch.empty(B, res, res, dtype=torch.uint8, device=dev) if return_mask else None
    check(load().fp_mask_bbox(ptr(depth), B, res, int(fallback[0]), int(fallback[1]), min_count, ptr(bbox),
                              ptr(count), ptr(mask), stream_ptr()), "fp_mask_bbox")
    return bbox, count, mask


@functools.lru_cache(maxsize=None)
def _norm_lut_host() -> torch.Tensor:
    """Normalize(bf16(v/255)) per channel, computed with the reference's own torch ops (dino.py:12,16 on a
    bf16 tensor; the /255 and .float() of renderer.py:121)."""
    v = torch.from_numpy(np.arange(256) / 255).float().to(bf16)               # (256,)
    mean = torch.as_tensor((0.485, 0.456, 0.406), dtype=bf16).view(3, 1)
    std = torch.as_tensor((0.229, 0.224, 0.225), dtype=bf16).view(3, 1)
    x = v.view(1, 256).repeat(3, 1)
    return x.sub_(mean).div_(std).contiguous()                                   # (3,256) bf16


_norm_lut_dev = {}


def norm_lut(device) -> torch.Tensor:
    key = str(device)
    if key not in _norm_lut_dev:
        _norm_lut_dev[key] = _norm_lut_host().to(device)
    return _norm_lut_dev[key]


def crop_resize_pad(src: torch.Tensor, boxes: torch.Tensor, target: int, to_patches: bool = False,
                    out: torch.Tensor | None = None):
    """CropResizePad gather.  src: u8 (B,H,W,3) or fp32 (B,3,H,W); boxes (B,4) int32 xyxy (exclusive).
    `out` (patch-matrix mode): a preallocated (>= B*g*g, 640) bf16 buffer whose first rows are written."""
    dev = src.device
    B = src.shape[0]
    u8 = src.dtype == torch.uint8
    H, W = (src.shape[1], src.shape[2]) if u8 else (src.shape[2], src.shape[3])
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    if to_patches:
        g = target // 14
        if out is not None:
            assert out.dtype == bf16 and out.is_contiguous() and out.shape[1] == KPAD and out.shape[0] >= B * g * g
            dst = out
        else:
            dst = torch.empty(B * g * g, KPAD, dtype=bf16, device=dev)
    else:
        dst = torch.empty(B, 3, target, target, dtype=torch.float32, device=dev)
    boxes = boxes.to(torch.int32).contiguous()
    check(load().fp_crop_resize_pad(ptr(src), int(u8), ptr(boxes), ptr(norm_lut(dev)), ptr(dst), int(to_patches), B,
                                    H, W, target, KPAD, ptr(status), stream_ptr()), "fp_crop_resize_pad")
    return dst, status


@functools.lru_cache(maxsize=256)
def _kinv_device(k_bytes: bytes, dev_index: int) -> torch.Tensor:
    """inv(K) on the device, uploaded once per intrinsic matrix (a blocking pageable copy per call would make the host
    wait for everything queued on the stream -- the whole ViT forward in the per-frame loop)."""
    K = np.frombuffer(k_bytes, dtype=np.float64).reshape(3, 3)
    return torch.from_numpy(np.linalg.inv(K).reshape(9).copy()).to(torch.device("cuda", dev_index))


def depth_extents(depth: torch.Tensor, K, view_idx=None):
    """(n,8) fp64: xmin,xmax,ymin,ymax,sum_x,sum_y,sum_z,count of K^-1 [u v 1]^T d over non-zero points."""
    B, res, _ = depth.shape
    dev = depth.device
    kinv = _kinv_device(np.ascontiguousarray(K, dtype=np.float64).tobytes(), dev.index)
    if view_idx is not None:
        view_idx = view_idx.to(torch.int32).contiguous()
        n = view_idx.numel()
    else:
        n = B
    out = torch.empty(n, 8, dtype=torch.float64, device=dev)
    check(load().fp_depth_extents(ptr(depth), ptr(view_idx), n, res, ptr(kinv), ptr(out), stream_ptr()),
          "fp_depth_extents")
    return out


# ------------------------------------------------------------------------------------------- refiner confidence pass
def roi_align(image: torch.Tensor, boxes: torch.Tensor, out_h: int, out_w: int, sampling_ratio: int = 2):
    """torchvision.ops.roi_align(image[None], [0 | boxes], (out_h, out_w), sampling_ratio=...) for one (C,H,W) fp32
    image and (n,4) fp32 xyxy boxes -> (n,C,out_h,out_w) fp32 (reference refiner_utils.py:128-133)."""
    assert image.dtype == torch.float32 and image.dim() == 3 and boxes.dtype == torch.float32 and boxes.shape[1] == 4
    Cn, H, W = image.shape
    n = boxes.shape[0]
    out = torch.empty(n, Cn, out_h, out_w, dtype=torch.float32, device=image.device)
    check(load().fp_roi_align(ptr(image.contiguous()), Cn, H, W, ptr(boxes.contiguous()), n, out_h, out_w,
                              sampling_ratio, ptr(out), stream_ptr()), "fp_roi_align")
    return out


def depth_mask_cubic(depth: torch.Tensor, g: int, res: int | None = None):
    """(B,S,S) fp32 depth, image = its top-left res x res (default S) -> (B,g,g) bool:
    cv2.resize((depth > 0).astype(float32), (g,g), INTER_CUBIC) > 0.5 (reference tracking_refiner.py:75)."""
    B, S, _ = depth.shape
    res = S if res is None else res
    mask = torch.empty(B, g, g, dtype=torch.uint8, device=depth.device)
    check(load().fp_depth_mask_cubic(ptr(depth.contiguous()), B, res, S, g, ptr(mask), stream_ptr()),
          "fp_depth_mask_cubic")
    return mask.bool()


def patch_cosine(feats_a: torch.Tensor, feats_b: torch.Tensor, mask: torch.Tensor | None = None):
    """(..., D) bf16 token rows x 2 (+ optional bool/u8 mask over the rows) -> (...) fp32 masked cosine
    (reference tracking_refiner.py:80-88)."""
    assert feats_a.shape == feats_b.shape and feats_a.dtype == bf16 and feats_b.dtype == bf16
    D = feats_a.shape[-1]
    rows = feats_a.numel() // D
    m = None if mask is None else mask.to(torch.uint8).contiguous()
    out = torch.empty(feats_a.shape[:-1], dtype=torch.float32, device=feats_a.device)
    check(load().fp_patch_cosine(ptr(feats_a.contiguous()), ptr(feats_b.contiguous()), ptr(m), rows, D, ptr(out),
                                 stream_ptr()), "fp_patch_cosine")
    return out
