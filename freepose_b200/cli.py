"""The three hot-path CLIs of the reference, re-hosted on the B200 engine with the reference's flags:

    python -m scripts.extract_retrieval_features   (reference scripts/extract_retrieval_features.py:12-76)
    python -m scripts.dino_inference               (reference scripts/dino_inference.py:22-133)
    python -m scripts.dino_inference_video         (reference scripts/dino_inference_video.py:43-245)

Dataset readers (BOPDataset, WebTemplateDataset: PNG/tar decoding, out of scope per SURVEY.md section 2) are taken from a
reference checkout through the ``src`` overlay ($FREEPOSE_REFERENCE_ROOT).  ``--synthetic`` replaces them with seeded
synthetic scenes so every CLI runs self-contained (tests, demos, the GPU box that has no datasets).
Outputs keep the reference's formats: ``(views, 1024)`` fp32 ``.npy`` per mesh; BOP CSV rows
``scene_id,im_id,obj_id,score,R,t,bbox_visib,scale,time`` (t in mm for static images, metres for video).
"""
from __future__ import annotations

import argparse
import os
from pathlib import Path

import numpy as np
import torch

from . import ops
from .pipeline.estimators.online_pose_estimator import DinoOnlinePoseEstimator
from .pipeline.estimators.pose_estimator import DinoPoseEstimator
from .pipeline.proposals import Proposals
from .pipeline.retrieval.dino import DINOv2FeatureExtractor
from .pipeline.retrieval.renderer import MeshRenderer
from .synthetic import synthetic_mesh


def _extractor_kwargs(args):
    kw = {}
    if getattr(args, "weights", None):
        kw["weights"] = args.weights
    if getattr(args, "synthetic_depth", None):
        kw["depth"] = args.synthetic_depth
    return kw


def _csv_row(results, scene_id, im_id, obj_id, out, bbox_xyxy, scale, t_unit, time_value):
    R = out["TCO"][0][:3, :3].flatten().tolist()
    t = out["TCO"][0][:3, 3].tolist()
    b = [float(x) for x in np.asarray(bbox_xyxy)]
    results["scene_id"].append(int(scene_id))
    results["im_id"].append(int(im_id))
    results["obj_id"].append(obj_id)
    results["score"].append(float(np.asarray(out["scores"][0])))
    results["R"].append(" ".join(str(x) for x in R))
    results["t"].append(" ".join(str(x * t_unit) for x in t))
    results["bbox_visib"].append(" ".join(str(x) for x in [b[0], b[1], b[2] - b[0], b[3] - b[1]]))
    results["scale"].append(scale)
    results["time"].append(time_value)


def _new_results():
    return {k: [] for k in ("scene_id", "im_id", "obj_id", "score", "R", "t", "bbox_visib", "scale", "time")}


def _write_csv(results, path: Path):
    import pandas as pd
    path.parent.mkdir(parents=True, exist_ok=True)
    pd.DataFrame(results).to_csv(path, index=False, header=True)


# ------------------------------------------------------------------------------------------ synthetic scenes
class SyntheticTemplates:
    """Stand-in for WebTemplateDataset: renders the views on the device instead of decoding PNG shards; returns the
    reference's sample schema (src/dataloader/template.py:98-99)."""

    def __init__(self, n_meshes, n_views, resolution, crop=True, subdivisions=4, device="cuda"):
        self.n, self.res, self.crop = n_meshes, resolution, crop
        self.renderer = MeshRenderer(n_views, resolution=resolution, device=device)
        self.subdiv = subdivisions

    def __len__(self):
        return self.n

    def mesh(self, idx):
        return synthetic_mesh(seed=idx, subdivisions=self.subdiv)

    def __getitem__(self, idx):
        rgb, depth = self.renderer.render_device(self.mesh(idx))
        if self.crop:
            templates, _, masks, _ = self.renderer.proposals_device(rgb, depth, self.res, to_patches=False)
            masks = masks.bool()
        else:
            templates = (rgb.float() / 255).permute(0, 3, 1, 2).contiguous()
            masks = depth > 0
        f, c = self.renderer.focal, self.res / 2
        return {"templates": templates, "masks": masks, "depths": depth, "model_name": f"synthetic_{idx:06d}",
                "tar_file": "", "intrinsic": torch.tensor([[f, 0, c], [0, f, c], [0, 0, 1]], dtype=torch.float64)}

    def get_template_by_name(self, name):
        return self[int(name.rsplit("_", 1)[1])]


def synthetic_frame(meshes, h=480, w=640, seed=0, device="cuda"):
    """A frame with the meshes projected at random poses over a noise background + per-object boxes/masks."""
    rng = np.random.default_rng(seed)
    f = float(np.sqrt(h ** 2 + w ** 2))  # reference: K from the image diagonal (dino_inference_video.py:116-118)
    K = np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1]])
    img = rng.integers(0, 60, (h, w, 3), dtype=np.uint8)
    boxes, masks, poses = [], [], []
    side = max(h, w)
    side += (-side) % 4
    for i, mesh in enumerate(meshes):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        pose = np.eye(4)
        pose[:3, :3] = q
        pose[:3, 3] = [rng.uniform(-0.5, 0.5), rng.uniform(-0.3, 0.3), rng.uniform(2.2, 3.0)]
        rgb, depth = ops.rasterize_mesh(mesh, torch.from_numpy(pose[None]).float().to(device), f, f, w / 2, h / 2, side)
        rgb, depth = rgb[0, :h, :w].cpu().numpy(), depth[0, :h, :w].cpu().numpy()
        m = depth > 0
        if m.sum() < 50:
            continue
        img[m] = rgb[m]
        ys, xs = np.nonzero(m)
        boxes.append([xs.min(), ys.min(), xs.max(), ys.max()])
        masks.append(m)
        poses.append(pose)
    return img, K, np.array(boxes), np.array(masks), poses


# ------------------------------------------------------------------------------------------ extract_retrieval_features
def run_extract_retrieval_features(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--shards_folder", type=str, default="objaverse_shards")
    ap.add_argument("--filelist", type=str, default="mesh_cache.csv")
    ap.add_argument("--feature", type=str, default="ffa", choices=["ffa", "cls"])
    ap.add_argument("--layer", type=int, default=22)
    ap.add_argument("--mesh_per_job", type=int, default=100)
    ap.add_argument("--batch_size", type=int, default=128)
    ap.add_argument("--weights", type=str, default=None)
    ap.add_argument("--synthetic", type=int, default=0, help="number of synthetic meshes instead of the shard dataset")
    ap.add_argument("--synthetic_views", type=int, default=42)
    ap.add_argument("--synthetic_depth", type=int, default=None)
    ap.add_argument("--resolution", type=int, default=420)
    ap.add_argument("--out_dir", type=str, default=None)
    args = ap.parse_args(argv)

    out_dir = Path(args.out_dir) if args.out_dir else \
        Path("data/datasets").resolve() / f"{args.shards_folder}_{args.feature}_{args.layer}"
    out_dir.mkdir(parents=True, exist_ok=True)
    model = DINOv2FeatureExtractor(chunk=args.batch_size, **_extractor_kwargs(args))
    feature_type = "cls" if args.feature == "cls" else "patch"
    if args.synthetic:
        dataset = SyntheticTemplates(args.synthetic, args.synthetic_views, args.resolution, crop=False)
    else:
        from src.dataloader.template import WebTemplateDataset  # reference reader through the overlay
        dataset = WebTemplateDataset((Path("data/datasets").resolve() / args.shards_folder).as_posix(),
                                     (Path("data").resolve() / args.filelist).as_posix(), crop=False)
    job_id = int(os.environ.get("SLURM_ARRAY_TASK_ID", 0))
    start = job_id * args.mesh_per_job
    end = min(start + args.mesh_per_job, len(dataset))
    written = []
    for idx in range(start, end):
        sample = dataset[idx]
        if sample["templates"] is None:
            continue
        feats = model(sample["templates"], layer=args.layer, feature_type=feature_type)
        if args.feature == "ffa":
            pooled, valid = ops.ffa_pool(feats, torch.as_tensor(sample["masks"]).to(feats.device))
            keep = (valid > 0).cpu().numpy()          # the reference skips NaN (empty-mask) views with a warning
            arr = pooled.cpu().numpy()[keep]
        else:
            arr = feats.float().cpu().numpy()
        path = out_dir / f"{sample['model_name']}.npy"
        np.save(path.as_posix(), arr)
        written.append(path)
    return written


# ------------------------------------------------------------------------------------------ dino_inference
def run_dino_inference(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", type=str)
    ap.add_argument("--split", type=str, default="test")
    ap.add_argument("--proposals", type=str)
    ap.add_argument("--layer", type=int, default=22)
    ap.add_argument("--depth_method", type=str, default="zoedepth")
    ap.add_argument("--bbox_extend", type=float, default=0.05)
    ap.add_argument("--batch_size", type=int, default=128)
    ap.add_argument("--cache_size", type=int, default=50)
    ap.add_argument("--save_all_cache", action="store_true")
    ap.add_argument("--weights", type=str, default=None)
    ap.add_argument("--synthetic", type=int, default=0, help="number of synthetic images instead of a BOP dataset")
    ap.add_argument("--synthetic_depth", type=int, default=None)
    ap.add_argument("--n_poses", type=int, default=600)
    ap.add_argument("--resolution", type=int, default=420)
    ap.add_argument("--out", type=str, default=None)
    args = ap.parse_args(argv)

    task = int(os.getenv("SLURM_ARRAY_TASK_ID", 0))
    model = DinoPoseEstimator(n_poses=args.n_poses, cache_size=args.cache_size, save_all=args.save_all_cache,
                              cache_dir=f"./data/cache_{task}_{args.dataset}", resolution=args.resolution,
                              chunk=args.batch_size, **_extractor_kwargs(args))
    results = _new_results()
    if args.synthetic:
        templates = SyntheticTemplates(4, args.n_poses, args.resolution, crop=True)
        for im in range(args.synthetic):
            meshes = [templates.mesh(i) for i in range(2)]
            full = [m.copy().apply_scale(4.0 * 0.3) for m in meshes]   # metric objects of ~0.3 m half-extent
            img, K, boxes, masks, _ = synthetic_frame(full, seed=im)
            if len(boxes) == 0:
                continue
            props = Proposals(img, {"boxes": torch.from_numpy(boxes), "masks": torch.from_numpy(masks)},
                              args.resolution, bbox_extend=args.bbox_extend)
            for j, prop in enumerate(props.proposals):
                entry = templates[j]
                out = model(prop, entry, K, boxes[j].astype(np.float64), 0.3, layer=args.layer, batch_size=args.batch_size)
                _csv_row(results, 0, im, entry["model_name"], out, boxes[j], 0.3, 1000.0, 0.2)
        out_path = Path(args.out or "./data/results/synthetic/pose_outputs_0.csv")
    else:
        import json
        from sam2.utils.amg import rle_to_mask
        from src.dataloader.bop import BOPDataset
        from src.dataloader.template import WebTemplateDataset
        proposals_path = Path("./data/results").resolve() / args.dataset / args.proposals
        out_path = Path("./data/results").resolve() / args.dataset / args.proposals.replace(
            ".json", f"_dinopose_layer_{args.layer}_bbext_{args.bbox_extend}_depth_{args.depth_method}_cache_{args.cache_size}")
        out_path = out_path / f"pose_outputs_{task}.csv"
        dataset = BOPDataset(f"data/datasets/{args.dataset}/", args.split)
        templates = WebTemplateDataset("data/datasets/objaverse_shards", "data/mesh_cache.csv", bbox_extend=args.bbox_extend)
        with open(proposals_path) as f:
            props_json = json.load(f)
        per_task = 30
        for scene_idx in range(task * per_task, min((task + 1) * per_task, len(dataset))):
            entry = dataset[scene_idx]
            scene_id, frame_id = int(entry["scene_id"]), int(entry["frame_id"])
            sp = [p for p in props_json if p["scene_id"] == scene_id and p["image_id"] == frame_id]
            if not sp:
                continue
            masks = torch.from_numpy(np.stack([rle_to_mask(p["segmentation"]) for p in sp]))
            boxes = torch.from_numpy(np.stack([np.array(p["bbox"]) for p in sp]))
            boxes[:, 2:] += boxes[:, :2]
            if args.depth_method.startswith("const-"):
                scales = [float(args.depth_method.split("-")[1])] * len(sp)
            elif args.depth_method == "zoedepth":
                scales = [np.clip(p["scale"], a_min=0.01, a_max=None) for p in sp]
            else:
                raise SystemExit("--depth_method depthmap needs the reference's scale stage (out of scope here)")
            props = Proposals(entry["image"], {"boxes": boxes, "masks": masks}, 420, bbox_extend=args.bbox_extend)
            for j, prop in enumerate(props.proposals):
                mesh_entry = templates.get_template_by_name(sp[j]["mesh"])
                out = model(prop, mesh_entry, entry["intrinsic"], boxes[j], scales[j], layer=args.layer,
                            batch_size=args.batch_size)
                _csv_row(results, scene_id, frame_id, sp[j]["mesh"], out, boxes[j].numpy(), scales[j], 1000.0, 0.2)
    _write_csv(results, out_path)
    return out_path


# ------------------------------------------------------------------------------------------ dino_inference_video
def run_dino_inference_video(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--video_folder", type=str, default=None)
    ap.add_argument("--proposals", type=str, default=None)
    ap.add_argument("--layer", type=int, default=22)
    ap.add_argument("--bbox_extend", type=float, default=0.05)
    ap.add_argument("--batch_size", type=int, default=128)
    ap.add_argument("--cache_size", type=int, default=50)
    ap.add_argument("--neighborhood", type=float, default=15)
    ap.add_argument("--n_fine_poses", type=int, default=20000)
    ap.add_argument("--no_rescore", action="store_true")
    ap.add_argument("--mask_scores", action="store_true")
    ap.add_argument("--weights", type=str, default=None)
    ap.add_argument("--synthetic", type=int, default=0, help="number of synthetic frames instead of a video folder")
    ap.add_argument("--synthetic_depth", type=int, default=None)
    ap.add_argument("--n_poses", type=int, default=600)
    ap.add_argument("--resolution", type=int, default=420)
    ap.add_argument("--out", type=str, default=None)
    args = ap.parse_args(argv)
    if not args.synthetic:
        raise SystemExit("video folders need the reference's frame/proposal readers; run with the src overlay and "
                         "FREEPOSE_REFERENCE_ROOT set, or use --synthetic N")
    model = DinoOnlinePoseEstimator(n_coarse_poses=args.n_poses, n_fine_poses=args.n_fine_poses,
                                    cache_size=args.cache_size, cache_dir="./data/cache_video",
                                    resolution=args.resolution, chunk=args.batch_size, **_extractor_kwargs(args))
    templates = SyntheticTemplates(2, args.n_poses, args.resolution, crop=True)
    meshes_r = [templates.mesh(i) for i in range(2)]
    meshes_full = [m.copy().apply_scale(4.0) for m in meshes_r]
    metric = [m.copy().apply_scale(4.0 * 0.3) for m in meshes_r]
    entries = [templates[i] for i in range(2)]
    prev_poses = [None, None]
    results = _new_results()
    img0, K, boxes0, masks0, _ = synthetic_frame(metric, seed=0)
    for frame in range(args.synthetic):
        # a static synthetic scene observed for N frames: exercises the prev_pose carry of the reference loop
        props = Proposals(img0, {"boxes": torch.from_numpy(boxes0), "masks": torch.from_numpy(masks0)},
                          args.resolution, bbox_extend=args.bbox_extend)
        if args.no_rescore:
            outs = [model.coarse_estimator(props.proposals[j], entries[j], K, boxes0[j].astype(np.float64), 0.3,
                                           layer=args.layer, batch_size=args.batch_size) for j in range(len(boxes0))]
        else:
            # all proposals of the frame in one ViT pass (same results as one model(...) call per proposal)
            items = [dict(proposal=props.proposals[j], proposal_mask=props.proposals_masks[j], template_dict=entries[j],
                          mesh=meshes_full[j], K=K, bbox=boxes0[j].astype(np.float64), est_scale=0.3,
                          prev_pose=prev_poses[j]) for j in range(len(boxes0))]
            outs = model.forward_batch(items, neighborhood=args.neighborhood, layer=args.layer,
                                       batch_size=args.batch_size, mask_scores=args.mask_scores)
            for j, out in enumerate(outs):
                prev_poses[j] = out["TCO"][0]
        for j, out in enumerate(outs):
            _csv_row(results, 0, frame, entries[j]["model_name"], out, boxes0[j], 0.3, 1.0, -1)
    out_path = Path(args.out or "./data/results/synthetic/video_pose_outputs.csv")
    _write_csv(results, out_path)
    return out_path
