"""Mints tests/golden/*.npz by running the REFERENCE'S OWN CODE (imported unmodified from /root/reference via
oracle/refimport.py) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures travel with the repository; tests compare the oracle restatements (and, on the GPU, the CUDA
kernels) against them.  Nothing at test time reads /root/reference.
"""
from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import refimport  # noqa: E402

OUT = Path(__file__).resolve().parent


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_poses():
    pe = refimport.import_reference("src.pipeline.estimators.pose_estimator")
    gen = pe.DinoPoseEstimator.generate_poses
    p42 = np.array(gen(42))
    p600 = np.array(gen(600))
    p20k = np.array(gen(20000))
    np.savez_compressed(OUT / "poses.npz", p42=p42, p600_first8=p600[:8], p600_last8=p600[-8:],
                        p600_sha=sha(p600), p20000_sha=sha(p20k), p20000_rows=p20k[[0, 1, 7777, 19999]])


def golden_geometry():
    ru = refimport.import_reference("src.pipeline.utils")
    rng = np.random.default_rng(7)
    cases = []
    K_t = np.array([[600, 0, 210], [0, 600, 210], [0, 0, 1]])
    for i in range(4):
        res = 420 if i % 2 == 0 else 224
        Kt = K_t if res == 420 else np.array([[320.0, 0, 112], [0, 320.0, 112], [0, 0, 1]])
        d = np.zeros((res, res), np.float32)
        y0, x0 = rng.integers(10, res // 3, 2)
        h, w = rng.integers(res // 4, res // 2, 2)
        d[y0:y0 + h, x0:x0 + w] = rng.uniform(0.85, 1.35, (h, w)).astype(np.float32)
        d[rng.random((res, res)) < 0.3] = 0
        bbox = np.array([100.0 + i, 80.0, 300.0, 260.0 + 2 * i])
        Kq = np.array([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
        T0 = np.eye(4); T0[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]; T0[2, 3] = 1.1
        est_scale = float(rng.uniform(0.05, 0.6))
        pc = ru.depthmap_to_pointcloud(d, Kt)
        # coarse rescaling, pose_estimator.py:104-111
        pcc = pc.copy(); m = pcc.mean(axis=0); pcc -= m; pcc /= 0.25; pcc *= est_scale; pcc += m
        tco_coarse = ru.get_z_from_pointcloud(bbox, pcc, Kq, T0)
        # fine rescaling, online_pose_estimator.py:82-86
        pcf = pc.copy(); pcf /= 0.25; pcf *= est_scale
        tco_fine = ru.get_z_from_pointcloud(bbox, pcf, Kq, T0)
        cases.append(dict(depth=d, Kt=Kt, bbox=bbox, Kq=Kq, T0=T0, est_scale=est_scale, n_points=pc.shape[0],
                          pc_sha=sha(pc), tco_coarse=tco_coarse, tco_fine=tco_fine,
                          bbox_of_mask=ru.mask_to_bbox(d > 0)))
    np.savez_compressed(OUT / "geometry.npz", **{f"c{i}_{k}": v for i, c in enumerate(cases) for k, v in c.items()})


def golden_crop():
    bu = refimport.import_reference("src.utils.bbox_utils")
    rng = np.random.default_rng(11)
    store = {}
    # (H, W, T, bbox_extend, boxes): small full-output cases + large checksum cases
    specs = [
        (60, 90, 84, 0, [[5, 7, 50, 40], [10, 10, 40, 40], [0, 0, 90, 60]]),
        (60, 90, 84, 0.2, [[5, 7, 50, 40], [30, 20, 88, 58], [12, 3, 31, 57]]),
        (420, 420, 420, 0, [[100, 120, 330, 344], [150, 100, 287, 390], [105, 105, 314, 314]]),
        (420, 420, 224, 0, [[100, 120, 330, 344], [50, 60, 380, 200], [0, 0, 419, 419]]),
        (480, 640, 420, 0.2, [[345, 136, 575, 360], [274, 255, 498, 479], [10, 20, 600, 300]]),
        (480, 640, 224, 0.2, [[317, 4, 617, 272], [0, 0, 100, 470], [500, 400, 639, 479]]),
    ]
    for i, (H, W, T, ext, boxes) in enumerate(specs):
        imgs = rng.random((len(boxes), 3, H, W)).astype(np.float32)
        boxes = np.array(boxes, dtype=np.int64)
        out = bu.CropResizePad(T, (H, W), bbox_extend=ext)(torch.from_numpy(imgs), torch.from_numpy(boxes)).numpy()
        store[f"c{i}_spec"] = np.array([H, W, T], dtype=np.int64)
        store[f"c{i}_ext"] = np.float64(ext)
        store[f"c{i}_boxes"] = boxes
        store[f"c{i}_seed_imgs_sha"] = sha(imgs)
        store[f"c{i}_out_sha"] = sha(out)
        if H * W < 10000:
            store[f"c{i}_imgs"] = imgs
            store[f"c{i}_out"] = out
    store["rng_seed"] = np.int64(11)
    np.savez_compressed(OUT / "crop.npz", **store)


def golden_dino_forward():
    """The reference's DINOv2FeatureExtractor.forward (dino.py:14-32), unmodified, around the hub-shaped oracle
    ViT (2 synthetic blocks, 56 px crops) in fp32: pins Normalize + prepare_tokens + block loop + norm + slices."""
    from freepose_b200.vit_weights import synthetic_state_dict
    from oracle.vit import OracleViT
    sd = synthetic_state_dict(seed=3, depth=2)
    fe = refimport.reference_feature_extractor(OracleViT(sd).float())
    g = torch.Generator().manual_seed(5)
    imgs = torch.rand(2, 3, 56, 56, generator=g)
    out = {"imgs": imgs.numpy()}
    for ft in ("cls", "reg", "patch"):
        for layer in (1, 2):
            out[f"{ft}_{layer}"] = fe(imgs, layer=layer, feature_type=ft).numpy()
    np.savez_compressed(OUT / "dino_forward.npz", **out)


def golden_score():
    """The reference's scoring expression (pose_estimator.py:85-90 / online_pose_estimator.py:68-79) evaluated
    verbatim with einops on CPU bf16 tensors."""
    import torch.nn.functional as F
    from einops import einsum
    g = torch.Generator().manual_seed(9)
    q = torch.randn(1, 9, 1024, generator=g)
    t = (0.6 * q + 0.8 * torch.randn(12, 9, 1024, generator=g))
    t[3] = t[5]  # exact tie
    feats_template, query_feat = t.to(torch.bfloat16), q.to(torch.bfloat16)
    signature = "b n d, b n d -> b n"
    scores = einsum(F.normalize(feats_template, dim=-1), F.normalize(query_feat, dim=-1), signature).mean(dim=-1)
    top_scores, top_indices = torch.topk(scores, 3)
    masks = torch.rand(12, 9, generator=g)
    qn = F.normalize(query_feat, dim=-1)
    s_fine = einsum(qn, F.normalize(feats_template, dim=-1), signature)
    weighted = (s_fine * masks).sum(dim=-1) / masks.sum(dim=-1)
    np.savez_compressed(OUT / "score.npz", feats_t=feats_template.view(torch.int16).numpy(),
                        feat_q=query_feat.view(torch.int16).numpy(), scores=scores.float().numpy(),
                        top_scores=top_scores.float().numpy(), top_indices=top_indices.numpy(),
                        masks=masks.numpy(), weighted=weighted.numpy(),
                        argmax=torch.argmax(scores).numpy(), maxval=torch.max(scores).float().numpy())


def golden_retrieval():
    """The reference's retrieval lines (extract_proposals_ground.py:39-41,136-160, ..._video.py:148-190) restated in
    oracle/retrieval.py reference_* and run on CPU bf16; inputs are regenerated from the seed at test time."""
    import torch.nn.functional as F
    from oracle import retrieval as R
    case = R.synthetic_case(0)
    dbn = R.reference_database(case["db"])
    feats = F.normalize(torch.from_numpy(case["queries"]).to(torch.bfloat16), dim=-1)
    store = {"db_sha": np.frombuffer(bytes.fromhex(sha(case["db"])), dtype=np.uint8),
             "db_norm": dbn[:16].view(torch.int16).numpy()}
    for topk in (0, 3, 10):
        best, score, cand, dense = [], [], [], []
        for f in feats:
            m, sc, I, s = R.reference_retrieve(dbn, case["fine"], f, topk)
            best.append(m); score.append(sc); cand.append(I); dense.append(s[torch.from_numpy(I)].numpy())
        store[f"best_{topk}"] = np.array(best)
        store[f"score_{topk}"] = np.array(score, dtype=np.float64)
        store[f"cand_{topk}"] = np.stack(cand)
        store[f"cand_scores_{topk}"] = np.stack(dense)
    per_frame = []
    for fr in case["video"]:
        ff = F.normalize(torch.from_numpy(fr).to(torch.bfloat16), dim=-1)
        per_frame.append(torch.stack([R.reference_retrieve(dbn, case["fine"], f, 3)[3] for f in ff]))
    I, sc = R.reference_softvote(per_frame)
    store["vote_best"], store["vote_score"] = I, sc
    np.savez_compressed(OUT / "retrieval.npz", **store)


def golden_refiner():
    """The reference's OWN TrackingRefiner.pose_confidence / crop_image / update_K_with_crop /
    _get_threshold_for_confidence, unmodified; torch.hub.load returns the hub-shaped oracle ViT-B (2 blocks, seeded
    synthetic weights) and _render -- pyrender, not installable -- is replaced by the C raster restatement."""
    from PIL import Image
    from freepose_b200.vit_weights import VITB14_REG, synthetic_state_dict
    from oracle import refiner as OR
    from oracle.vit import OracleViT
    sd = synthetic_state_dict(VITB14_REG, seed=3, depth=2)

    class Hub(OracleViT):                       # what torch.hub's DinoVisionTransformer exposes to the refiner
        def forward_features(self, x):
            return {"x_norm_patchtokens": OracleViT.forward_features(self, x, len(self.blocks))[:, 5:]}

    hub = Hub(sd, VITB14_REG)
    tr = refimport.import_reference("src.pipeline.estimators.tracking_refiner")
    ru = sys.modules.get("src.pipeline.refiner_utils") or refimport.import_reference("src.pipeline.refiner_utils")
    ru = tr.refiner_utils
    orig = torch.hub.load
    torch.hub.load = lambda repo, name, *a, **k: hub if "dinov2" in repo else torch.nn.Identity()
    try:
        ref = tr.TrackingRefiner(dino_device="cpu", cotracker_device="cpu")
    finally:
        torch.hub.load = orig
    assert ref.image_size == 518 and ref.feats_size == 37
    ref._render = lambda m, w, h, K, T: OR.render(m, K, T, w)          # the one replaced piece
    rng = np.random.default_rng(11)
    store = {}
    confs = []
    for i in range(2):
        mesh, frame, K, T = OR.synthetic_case(i)
        conf = ref.pose_confidence(mesh, Image.fromarray(frame), K, T)
        crop, bbox, new_K = ref._crop_image(mesh, Image.fromarray(frame), K, T)
        store[f"frame_{i}_sha"], store[f"T_{i}"] = sha(frame), T
        store[f"conf_{i}"], store[f"bbox_{i}"], store[f"new_K_{i}"] = conf, bbox.numpy(), new_K.numpy()
        store[f"crop_{i}_rows"] = crop[:, ::37, ::37].numpy()          # a 14x14 lattice of the 518^2 crop per channel
        store[f"crop_{i}_sha"] = sha(crop.numpy())
        confs.append(conf)
    store["K"] = K
    store["thr"] = np.float64(ref._get_threshold_for_confidence(np.stack(confs)))
    sim = rng.uniform(-0.2, 1.0, size=(5, 37, 37)).astype(np.float32)
    store["sim"], store["sim_thr"] = sim, np.float64(ref._get_threshold_for_confidence(sim))
    # crop_image / update_K_with_crop on a batch of poses (the reference functions themselves)
    pts = torch.from_numpy(np.pad(rng.normal(scale=0.05, size=(100, 3)), ((0, 0), (0, 1)), constant_values=1.)).float()
    Ts = torch.eye(4).repeat(3, 1, 1); Ts[:, 2, 3] = torch.tensor([0.5, 0.8, 1.3]); Ts[:, 0, 3] = torch.tensor([0.0, 0.1, -0.2])
    img = torch.from_numpy(rng.random((3, 120, 160)).astype(np.float32))
    Kt = torch.tensor([[200.0, 0, 80], [0, 200.0, 60], [0, 0, 1]])
    crops, boxes = ru.crop_image(img, Ts, pts, Kt, 64, 48)
    store["ci_pts"], store["ci_Ts"], store["ci_img"], store["ci_K"] = pts.numpy(), Ts.numpy(), img.numpy(), Kt.numpy()
    store["ci_crops"], store["ci_boxes"] = crops.numpy(), boxes.numpy()
    store["ci_newK"] = ru.update_K_with_crop(Kt, boxes, 64, 48).numpy()
    np.savez_compressed(OUT / "refiner.npz", **store)


def online_fine_inputs(seed=21, n_views=6, res=420):
    """Seeded inputs of the fine-stage fixture (regenerated identically by the tests: only outputs are stored)."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(1, 900, 1024, generator=g)
    feats = (0.5 * q + 0.8 * torch.randn(n_views, 900, 1024, generator=g)).to(torch.bfloat16)
    query = (q * (1.0 + 2.0 * torch.rand(1, 900, 1, generator=g))).to(torch.bfloat16)     # per-patch norms differ (raw query)
    yy, xx = torch.meshgrid(torch.arange(res), torch.arange(res), indexing="ij")
    tmasks = torch.stack([((yy - 210 - 9 * i) ** 2 / (60.0 + 9 * i) ** 2 + (xx - 200 + 7 * i) ** 2 / (90.0 - 5 * i) ** 2) < 1
                          for i in range(n_views)])
    pmask = ((yy - 180) ** 2 + (xx - 230) ** 2) < 100 ** 2
    return feats, query, tmasks, pmask


def golden_online_fine():
    """Fine stage of the video path, the reference's own code (online_pose_estimator.py:26-34, 49-79):
    * DinoOnlinePoseEstimator.geodesic_distance + np.where(dists < neighborhood) over the 20 000 fine poses for seeded
      previous poses (executed from the imported reference class);
    * the scoring lines 67-76 executed verbatim on CPU tensors: mask_scores with the (30, 30) bilinear resize of
      OR(template masks, proposal mask), with a normalised query (frames > 0) and with the RAW coarse query feature
      (first frame: online_pose_estimator.py:41,50 never normalises it)."""
    import torch.nn.functional as F
    from einops import einsum
    ope = refimport.import_reference("src.pipeline.estimators.online_pose_estimator")
    pe = refimport.import_reference("src.pipeline.estimators.pose_estimator")
    fine = np.array(pe.DinoPoseEstimator.generate_poses(20000))
    rng = np.random.default_rng(5)
    prevs = []
    for i in range(6):
        qm, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(qm) < 0:
            qm[:, 0] = -qm[:, 0]
        T = np.eye(4); T[:3, :3] = qm; T[:3, 3] = rng.normal(size=3)
        prevs.append(T)
    prevs.append(fine[1234].copy())                       # exactly a fine pose: distance 0 to itself
    store = {"prev_poses": np.array(prevs), "fine_sha": np.array(sha(fine))}
    for i, T in enumerate(prevs):
        d = ope.DinoOnlinePoseEstimator.geodesic_distance(fine[:, :3, :3], T)      # as called at line 55
        for nb in (15, 5):
            close = np.where(d < nb)[0]
            store[f"close_{i}_{nb}"] = close.astype(np.int32)
        store[f"dist_{i}"] = d[store[f"close_{i}_15"]]
    # ---- scoring lines, verbatim
    feats_fine_template, query_raw, masks_fine_template, proposal_mask = online_fine_inputs()
    signature = 'b n d, b n d -> b n'
    for tag, query_feat in (("norm", F.normalize(query_raw, dim=-1)), ("raw", query_raw)):
        scores = einsum(query_feat, F.normalize(feats_fine_template, dim=-1), signature)
        masks = torch.logical_or(masks_fine_template, proposal_mask[None]).to(torch.float16)
        n_views = feats_fine_template.shape[0]
        masks = F.interpolate(masks[None].float(), size=(30, 30), mode='bilinear').reshape(n_views, 900)
        scores_m = (scores * masks).sum(dim=-1) / masks.sum(dim=-1)
        scores_u = einsum(query_feat, F.normalize(feats_fine_template, dim=-1), signature).mean(dim=-1)
        store[f"masked_{tag}"] = scores_m.float().numpy()
        store[f"plain_{tag}"] = scores_u.float().numpy()
        store[f"argmax_masked_{tag}"] = torch.argmax(scores_m).numpy()
        store[f"argmax_plain_{tag}"] = torch.argmax(scores_u).numpy()
    store["weights"] = masks.numpy()
    np.savez_compressed(OUT / "online_fine.npz", **store)


def template_store_case(n_views):
    """Seeded renders for the template-store fixture (shared with tests/test_template_store.py)."""
    from freepose_b200.pipeline.utils import generate_poses
    from freepose_b200.synthetic import synthetic_mesh
    from oracle import raster as R
    out = {}
    for j, mesh_id in enumerate(("mesh_a_1", "b2")):
        m = synthetic_mesh(7 + j, subdivisions=1 + j)
        poses = np.array(generate_poses(600))[:n_views]
        if j == 1:
            poses[3, 2, 3] = 40.0            # a view with a tiny mask: the reader forces the centre square
        out[mesh_id] = R.render_mesh(m, poses, 600, 600, 210, 210, 420, msaa=1)
    return out


def golden_template_store():
    """Shards written by THIS repository's TemplateShardWriter, read back by the reference's OWN WebTemplateDataset
    (src/dataloader/template.py, unmodified; it hard-codes 600 views per mesh)."""
    import tempfile
    from freepose_b200.pipeline.template_store import TemplateShardWriter
    tm = refimport.import_reference("src.dataloader.template")
    d = Path(tempfile.mkdtemp())
    with TemplateShardWriter(d) as w:
        for mesh_id, (rgb, depth) in template_store_case(600).items():
            w.write_mesh(mesh_id, rgb, depth)
    (d / "list.csv").write_text("model_name\nmesh_a_1\nb2\n")
    ds = tm.WebTemplateDataset(d.as_posix(), (d / "list.csv").as_posix(), crop=False)
    store = {}
    for name in ("mesha1", "b2"):
        o = ds.get_template_by_name(name)
        assert o["templates"].shape == (600, 3, 420, 420)
        for k in ("templates", "masks", "depths"):
            store[f"{name}_{k}_sha40"] = sha(o[k][:40].numpy())
        store[f"{name}_tar"], store[f"{name}_intrinsic"] = o["tar_file"], o["intrinsic"].numpy()
        store[f"{name}_mask_counts40"] = o["masks"][:40].sum((1, 2)).numpy()
    np.savez_compressed(OUT / "template_store.npz", **store)


if __name__ == "__main__":
    assert refimport.available(), "/root/reference is required to mint fixtures"
    torch.set_num_threads(1)
    golden_poses()
    golden_geometry()
    golden_crop()
    golden_dino_forward()
    golden_score()
    golden_retrieval()
    golden_refiner()
    golden_template_store()
    golden_online_fine()
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)
