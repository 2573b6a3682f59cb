"""CPU: host-side logic -- sharding, the score all-gather over gloo (world_size 2), synthetic inputs,
product paths failing loudly without CUDA."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def test_shard_bounds_cover_exactly():
    from freepose_b200.distributed import shard_bounds
    for n in (0, 1, 7, 520, 521, 600):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == -(-n // w) or n == 0
    assert shard_bounds(520, 3, 8) == (195, 260)


def test_stable_topk_tie_rule():
    from freepose_b200.distributed import stable_topk_host
    s = torch.tensor([0.5, 0.75, 0.75, 0.25, 0.75])
    idx, vals = stable_topk_host(s, 3)
    assert idx.tolist() == [1, 2, 4] and vals.tolist() == [0.75, 0.75, 0.75]


def test_synthetic_mesh_is_normalised_like_the_reference():
    from freepose_b200.synthetic import camera_for, synthetic_mesh
    m = synthetic_mesh(0, subdivisions=3)
    v = np.asarray(m.vertices)
    # resize_meshes.py:18-23 normalisation (bbox centred, max half extent 1) then x0.25
    assert np.allclose((v.min(0) + v.max(0)) / 2, 0, atol=1e-12) and np.isclose(np.abs(v).max(), 0.25)
    assert m.faces.shape == (1280, 3) and m.vertex_colors.dtype == np.uint8 and m.vertex_colors.max() < 128
    assert camera_for(420) == (600.0, 600.0, 210.0, 210.0) and camera_for(224)[0] == 320.0
    # closed manifold: every edge shared by exactly two faces
    e = np.sort(np.concatenate([m.faces[:, [0, 1]], m.faces[:, [1, 2]], m.faces[:, [2, 0]]]), axis=1)
    _, counts = np.unique(e, axis=0, return_counts=True)
    assert np.all(counts == 2)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    from freepose_b200 import ops
    from freepose_b200.vit_engine import ViTEngine
    from freepose_b200.vit_weights import synthetic_state_dict
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ViTEngine(synthetic_state_dict(depth=1))
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ops.score_topk(torch.zeros(2, 4, 1024, dtype=torch.bfloat16), torch.zeros(4, 1024, dtype=torch.bfloat16))


def test_missing_library_is_an_error(tmp_path, monkeypatch):
    from freepose_b200 import _lib
    monkeypatch.setenv("FREEPOSE_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


WORKER = textwrap.dedent("""
    import sys, torch, numpy as np
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from freepose_b200.distributed import ScoreGather, init_from_env, shard_bounds, stable_topk_host
    rank, local, world = init_from_env("gloo")
    n = 37                                   # not divisible by the world size: padding slots must never win
    g = torch.Generator().manual_seed(0)
    all_scores = torch.randn(n, generator=g)
    all_scores[5] = all_scores[30] = all_scores.max() + 1.0      # a tie across the two shards
    lo, hi = shard_bounds(n, rank, world)
    sg = ScoreGather(n, world, "cpu")
    sg.local_view(rank)[: hi - lo] = all_scores[lo:hi]           # "score kernel writes into the gather buffer"
    got = sg.gather(rank)
    assert torch.equal(got, all_scores), (rank, got, all_scores)
    idx, vals = stable_topk_host(got, 3)
    assert idx[:2].tolist() == [5, 30]
    gathered = [None] * world
    dist.all_gather_object(gathered, idx.tolist())
    assert all(x == gathered[0] for x in gathered)               # every rank picks the identical winners
    dist.barrier()
    sys.stdout.write("rank" + str(rank) + "-ok" + chr(10)); sys.stdout.flush()
""")


def test_score_allgather_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % str(ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank0-ok" in r.stdout and "rank1-ok" in r.stdout


def test_rle_matches_the_reference_format():
    """Uncompressed column-major RLE of the proposal JSON (reference src/pipeline/utils.py:62 ->
    sam2.utils.amg.mask_to_rle_pytorch); the expected vectors were produced by that function."""
    from freepose_b200.pipeline.proposals import mask_to_rle, rle_to_mask
    m = np.array([[1, 0], [1, 1]], dtype=bool)
    assert mask_to_rle(m) == {"size": [2, 2], "counts": [0, 2, 1, 1]}
    m = np.zeros((3, 4), dtype=bool)
    m[1, 2] = m[2, 2] = True
    assert mask_to_rle(m) == {"size": [3, 4], "counts": [7, 2, 3]}
    rng = np.random.default_rng(0)
    for _ in range(20):
        x = rng.random((9, 7)) > 0.5
        assert np.array_equal(rle_to_mask(mask_to_rle(x)), x)
    assert mask_to_rle(np.zeros((2, 2), bool)) == {"size": [2, 2], "counts": [4]}


def test_src_overlay_exposes_the_reference_names():
    """`from src.pipeline... import X` (the imports the reference scripts use) resolve to the B200 classes."""
    import importlib
    for mod, name in (("src.pipeline.retrieval.dino", "DINOv2FeatureExtractor"),
                      ("src.pipeline.retrieval.renderer", "MeshRenderer"),
                      ("src.pipeline.estimators.pose_estimator", "DinoPoseEstimator"),
                      ("src.pipeline.estimators.online_pose_estimator", "DinoOnlinePoseEstimator"),
                      ("src.pipeline.utils", "Proposals"), ("src.pipeline.utils", "get_z_from_pointcloud"),
                      ("src.utils.bbox_utils", "CropResizePad")):
        obj = getattr(importlib.import_module(mod), name)
        assert obj.__module__.startswith("freepose_b200."), (mod, name, obj.__module__)


def test_feature_cache_lru_evict_and_reload(tmp_path):
    """Row K (reference pose_estimator.py:38-77): RAM LRU of `cache_size` meshes kept on the HOST, oldest entry evicted
    to `<cache_dir>/<name>.pth`, reloaded from disk on the next request, `save_all` writes through, the cache
    directory is removed with the estimator.  Runs on the CPU with a counting stand-in for the ViT."""
    from freepose_b200.pipeline.estimators.pose_estimator import DinoPoseEstimator

    class FakeEngine:
        device = torch.device("cpu")

    class FakeExtractor(torch.nn.Module):
        engine = FakeEngine()

    calls = []

    class Est(DinoPoseEstimator):
        def _extract_features(self, proposals, layer=22, batch_size=128):
            calls.append(float(proposals.flatten()[0]))
            return (proposals.flatten()[0] * torch.ones(len(proposals), 4, 1024)).to(torch.bfloat16)

    cdir = tmp_path / "cache"
    est = Est(n_poses=4, cache_size=2, cache_dir=str(cdir), feature_extractor=FakeExtractor(), resolution=224,
              device_cache_bytes=3 * 4 * 1024 * 2)                      # room for ONE entry on the "device"
    td = lambda name, v: {"model_name": name, "templates": torch.full((3, 3, 28, 28), float(v))}
    fa = est._get_template_features(td("a", 1))
    fb = est._get_template_features(td("b", 2))
    assert calls == [1.0, 2.0] and list(est.feature_cache) == ["a", "b"] and not list(cdir.glob("*.pth"))
    assert list(est._device_cache) == ["b"]                             # byte cap: only the most recent stays resident
    assert all(not t.is_cuda for t in est.feature_cache.values())       # the LRU itself is host memory
    # hit: no recompute, moves to the MRU end, value identical
    assert torch.equal(est._get_template_features(td("a", 1)), fa) and calls == [1.0, 2.0]
    assert list(est.feature_cache) == ["b", "a"]
    # third mesh: "b" (least recently used) is evicted to disk
    est._get_template_features(td("c", 3))
    assert list(est.feature_cache) == ["a", "c"] and [p.name for p in cdir.glob("*.pth")] == ["b.pth"]
    assert "b" not in est._device_cache
    # reload "b" from disk: no recompute, bit-identical, and now "a" is the one evicted
    fb2 = est._get_template_features(td("b", 2))
    assert calls == [1.0, 2.0, 3.0] and torch.equal(fb2.cpu(), fb.cpu()) and fb2.dtype == torch.bfloat16
    assert list(est.feature_cache) == ["c", "b"] and (cdir / "a.pth").exists()
    # save_all writes through immediately
    est.save_all = True
    est._get_template_features(td("d", 4))
    assert (cdir / "d.pth").exists() and torch.equal(torch.load(cdir / "d.pth"), est.feature_cache["d"])
    # reference __del__ (pose_estimator.py:76-77): the cache directory goes with the estimator
    est.__del__()
    assert not cdir.exists()
