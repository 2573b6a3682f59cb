"""Template store (reference scripts/render_templates.py:49-72, src/dataloader/template.py):
shards written here are read back (a) by this repository's WebTemplateDataset and compared with what the reference's own
reader returned for the same shards (tests/golden/template_store.npz), (b) on the GPU, rendered by the device rasteriser
and fed through the estimator's template-dict contract."""
import hashlib
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_reader_equals_reference_reader_on_our_shards(golden, tmp_path):
    from make_golden import template_store_case
    from freepose_b200.pipeline.template_store import TemplateShardWriter, WebTemplateDataset, collate_fn
    g = golden["template_store"]
    with TemplateShardWriter(tmp_path) as w:
        for mesh_id, (rgb, depth) in template_store_case(40).items():
            w.write_mesh(mesh_id, rgb, depth)
    (tmp_path / "list.csv").write_text("model_name\nmesh_a_1\nb2\n")
    ds = WebTemplateDataset(tmp_path.as_posix(), (tmp_path / "list.csv").as_posix(), crop=False, n_views=40)
    assert len(ds) == 2
    outs = []
    for name in ("mesha1", "b2"):
        o = ds.get_template_by_name(name)
        outs.append(o)
        assert o["templates"].shape == (40, 3, 420, 420) and o["templates"].dtype == torch.float32
        assert o["masks"].dtype == torch.bool and o["depths"].dtype == torch.float32
        for k in ("templates", "masks", "depths"):
            assert sha(o[k].numpy()) == str(g[f"{name}_{k}_sha40"]), (name, k)
        assert o["tar_file"] == str(g[f"{name}_tar"]) and o["model_name"] == name
        assert np.array_equal(o["intrinsic"].numpy(), g[f"{name}_intrinsic"])
        assert np.array_equal(o["masks"].sum((1, 2)).numpy(), g[f"{name}_mask_counts40"])
    assert outs[1]["masks"][3, 105:315, 105:315].all() and int(outs[1]["masks"][3].sum()) == 210 * 210   # fallback square
    mm = outs[0]["depths"].double() * 1000
    assert float(outs[0]["depths"].max()) < 2 and float((mm - mm.round()).abs().max()) < 1e-3            # whole mm
    b = collate_fn(outs + [{"templates": None}])
    assert b["templates"].shape[0] == 80 and b["model_name"] == ["mesha1", "b2"]
    # shard roll-over: 10 meshes per shard like the reference (idx // 10)
    w = TemplateShardWriter(tmp_path / "many", meshes_per_shard=10)
    tiny_rgb, tiny_d = np.zeros((1, 4, 4, 3), np.uint8), np.ones((1, 4, 4), np.float32)
    for i in range(12):
        w.write_mesh(f"m_{i}", tiny_rgb, tiny_d)
    w.close()
    assert sorted(p.name for p in (tmp_path / "many").iterdir()) == ["shard-000000.tar", "shard-000001.tar"]


@pytest.mark.gpu
def test_render_templates_to_store_and_estimate_from_it(lib, tmp_path):
    """render_templates (device rasteriser) -> shards -> WebTemplateDataset(crop=True) -> DinoPoseEstimator.forward with
    the template dict: the same best hypothesis as rendering online (forward_mesh)."""
    from freepose_b200.pipeline.estimators.pose_estimator import DinoPoseEstimator
    from freepose_b200.pipeline.template_store import WebTemplateDataset, render_templates
    from freepose_b200.synthetic import synthetic_mesh
    from freepose_b200.vit_weights import synthetic_state_dict
    from oracle.pipeline import synthetic_query
    n = 12
    unit = synthetic_mesh(0, subdivisions=3, scale=1.0)
    render_templates({"obj_000_1": unit}, tmp_path, n_poses=n, resolution=420)
    (tmp_path / "list.csv").write_text("model_name\nobj_000_1\n")
    ds = WebTemplateDataset(tmp_path.as_posix(), (tmp_path / "list.csv").as_posix(), n_views=n)
    entry = ds.get_template_by_name("obj0001")
    assert entry["templates"].shape == (n, 3, 420, 420) and entry["depths"].shape == (n, 420, 420)
    est = DinoPoseEstimator(n_poses=n, cache_size=0, cache_dir=str(tmp_path / "cache"),
                            weights=synthetic_state_dict(seed=0, depth=2), resolution=420)
    mesh = synthetic_mesh(0, subdivisions=3)            # = unit scaled by 0.25
    query, _ = synthetic_query(mesh, 420, seed=1)
    K = np.array([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
    bbox = torch.tensor([200.0, 150.0, 330.0, 290.0])
    a = est.forward(query, entry, K, bbox, 0.3, layer=2)
    b = est.forward_mesh(query, mesh, K, bbox.numpy(), 0.3, layer=2)
    assert int(np.argmax(a["scores"])) == 0 and len(a["TCO"]) == 3
    sa, sb = np.asarray(a["scores"], dtype=np.float64), np.asarray(b["scores"], dtype=np.float64)
    assert np.allclose(sa, sb, atol=2e-2), (sa, sb)     # stored depth is truncated to mm; RGB is identical
    np.testing.assert_allclose(a["TCO"][0][:3, :3], b["TCO"][0][:3, :3], atol=1e-12)
