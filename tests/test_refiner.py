"""Refiner pose-confidence pass (reference tracking_refiner.py:45-100,
refiner_utils.py:92-176).  CPU: the oracle restatement and the product's host logic against tests/golden/refiner.npz,
minted by the reference's own TrackingRefiner / refiner_utils (tests/golden/make_golden.py).  GPU: the kernels against
torchvision / cv2 / the oracle, and the whole pass against the reference fixture."""
import hashlib

import numpy as np
import pytest
import torch

bf = torch.bfloat16


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def sd_b2():
    from freepose_b200.vit_weights import VITB14_REG, synthetic_state_dict
    return synthetic_state_dict(VITB14_REG, seed=3, depth=2)


# ----------------------------------------------------------------------------------------------- CPU: oracle + host logic
def test_oracle_pose_confidence_equals_reference_fixture(golden, sd_b2):
    """oracle/refiner.py (fp32 mode) reproduces the reference class's confidence map, crop box, cropped intrinsics and
    roi_align crop on the seeded scene."""
    from oracle import refiner as OR
    g = golden["refiner"]
    mesh, frame, K, T = OR.synthetic_case(0)
    assert sha(frame) == str(g["frame_0_sha"]) and np.array_equal(T, g["T_0"])
    photo = torch.from_numpy(frame.astype(np.float32) / 255).permute(2, 0, 1).contiguous()     # ToTensor
    conf, parts = OR.OracleRefiner(sd_b2, "fp32").pose_confidence(mesh, photo, K, T)
    np.testing.assert_array_equal(parts["boxes"], g["bbox_0"])
    np.testing.assert_array_equal(parts["new_K"], g["new_K_0"])
    assert sha(parts["crop"]) == str(g["crop_0_sha"])
    np.testing.assert_allclose(conf, g["conf_0"], rtol=0, atol=2e-6)
    assert (conf > 0).sum() > 300 and conf.max() > 0.9


def test_host_crop_geometry_and_threshold_equal_reference(golden):
    from freepose_b200.pipeline import refiner_utils as RU
    from oracle import refiner as OR
    g = golden["refiner"]
    pts, Ts, Kt = (torch.from_numpy(g[k]) for k in ("ci_pts", "ci_Ts", "ci_K"))
    boxes = RU.crop_boxes(Ts, pts, Kt, 64, 48)
    np.testing.assert_array_equal(boxes.numpy(), g["ci_boxes"])
    np.testing.assert_array_equal(RU.update_K_with_crop(Kt, boxes, 64, 48).numpy(), g["ci_newK"])
    np.testing.assert_array_equal(OR.crop_boxes(Ts, pts, Kt, 64, 48).numpy(), g["ci_boxes"])
    np.testing.assert_array_equal(OR.roi_crops(torch.from_numpy(g["ci_img"]), boxes, 48, 64).numpy(), g["ci_crops"])
    # histogram threshold: product method (no GPU needed for it) and oracle vs the reference's value
    from freepose_b200.pipeline.estimators.tracking_refiner import TrackingRefiner
    thr = TrackingRefiner._get_threshold_for_confidence(None, g["sim"])
    assert thr == g["sim_thr"] == OR.threshold_for_confidence(g["sim"])
    # the private RandomState draws the indices the reference gets from np.random.seed(42); np.random.choice
    np.random.seed(42)
    assert np.array_equal(np.random.choice(np.arange(642), 100), np.random.RandomState(42).choice(np.arange(642), 100))


def test_refiner_has_no_cpu_fallback():
    from freepose_b200.pipeline.estimators.tracking_refiner import TrackingRefiner
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            TrackingRefiner(weights={})
    with pytest.raises((RuntimeError, ValueError)):
        TrackingRefiner(dino_device="cpu", weights={})


# ----------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_roi_align_equals_torchvision(lib, golden):
    from freepose_b200 import ops
    g = golden["refiner"]
    img = torch.from_numpy(g["ci_img"])
    boxes = torch.from_numpy(g["ci_boxes"])
    got = ops.roi_align(img.cuda(), boxes.cuda(), 48, 64, 2).cpu().numpy()
    np.testing.assert_allclose(got, g["ci_crops"], rtol=0, atol=1e-6)        # the reference's own crops
    assert np.mean(got == g["ci_crops"]) > 0.99
    # boxes hanging over every image border, degenerate (< 1 px) and adaptive sampling (sampling_ratio = 0)
    import torchvision
    torch.manual_seed(0)
    img = torch.rand(3, 77, 131)
    boxes = torch.tensor([[-20.5, -10.2, 60.0, 50.0], [100.0, 40.0, 180.0, 99.0], [30.0, 30.0, 30.2, 30.4],
                          [0.0, 0.0, 131.0, 77.0], [5.5, 6.5, 100.25, 70.75]])
    for sr in (2, 0, 3):
        want = torchvision.ops.roi_align(img[None], torch.cat([torch.zeros(5, 1), boxes], 1), (23, 31), sampling_ratio=sr)
        got = ops.roi_align(img.cuda(), boxes.cuda(), 23, 31, sr).cpu()
        assert torch.allclose(got, want, rtol=0, atol=1e-6), (sr, (got - want).abs().max())


@pytest.mark.gpu
def test_depth_mask_cubic_equals_cv2(lib):
    import cv2
    from freepose_b200 import ops
    rng = np.random.default_rng(0)
    d = np.zeros((3, 520, 520), np.float32)
    yy, xx = np.mgrid[:520, :520]
    d[0][(yy - 250) ** 2 + (xx - 270) ** 2 < 150 ** 2] = 1.3
    d[1][rng.random((520, 520)) > 0.5] = 0.7                       # salt and pepper: every cubic tap matters
    d[2][:260] = 2.0
    d[2][517:] = 5.0                                               # outside the 518 x 518 image: must be ignored
    got = ops.depth_mask_cubic(torch.from_numpy(d).cuda(), 37, res=518).cpu().numpy()
    for i in range(3):
        want = cv2.resize((d[i, :518, :518] > 0).astype(np.float32), (37, 37), interpolation=cv2.INTER_CUBIC) > 0.5
        assert np.array_equal(got[i], want), i
    # other scales (non-integer ratio, border taps): contiguous input
    d2 = (rng.random((2, 224, 224)) > 0.6).astype(np.float32)
    got = ops.depth_mask_cubic(torch.from_numpy(d2).cuda(), 30).cpu().numpy()
    for i in range(2):
        want = cv2.resize(d2[i], (30, 30), interpolation=cv2.INTER_CUBIC)
        near = np.abs(want - 0.5) < 1e-4                           # cv2's SIMD summation order is not specified
        assert np.array_equal(got[i][~near], (want > 0.5)[~near])


@pytest.mark.gpu
def test_patch_cosine(lib):
    from freepose_b200 import ops
    torch.manual_seed(0)
    a = torch.randn(5, 1369, 768).to(bf)
    b = (a.float() * 0.7 + 0.5 * torch.randn(5, 1369, 768)).to(bf)
    mask = torch.rand(5, 1369) > 0.4
    got = ops.patch_cosine(a.cuda(), b.cuda(), mask.cuda()).cpu()
    fa, fb = a.float(), b.float()
    fa = fa / torch.linalg.norm(fa, dim=-1, keepdim=True)
    fb = fb / torch.linalg.norm(fb, dim=-1, keepdim=True)
    want = (fa * fb).sum(-1) * mask.float()
    assert torch.allclose(got, want, rtol=0, atol=2e-6)
    assert torch.equal(got == 0, ~mask)
    assert torch.allclose(ops.patch_cosine(a.cuda(), a.cuda()).cpu(), torch.ones(5, 1369), atol=2e-6)


@pytest.mark.gpu
def test_refiner_render_setup_bit_exact(lib):
    """ambient 5 / znear 1e-4 / zfar 9999 / back-face culling / one camera per view (tracking_refiner.py:31-45) against
    the C restatement, 518 px views inside 520 px targets."""
    from freepose_b200 import ops
    from oracle import raster as R
    from oracle import refiner as OR
    mesh, _, K, T = OR.synthetic_case(0)
    Ts = np.stack([T, OR.synthetic_case(1)[3]])
    ks = np.array([[1197.25, 1197.25, 232.4, 294.3], [903.5, 911.0, 250.0, 262.5]], np.float32)
    want_rgb, want_depth = R.render_mesh(mesh, Ts, 1, 1, 0, 0, 520, msaa=4, cull=True, ambient=5.0, znear=1e-4, zfar=9999.0,
                                         view_k=ks)
    rgb, depth = ops.rasterize_mesh(mesh, torch.from_numpy(Ts).float().cuda(), 1.0, 1.0, 0.0, 0.0, 520, msaa=4,
                                    cull_backfaces=True, ambient=5.0, znear=1e-4, zfar=9999.0,
                                    view_k=torch.from_numpy(ks).cuda())
    assert np.array_equal(rgb.cpu().numpy(), want_rgb) and np.array_equal(depth.cpu().numpy(), want_depth)
    assert (want_depth[0] > 0).sum() > 5000 and not np.array_equal(want_depth[0], want_depth[1])


@pytest.mark.gpu
def test_pose_confidence_end_to_end_vs_reference_fixture(lib, golden, sd_b2):
    """TrackingRefiner.pose_confidence / n_inliers_per_pose on the device against (a) the confidence maps the reference's
    own class produced in fp32 and (b) the oracle in the engine's bf16 rounding contract."""
    from freepose_b200.pipeline.estimators.tracking_refiner import TrackingRefiner
    from oracle import refiner as OR
    g = golden["refiner"]
    ref = TrackingRefiner(weights=sd_b2)
    cases = [OR.synthetic_case(i) for i in range(2)]
    mesh, K = cases[0][0], cases[0][2]
    frames, Ts = [c[1] for c in cases], [c[3] for c in cases]
    # stage: crop box, intrinsics, crop pixels
    crop, bbox, new_K = ref._crop_image(mesh, frames[0], K, Ts[0])
    np.testing.assert_array_equal(bbox.numpy(), g["bbox_0"])
    np.testing.assert_array_equal(new_K.numpy(), g["new_K_0"])
    np.testing.assert_allclose(crop.cpu().numpy()[:, ::37, ::37], g["crop_0_rows"], rtol=0, atol=1e-6)
    conf = ref.pose_confidences(mesh, frames, K, Ts).cpu().numpy()
    assert conf.shape == (2, 37, 37)
    oc = OR.OracleRefiner(sd_b2, "contract")
    for i in range(2):
        photo = torch.from_numpy(frames[i].astype(np.float32) / 255).permute(2, 0, 1).contiguous()
        want, parts = oc.pose_confidence(mesh, photo, K, Ts[i])
        assert np.array_equal(conf[i] != 0, parts["mask"]), "validity mask differs"
        assert np.abs(conf[i] - want).max() < 5e-3, np.abs(conf[i] - want).max()              # same rounding contract
        assert np.abs(conf[i] - g[f"conf_{i}"]).max() < 2e-2, np.abs(conf[i] - g[f"conf_{i}"]).max()   # reference fp32
    np.testing.assert_array_equal(ref.pose_confidence(mesh, frames[1], K, Ts[1]), conf[1])    # batch invariance
    counts, thr = ref.n_inliers_per_pose(mesh, frames, K, Ts)
    want_thr = float(g["thr"])
    assert abs(thr - want_thr) < 2e-2
    want_counts = np.array([(g[f"conf_{i}"] > want_thr).sum() for i in range(2)])
    assert np.all(np.abs(counts - want_counts) <= 0.1 * want_counts + 5), (counts, want_counts)
