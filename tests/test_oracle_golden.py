"""CPU: the oracle restatements against fixtures minted from the reference's own code
(tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import torch

from oracle import crop as ocrop
from oracle import score as oscore


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_generate_poses_matches_reference(golden):
    from freepose_b200.pipeline.utils import generate_poses
    g = golden["poses"]
    assert np.array_equal(np.array(generate_poses(42)), g["p42"])
    p600 = np.array(generate_poses(600))
    assert np.array_equal(p600[:8], g["p600_first8"]) and np.array_equal(p600[-8:], g["p600_last8"])
    assert sha(p600) == str(g["p600_sha"])
    p20k = np.array(generate_poses(20000))
    assert sha(p20k) == str(g["p20000_sha"])
    assert np.array_equal(p20k[[0, 1, 7777, 19999]], g["p20000_rows"])
    # structure: proper rotations at distance 1.1
    R = p600[:, :3, :3]
    assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3), atol=1e-12)
    assert np.allclose(np.linalg.det(R), 1.0) and np.all(p600[:, 2, 3] == 1.1)


def test_geometry_matches_reference(golden):
    from freepose_b200.pipeline import utils as U
    g = golden["geometry"]
    for i in range(4):
        c = {k[len(f"c{i}_"):]: g[k] for k in g.files if k.startswith(f"c{i}_")}
        pc = U.depthmap_to_pointcloud(c["depth"], c["Kt"])
        assert pc.shape[0] == int(c["n_points"]) and sha(pc) == str(c["pc_sha"])
        s = float(c["est_scale"])
        pcc = pc.copy(); m = pcc.mean(axis=0); pcc -= m; pcc /= 0.25; pcc *= s; pcc += m
        assert np.array_equal(U.get_z_from_pointcloud(c["bbox"], pcc, c["Kq"], c["T0"]), c["tco_coarse"])
        pcf = pc.copy(); pcf /= 0.25; pcf *= s
        assert np.array_equal(U.get_z_from_pointcloud(c["bbox"], pcf, c["Kq"], c["T0"]), c["tco_fine"])
        assert np.array_equal(U.mask_to_bbox(c["depth"] > 0), c["bbox_of_mask"])
        # the O(1) extents formulation the engine uses gives the same TCO (to fp64 round-off of the mean)
        ext = np.array([pc[:, 0].min(), pc[:, 0].max(), pc[:, 1].min(), pc[:, 1].max(), pc[:, 0].sum(),
                        pc[:, 1].sum(), pc[:, 2].sum(), pc.shape[0]])
        for recentre, want in ((True, c["tco_coarse"]), (False, c["tco_fine"])):
            dx, dy = U.rescaled_extents(ext, s, recentre)
            got = U.tco_from_extents(c["bbox"], dx, dy, c["Kq"], c["T0"])
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


def test_crop_resize_pad_matches_reference(golden):
    g = golden["crop"]
    rng = np.random.default_rng(int(g["rng_seed"]))
    i = 0
    while f"c{i}_spec" in g.files:
        H, W, T = [int(v) for v in g[f"c{i}_spec"]]
        ext = float(g[f"c{i}_ext"])
        ext = int(ext) if ext == 0 else ext
        boxes = g[f"c{i}_boxes"]
        imgs = rng.random((len(boxes), 3, H, W)).astype(np.float32)  # same stream as the generator
        assert sha(imgs) == str(g[f"c{i}_seed_imgs_sha"])
        out = ocrop.crop_resize_pad(imgs, boxes, T, bbox_extend=ext, orig_size=(H, W))
        assert sha(out) == str(g[f"c{i}_out_sha"]), f"crop case {i}"
        if f"c{i}_out" in g.files:
            assert np.array_equal(out, g[f"c{i}_out"])
        i += 1
    assert i == 6


def test_extend_boxes_product_matches_oracle():
    from freepose_b200.pipeline.bbox_utils import extend_boxes
    rng = np.random.default_rng(0)
    for _ in range(200):
        w, h = 640, 480
        x1, y1 = int(rng.integers(0, 600)), int(rng.integers(0, 440))
        x2, y2 = int(rng.integers(x1 + 1, 640)), int(rng.integers(y1 + 1, 480))
        for ext in (0, 0.2, 0.1):
            got = extend_boxes(torch.tensor([[x1, y1, x2, y2]]), ext, w, h)[0].tolist()
            assert got == list(ocrop.extend_box([x1, y1, x2, y2], ext, w, h))


def test_dino_forward_slices_match_reference(golden):
    """oracle forward_features + token slices == the reference's DINOv2FeatureExtractor.forward output."""
    from freepose_b200.vit_weights import synthetic_state_dict
    from oracle.pipeline import reference_normalize
    from oracle.vit import OracleViT
    g = golden["dino_forward"]
    vit = OracleViT(synthetic_state_dict(seed=3, depth=2)).float()
    imgs = torch.from_numpy(g["imgs"])
    with torch.no_grad():
        for layer in (1, 2):
            x = vit.forward_features(reference_normalize(imgs), layer)
            np.testing.assert_allclose(x[:, 0].numpy(), g[f"cls_{layer}"], rtol=0, atol=2e-5)  # fp32 round-off (BLAS thread count)
            np.testing.assert_allclose(x[:, 1:5].numpy(), g[f"reg_{layer}"], rtol=0, atol=2e-5)  # fp32 round-off (BLAS thread count)
            np.testing.assert_allclose(x[:, 5:].numpy(), g[f"patch_{layer}"], rtol=0, atol=2e-5)  # fp32 round-off (BLAS thread count)


def test_score_oracles_match_reference_lines(golden):
    g = golden["score"]
    ft = torch.from_numpy(g["feats_t"]).view(torch.bfloat16)
    fq = torch.from_numpy(g["feat_q"]).view(torch.bfloat16)
    ref_scores = g["scores"]
    # the reference lines re-executed here reproduce the fixture bit for bit
    assert np.array_equal(oscore.reference_scores(ft, fq).float().numpy(), ref_scores)
    # the engine-order restatement: identical values on this fixture (at most 1 bf16 ulp apart in general)
    eng = oscore.engine_order_scores(ft, fq)
    assert np.array_equal(eng, ref_scores)
    idx, vals = oscore.stable_topk(eng, 3)
    assert np.array_equal(vals, g["top_scores"])             # values are tie-order independent
    assert eng[3] == eng[5]                                    # the planted exact tie
    assert int(idx[0]) == int(g["argmax"]) or eng[int(g["argmax"])] == vals[0]
    assert float(g["maxval"]) == float(vals[0])
    w = torch.from_numpy(g["masks"])
    np.testing.assert_allclose(oscore.engine_order_scores(ft, fq, weights=w), g["weighted"], rtol=2e-6)
    np.testing.assert_allclose(oscore.reference_scores(ft, fq, weights=w).numpy(), g["weighted"], rtol=1e-6)


def test_ffa_oracles_agree():
    torch.manual_seed(0)
    feats = torch.randn(5, 16, 1024).to(torch.bfloat16)
    rng = np.random.default_rng(0)
    masks = rng.random((5, 56, 56)) > 0.995
    masks[4] = False  # empty mask -> NaN row, which the reference detects and skips
    ref = oscore.ffa_reference(feats, masks)
    eng, counts = oscore.ffa_engine_order(feats, masks)
    assert counts[4] == 0 and np.isnan(eng[4]).all() and np.isnan(ref[4]).all()
    # sequential vs ATen summation order: identical up to rare single bf16-ulp flips
    d = np.abs(ref[:4] - eng[:4])
    assert (d > 0).mean() < 0.01 and np.all(d <= np.abs(ref[:4]) * 2 ** -7 + 1e-30)


# ------------------------------------------------------------------------------------------- mesh retrieval
def _tie_consistent(cand, vals, ref_cand, ref_vals):
    """Two top-k selections of the same score vector may differ only inside the tie at the cut-off value."""
    assert np.array_equal(np.sort(vals), np.sort(ref_vals))             # same multiset of values
    cut = vals.min()
    ours, theirs = dict(zip(cand.tolist(), vals.tolist())), dict(zip(ref_cand.tolist(), ref_vals.tolist()))
    for m in set(ours) ^ set(theirs):
        assert ours.get(m, theirs.get(m)) == cut
    return set(ours) & set(theirs)


def test_retrieval_oracles_match_reference_lines(golden):
    from oracle import retrieval as R
    g = golden["retrieval"]
    case = R.synthetic_case(0)
    assert sha(case["db"]) == bytes(g["db_sha"]).hex()                  # the seeded inputs are the minted ones
    dbn = R.engine_normalize(case["db"])
    ref16 = torch.from_numpy(g["db_norm"]).view(torch.bfloat16).float().numpy()
    assert np.array_equal(dbn[:16], ref16)
    qn = R.engine_normalize(case["queries"])
    for topk in (0, 3, 10):
        best, score, cand, cs = R.engine_retrieve(dbn, case["fine"], qn, topk)
        assert np.array_equal(best, g[f"best_{topk}"])                  # retrieved mesh: bit-exact
        assert np.array_equal(score.astype(np.float64), g[f"score_{topk}"])
        for q in range(cand.shape[0]):
            if topk == 0:
                _tie_consistent(cand[q], cs[q], g["cand_0"][q], g["cand_scores_0"][q])
            else:  # candidates present in both selections carry identical fine scores
                coarse = R.engine_topk(R.engine_scan(dbn, qn[q:q + 1])[0], 100)[1]
                ours = dict(zip(cand[q].tolist(), cs[q].tolist()))
                theirs = dict(zip(g[f"cand_{topk}"][q].tolist(), g[f"cand_scores_{topk}"][q].tolist()))
                shared = set(ours) & set(theirs)
                assert len(shared) >= 90 and all(ours[m] == theirs[m] for m in shared), coarse[-1]
    # duplicated rows 3 / 7: identical scores, the lower index wins
    assert best[1] == 3 and 7 in cand[1]


def test_retrieval_softvote_matches_reference_lines(golden):
    from oracle import retrieval as R
    g = golden["retrieval"]
    case = R.synthetic_case(0)
    dbn = R.engine_normalize(case["db"])
    per_frame = []
    for fr in case["video"]:
        _, _, cand, cs = R.engine_retrieve(dbn, case["fine"], R.engine_normalize(fr), 3)
        per_frame.append((cand, cs))
    best, score, _ = R.engine_softvote_dense(per_frame, dbn.shape[0])
    assert np.array_equal(best, g["vote_best"])
    # torch.mean sums the frames in its own (cascade) order: values agree to fp32 round-off, the argmax exactly
    np.testing.assert_allclose(score, g["vote_score"], rtol=1e-6)


def test_numpy_order_mean_is_numpy_mean():
    from oracle.retrieval import numpy_order_mean
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 7, 8, 9, 15, 16, 17, 31, 64, 100, 128):
        for _ in range(10):
            v = rng.standard_normal(n).astype(np.float32)
            assert numpy_order_mean(v) == v.mean()


# ------------------------------------------------------------------------------------------- raster oracle: surfaces
def test_raster_oracle_uniform_texture_matches_closed_form():
    """A constant texture must come out as gamma(2 * (c/255)^2.2) wherever the mesh is hit (pyrender: srgb_to_linear on
    the texel, ambient (2,2,2), 1/2.2 gamma), for every mip level and filter weight."""
    from freepose_b200.pipeline.utils import Mesh, generate_poses
    from freepose_b200.synthetic import synthetic_textured_mesh
    from oracle import raster as R
    m = synthetic_textured_mesh(0, 2)
    poses = np.array(generate_poses(3))
    for c in (0, 37, 100, 180, 255):
        tex = np.full((16, 32, 3), c, np.uint8)
        rgb, depth = R.render_mesh(Mesh(m.vertices, m.faces, None, m.uv, tex), poses, 320, 320, 112, 112, 224, msaa=1)
        want = int(np.floor(255 * min(1.0, 2 * (c / 255) ** 2.2) ** (1 / 2.2) + 0.5))
        got = np.unique(rgb[depth > 0])
        assert len(got) == 1 and abs(int(got[0]) - want) <= 1, (c, got, want)


def test_mip_chain_and_trimesh_adapters():
    from types import SimpleNamespace as NS
    from freepose_b200.pipeline.utils import as_mesh, build_mip_chain
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(5, 12, 3), dtype=np.uint8)
    chain, levels = build_mip_chain(img)
    sizes = [(12, 5), (6, 2), (3, 1), (1, 1)]
    assert levels == len(sizes) and chain.size == sum(w * h * 4 for w, h in sizes)
    l1 = chain[12 * 5 * 4:12 * 5 * 4 + 6 * 2 * 4].reshape(2, 6, 4)
    want = (img[0:2, 0:2].astype(int).sum((0, 1)) + 2) >> 2
    assert np.array_equal(l1[0, 0, :3], want) and l1[0, 0, 3] == 255
    tri = NS(vertices=np.zeros((3, 3)), faces=np.array([[0, 1, 2]]),
             visual=NS(kind="texture", uv=np.zeros((3, 2)), material=NS(image=img)))
    m = as_mesh(tri)
    assert m.texture is img or np.array_equal(m.texture, img)
    assert m.uv.shape == (3, 2) and not m.is_point_cloud
    pc = as_mesh(NS(vertices=np.zeros((4, 3)), colors=np.zeros((0, 4))))
    assert pc.is_point_cloud and pc.vertex_colors is None
    vc = as_mesh(NS(vertices=np.zeros((3, 3)), faces=np.array([[0, 1, 2]]),
                    visual=NS(kind="vertex", vertex_colors=np.full((3, 4), 9, np.uint8))))
    assert vc.vertex_colors.shape == (3, 4) and vc.texture is None


def test_raster_oracle_bilinear_texture_on_a_facing_quad():
    """Independent float64 evaluation of the texture path (orientation v-up, REPEAT-free interior, bilinear on level 0
    when magnified, x^2.2 after filtering, ambient 2, 1/2.2 gamma) on a quad parallel to the image plane."""
    from freepose_b200.pipeline.utils import Mesh
    from oracle import raster as R
    rng = np.random.default_rng(5)
    tex = rng.integers(0, 256, size=(24, 24, 3), dtype=np.uint8)
    v = np.array([[-0.3, -0.3, 0.0], [0.3, -0.3, 0.0], [0.3, 0.3, 0.0], [-0.3, 0.3, 0.0]])
    uv = np.stack([(v[:, 0] + 0.3) / 0.6, (0.3 - v[:, 1]) / 0.6], axis=1)       # top edge of the quad (y = -0.3) is v = 1
    mesh = Mesh(v, np.array([[0, 1, 2], [0, 2, 3]]), None, uv, tex)
    pose = np.eye(4)[None].copy()
    pose[0, 2, 3] = 1.0
    rgb, depth = R.render_mesh(mesh, pose, 320.0, 320.0, 112.0, 112.0, 224, msaa=1)
    assert abs(float(depth[0, 100, 100]) - 1.0) < 1e-6 and depth[0, 5, 5] == 0
    ys, xs = np.mgrid[24:200, 24:200]                                            # interior of the 192 px quad
    X, Y = (xs + 0.5 - 112) / 320, (ys + 0.5 - 112) / 320
    tx, ty = (X + 0.3) / 0.6 * 24 - 0.5, (Y + 0.3) / 0.6 * 24 - 0.5
    x0, y0 = np.floor(tx).astype(int), np.floor(ty).astype(int)
    fx, fy = (tx - x0)[..., None], (ty - y0)[..., None]
    t = tex.astype(np.float64)
    c = (t[y0, x0] * (1 - fx) + t[y0, x0 + 1] * fx) * (1 - fy) + (t[y0 + 1, x0] * (1 - fx) + t[y0 + 1, x0 + 1] * fx) * fy
    want = np.floor(255 * np.minimum(1.0, 2 * (c / 255) ** 2.2) ** (1 / 2.2) + 0.5)
    got = rgb[0, 24:200, 24:200].astype(np.float64)
    assert np.abs(got - want).max() <= 1 and np.mean(got == want) > 0.97
    assert got.std() > 20                                                        # a real image, not a constant


def test_raster_oracle_clips_at_the_near_plane_like_an_analytic_ray_caster():
    """The restated spec must CLIP triangles that straddle the near plane or reach behind the camera (GL does), not drop
    them.  A ground quad from z = -1 (behind the camera) to z = 6: coverage and depth of every pixel against the closed
    form ray / plane intersection t = 0.3 / d_y, inside |X| < 2, 0.05 < t < 6 -- and vertex colours against the bilinear
    closed form on the quad's two triangles."""
    from oracle import raster as R
    res, f, c = 128, 150.0, 64.0
    V = np.array([[-2, 0.3, -1.0], [2, 0.3, -1.0], [2, 0.3, 6.0], [-2, 0.3, 6.0]], np.float32)
    F = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    col = np.array([[250, 10, 10], [10, 250, 10], [10, 10, 250], [250, 250, 10]], np.uint8)
    rgb, depth = R.render(V, F, col, np.eye(4)[None], f, f, c, c, res, msaa=1)
    py, px = np.mgrid[0:res, 0:res]
    dx, dy = (px + 0.5 - c) / f, (py + 0.5 - c) / f
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(dy > 0, 0.3 / dy, np.inf)
    X = t * dx
    want = (np.abs(X) < 2) & (t < 6) & (t > 0.05)
    got = depth[0] > 0
    assert np.array_equal(got, want) and want.sum() > 5000
    assert np.max(np.abs(depth[0][want] - t[want]) / t[want]) < 1e-6
    # colours: barycentric in the plane (exact for perspective-correct interpolation), x ambient 2, gamma 1/2.2
    P = np.stack([X[want], t[want]], axis=1)                      # (X, Z) on the plane
    A, B, C, D = V[0, [0, 2]], V[1, [0, 2]], V[2, [0, 2]], V[3, [0, 2]]

    def bary(p, a, b, cc):
        T = np.array([[b[0] - a[0], cc[0] - a[0]], [b[1] - a[1], cc[1] - a[1]]], np.float64)
        uv = np.linalg.solve(T, (p - a).T).T
        return np.stack([1 - uv.sum(1), uv[:, 0], uv[:, 1]], axis=1)
    w1 = bary(P.astype(np.float64), A, B, C)
    w2 = bary(P.astype(np.float64), A, C, D)
    in1 = (w1 >= -1e-9).all(1)
    lin = np.where(in1[:, None], w1 @ col[[0, 1, 2]].astype(np.float64), w2 @ col[[0, 2, 3]].astype(np.float64)) / 255 * 2
    expect = np.floor(255 * np.clip(lin, 0, 1) ** (1 / 2.2) + 0.5)
    diff = np.abs(rgb[0][want].astype(np.float64) - expect)
    # away from the shared diagonal the interpolated colour is the closed form to LUT rounding
    interior = (np.abs(w1).min(1) > 1e-3) & (np.abs(w2).min(1) > 1e-3)
    assert diff[interior].max() <= 1.0
    # a wall whose far vertices project 60 000 px off screen (outside the fixed-point guard band) is kept as well
    W = np.array([[-400, -0.5, 1.0], [0.2, -0.5, 2.0], [0.2, 0.2, 2.0], [-400, 0.2, 1.0]], np.float32)
    _, dw = R.render(W, F, col, np.eye(4)[None], f, f, c, c, res, msaa=1)
    assert (dw[0, 40:60, :20] > 0).all() and not (dw[0, :10, :] > 0).any()


def test_fine_stage_neighbourhood_matches_reference(golden):
    """online_pose_estimator.py:26-34,55-56: geodesic distance of the 20 000 fine poses to the previous pose and the
    `dists < neighborhood` selection -- index sets minted by the reference's own static method (scipy rotvec norm); the
    product computes the same angle in closed form (atan2 of the skew part and the trace)."""
    from freepose_b200.pipeline.estimators.online_pose_estimator import DinoOnlinePoseEstimator
    from freepose_b200.pipeline.utils import generate_poses
    g = golden["online_fine"]
    fine = np.array(generate_poses(20000))
    assert hashlib.sha256(np.ascontiguousarray(fine).tobytes()).hexdigest() == str(g["fine_sha"])
    class Sel:                                   # the estimator's selection method without constructing the CUDA engine
        fine_mesh_poses, _fine_rot = fine, None
        geodesic_distance = staticmethod(DinoOnlinePoseEstimator.geodesic_distance)
        neighbourhood = DinoOnlinePoseEstimator.neighbourhood
    sel = Sel()
    for i, T in enumerate(g["prev_poses"]):
        d = DinoOnlinePoseEstimator.geodesic_distance(fine, T)
        for nb in (15, 5):
            assert np.array_equal(np.where(d < nb)[0], g[f"close_{i}_{nb}"]), (i, nb)
            assert np.array_equal(sel.neighbourhood(T, nb), g[f"close_{i}_{nb}"]), (i, nb)     # pre-filtered path
        np.testing.assert_allclose(d[g[f"close_{i}_15"]], g[f"dist_{i}"], rtol=0, atol=1e-6)
    assert 1234 in g["close_6_5"] and len(g["close_1_5"]) == 0        # a fine pose finds itself; 5 degrees can be empty


def test_reciprocal_normalisation_is_exact():
    """score_rows_kernel evaluates the reference's bf16(t / norm) (F.normalize on a bf16 tensor, pose_estimator.py:85-86)
    as bf16(fl32(t * fl32(1 / norm))).  t and norm are bf16 values, so the quotient of their 8-bit significands can never
    come closer than 2^-17 (relative) to a bf16 rounding boundary while the two fp32 roundings err by < 2^-22: both
    forms round to the same bf16.  All significand pairs, a spread of exponents, plus subnormal-bf16 numerators."""
    def bf16_round(x):
        u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
        u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
        return u.astype(np.uint32).view(np.float32)

    sig = np.arange(128, 256, dtype=np.float32)
    T, N = np.meshgrid(sig, sig, indexing="ij")
    for et, en in [(0, 0), (-7, 0), (5, -3), (-20, 6), (12, 12), (-30, -10)]:
        t = np.ldexp(T, et - 7).astype(np.float32)
        n = np.ldexp(N, en - 7).astype(np.float32)
        want = bf16_round((t / n).astype(np.float32))
        r = (np.float32(1.0) / n).astype(np.float32)
        got = bf16_round((t * r).astype(np.float32))
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
        # an approximate reciprocal 1 ulp off still rounds the same way (the margin is 2^5 ulps)
        for off in (-1, 1):
            r1 = (r.view(np.uint32) + np.uint32(off)).view(np.float32) if off > 0 else (r.view(np.uint32) - np.uint32(1)).view(np.float32)
            assert np.array_equal(want.view(np.uint32), bf16_round((t * r1).astype(np.float32)).view(np.uint32))
    # numerators with fewer significant bits (any integer below 256)
    small = np.arange(1, 128, dtype=np.float32)
    T, N = np.meshgrid(small, sig, indexing="ij")
    want = bf16_round((T / N).astype(np.float32))
    got = bf16_round((T * (np.float32(1.0) / N).astype(np.float32)).astype(np.float32))
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


def test_raster_oracle_agrees_with_a_float64_ray_caster_on_a_closed_mesh():
    """A second, independent derivation of the restated raster spec (row R is unpinned: pyrender / OpenGL are not
    installable): Moller-Trumbore ray casting in float64 through the pixel centres (and, for 4x MSAA, through the four
    sample positions) of a vertex-coloured icosphere at a generic pose -- which face is nearest, its camera-space depth,
    and the barycentric (= perspective-correct) vertex colour x ambient 2 -> gamma 1/2.2 -> unorm8.  The C oracle must
    agree on coverage away from the silhouette, on the depth to 5e-4 relative (median 2e-5: the spec's 1/256 px vertex
    snapping) and on the colour to one unorm8 step
    wherever the hit is not within a hair of a triangle edge."""
    from freepose_b200.pipeline.utils import generate_poses
    from freepose_b200.synthetic import synthetic_mesh
    from oracle import raster as R
    mesh = synthetic_mesh(3, subdivisions=2)
    V = np.asarray(mesh.vertices, np.float64)
    F = np.asarray(mesh.faces, np.int64)
    col = np.asarray(mesh.vertex_colors, np.float64)[:, :3]
    pose = np.array(generate_poses(30))[17].astype(np.float32)
    res, f, c = 96, 137.0, 48.0
    Vc = V @ pose[:3, :3].astype(np.float64).T + pose[:3, 3].astype(np.float64)       # camera frame (OpenCV)
    A, B, C3 = Vc[F[:, 0]], Vc[F[:, 1]], Vc[F[:, 2]]

    def cast(offx, offy):
        """rays through (px + offx, py + offy): nearest face, depth Z, barycentrics; -1 where nothing is hit"""
        py, px = np.mgrid[0:res, 0:res]
        d = np.stack([(px + offx - c) / f, (py + offy - c) / f, np.ones((res, res))], -1).reshape(-1, 1, 3)   # (P,1,3)
        e1, e2 = (B - A)[None], (C3 - A)[None]
        h = np.cross(d, e2)
        det = (e1 * h).sum(-1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            s = -A[None]                                                    # origin - A
            u = (s * h).sum(-1) * inv
            q = np.cross(s, e1)
            v = (d * q).sum(-1) * inv
            t = (e2 * q).sum(-1) * inv                                      # d has Z = 1: t IS the camera-space depth
        ok = (np.abs(det) > 1e-14) & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > 0.05) & (t < 100)
        t = np.where(ok, t, np.inf)
        face = t.argmin(1)
        rows = np.arange(len(face))
        hit = np.isfinite(t[rows, face])
        w = np.stack([1 - u[rows, face] - v[rows, face], u[rows, face], v[rows, face]], -1)
        return (np.where(hit, face, -1).reshape(res, res), np.where(hit, t[rows, face], 0.0).reshape(res, res),
                w.reshape(res, res, 3))

    def shade(face, w):
        lin = (w[..., None] * col[F[face]]).sum(-2) / 255 * 2                # (…,3 verts,1) x (…,3 verts,3 ch)
        return np.floor(255 * np.clip(lin, 0, 1) ** (1 / 2.2) + 0.5)

    face, z, w = cast(0.5, 0.5)
    inside = face >= 0
    safe = inside & (w.min(-1) > 0.03)                                       # not within a hair of an edge
    # ---- msaa 1: everything at the pixel centre
    rgb, depth = R.render(V.astype(np.float32), F.astype(np.int32), col.astype(np.uint8), pose[None], f, f, c, c, res, msaa=1)
    got = depth[0] > 0
    assert safe.sum() > 1500 and got[safe].all()
    assert not got[~inside & ~_dilate(inside)].any()                         # nothing away from the silhouette
    # (the spec snaps vertices to 1/256 px like GL's sub-pixel grid: depth moves by up to the surface slope x 1/512 px)
    assert np.max(np.abs(depth[0][safe] - z[safe]) / z[safe]) < 5e-4        # (steep faces next to the silhouette)
    assert np.median(np.abs(depth[0][safe] - z[safe]) / z[safe]) < 2e-5
    assert np.abs(rgb[0][safe].astype(np.float64) - shade(face[safe], w[safe])).max() <= 1
    # ---- msaa 4: depth is sample 0's, the colour is shaded once per (face, pixel) at the centre and box-filtered
    offs = [(0.375, 0.125), (0.875, 0.375), (0.125, 0.625), (0.625, 0.875)]
    samples = [cast(ox, oy) for ox, oy in offs]
    same = safe.copy()
    for fs, _, ws in samples:
        same &= (fs == face) & (ws.min(-1) > 0.03)                            # all four samples on the centre's face
    rgb4, depth4 = R.render(V.astype(np.float32), F.astype(np.int32), col.astype(np.uint8), pose[None], f, f, c, c, res, msaa=4)
    assert same.sum() > 700
    z0 = samples[0][1]
    assert np.max(np.abs(depth4[0][same] - z0[same]) / z0[same]) < 5e-4
    assert np.abs(rgb4[0][same].astype(np.float64) - shade(face[same], w[same])).max() <= 1


def _dilate(mask):
    m = mask.copy()
    m[1:] |= mask[:-1]; m[:-1] |= mask[1:]; m[:, 1:] |= mask[:, :-1]; m[:, :-1] |= mask[:, 1:]
    return m
