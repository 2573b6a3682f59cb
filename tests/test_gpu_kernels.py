"""GPU (-m gpu): per-kernel parity through the C ABI against the CPU oracle on identical inputs.

Tolerances (north star: "bit-exact for index work, ViT tokens / RGB within 1e-3 rel"):
  * integer / index / byte outputs (raster RGB + coverage, bbox, crop gather, top-k indices): bit-exact;
  * score values: bit-exact against the engine-order oracle (the kernel fixes its fp32 summation order);
  * one ViT stage on identical bf16 inputs: REL_STAGE = 1e-3 relative L2 (observed ~1e-4: only isolated
    1-ulp bf16 rounding flips from fp32 accumulation order), and < 1 % of elements differ at all.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL_STAGE = 1e-3
bf = torch.bfloat16
dev = "cuda"


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def check_stage(out, ref, what, ulp_exact=True):
    """ulp_exact: outputs that are one rounding of an fp32 dot product (GEMM+bias, LayerNorm) may only differ by
    single bf16-ulp flips.  Stages with an approximated transcendental or a bf16-rounded intermediate (GELU tail,
    softmax P) are held to the tensor-level bounds only."""
    out, ref = out.float().cpu(), ref.float().cpu()
    assert out.shape == ref.shape, what
    assert torch.isfinite(out).all(), what
    r = rel_l2(out, ref)
    frac = (out != ref).float().mean().item()
    assert r < REL_STAGE, f"{what}: rel L2 {r:.3e}"
    assert frac < 0.01, f"{what}: {frac:.3%} elements differ"
    d = (out - ref).abs()
    assert d.max() <= 2 ** -7 * ref.abs().max(), f"{what}: max abs diff {d.max():.3e}"
    if ulp_exact:
        assert torch.all(d <= ref.abs() * 2 ** -7 + 2e-5 * (1 + ref.abs().mean())), what


@pytest.fixture(scope="module")
def ops(lib):
    from freepose_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M", [1, 127, 128, 129, 1000, 4177])
def test_gemm_bias_ragged_rows(ops, M):
    from freepose_b200._lib import FP_EPI_BIAS
    from oracle.vit import rb
    torch.manual_seed(M)
    a = torch.randn(M, 1024).to(bf)
    w = (torch.randn(3072, 1024) / 32).to(bf)
    b = (0.1 * torch.randn(3072)).to(bf)
    out = torch.full((M + 3, 3072), 7.0, dtype=bf, device=dev)
    ops.gemm(a.to(dev), w.to(dev), b.to(dev), FP_EPI_BIAS, out=out[:M])
    ref = rb(a.float() @ w.float().t() + b.float())
    check_stage(out[:M], ref, f"gemm bias M={M}")
    assert torch.all(out[M:] == 7.0), "rows past M were written"


def test_gemm_epilogues(ops):
    from freepose_b200._lib import FP_EPI_BIAS_GELU, FP_EPI_BIAS_LS_RES, FP_EPI_PATCH_EMBED
    from oracle.vit import rb
    torch.manual_seed(0)
    M = 777
    a = torch.randn(M, 1024).to(bf)
    w1 = (torch.randn(4096, 1024) / 32).to(bf)
    b1 = (0.1 * torch.randn(4096)).to(bf)
    out = ops.gemm(a.to(dev), w1.to(dev), b1.to(dev), FP_EPI_BIAS_GELU)
    ref = rb(torch.nn.functional.gelu(rb(a.float() @ w1.float().t() + b1.float())))
    check_stage(out, ref, "fc1 + gelu", ulp_exact=False)

    h = ref.to(bf)
    w2 = (torch.randn(1024, 4096) / 64).to(bf)
    b2 = (0.1 * torch.randn(1024)).to(bf)
    g = torch.rand(1024).to(bf)
    x = torch.randn(M, 1024).to(bf)
    ref2 = rb(x.float() + rb(rb(h.float() @ w2.float().t() + b2.float()) * g.float()))
    xd = x.clone().to(dev)
    out2 = ops.gemm(h.to(dev), w2.to(dev), b2.to(dev), FP_EPI_BIAS_LS_RES, gamma=g.to(dev), residual=xd)
    assert out2.data_ptr() == xd.data_ptr()  # in place on the residual stream
    check_stage(out2, ref2, "fc2 + layerscale + residual", ulp_exact=False)  # x + z may cancel

    B, P, T = 3, 16, 21
    pa = torch.zeros(B * P, 640, dtype=bf)
    pa[:, :588] = torch.randn(B * P, 588).to(bf)
    pw = torch.zeros(1024, 640, dtype=bf)
    pw[:, :588] = (torch.randn(1024, 588) / 24).to(bf)
    pb = (0.1 * torch.randn(1024)).to(bf)
    pos = torch.randn(1 + P, 1024).to(bf)
    tok = torch.zeros(B * T, 1024, dtype=bf, device=dev)
    ops.gemm(pa.to(dev), pw.to(dev), pb.to(dev), FP_EPI_PATCH_EMBED, out=tok, pos=pos.to(dev), patches_per_img=P,
             tokens_per_img=T, token_offset=5)
    ref3 = rb(rb(pa.float() @ pw.float().t() + pb.float()).view(B, P, 1024) + pos[1:].float())
    check_stage(tok.view(B, T, 1024)[:, 5:], ref3, "patch embed", ulp_exact=False)  # y + pos may cancel
    assert torch.all(tok.view(B, T, 1024)[:, :5] == 0), "special-token rows must not be touched by the GEMM"


def test_gemm_rejects_bad_shapes(ops):
    from freepose_b200._lib import FP_EPI_BIAS
    a = torch.zeros(8, 1000, dtype=bf, device=dev)
    w = torch.zeros(256, 1000, dtype=bf, device=dev)
    with pytest.raises(RuntimeError, match="K="):
        ops.gemm(a, w, torch.zeros(256, dtype=bf, device=dev), FP_EPI_BIAS)


# ------------------------------------------------------------------------------------------- LN / attention
def test_layernorm(ops):
    from oracle.vit import contract_layernorm
    torch.manual_seed(0)
    x = (torch.randn(1003, 1024) * 2 + 0.3).to(bf)
    w = (1 + 0.1 * torch.randn(1024)).to(bf)
    b = (0.1 * torch.randn(1024)).to(bf)
    out = ops.layernorm(x.to(dev), w.to(dev), b.to(dev))
    check_stage(out, contract_layernorm(x.float(), w.float(), b.float(), 1e-6), "layernorm")


def test_layernorm_768(ops):
    from oracle.vit import contract_layernorm
    torch.manual_seed(1)
    x = (torch.randn(517, 768) * 1.5 - 0.2).to(bf)
    w = (1 + 0.1 * torch.randn(768)).to(bf)
    b = (0.1 * torch.randn(768)).to(bf)
    out = ops.layernorm(x.to(dev), w.to(dev), b.to(dev))
    check_stage(out, contract_layernorm(x.float(), w.float(), b.float(), 1e-6), "layernorm 768")


@pytest.mark.parametrize("B,T", [(2, 261), (1, 17), (3, 272), (2, 256), (5, 128), (1, 129),
                                 # several (image, head) pairs per CTA: the cross-tile / cross-pair pipeline
                                 (40, 261), (30, 140), (24, 200), (40, 40), (21, 264)])
def test_attention(ops, B, T):
    from oracle.vit import contract_attention
    torch.manual_seed(T)
    qkv = torch.randn(B * T, 3072).to(bf)
    out = ops.attention(qkv.to(dev), B, T)
    q, k, v = qkv.float().view(B, T, 3, 16, 64).permute(2, 0, 3, 1, 4)
    ref = contract_attention(q, k, v, 0.125).transpose(1, 2).reshape(B * T, 1024)
    check_stage(out, ref, f"attention B={B} T={T}", ulp_exact=False)


@pytest.mark.parametrize("B,T", [(2, 905), (1, 273), (3, 512), (1, 529), (2, 1029), (1, 400),
                                 (8, 400), (5, 700), (6, 1025)])    # several work items per CTA; a 16-key last block
def test_attention_tiled_keys(ops, B, T):
    """Crops above 224^2 (905 tokens = the reference's 420^2 crops): key blocks of 256 with an online softmax."""
    from oracle.vit import contract_attention
    torch.manual_seed(T)
    qkv = torch.randn(B * T, 3072).to(bf)
    out = ops.attention(qkv.to(dev), B, T)
    q, k, v = qkv.float().view(B, T, 3, 16, 64).permute(2, 0, 3, 1, 4)
    ref = contract_attention(q, k, v, 0.125).transpose(1, 2).reshape(B * T, 1024)
    check_stage(out, ref, f"attention (tiled keys) B={B} T={T}", ulp_exact=False)


def test_attention_tiled_keys_growing_max(ops):
    """Keys ordered so that the running max rises in every block: the O rescale path carries all the weight."""
    from oracle.vit import contract_attention
    torch.manual_seed(11)
    B, T = 1, 905
    qkv = torch.randn(B * T, 3, 16, 64)
    qkv[:, 1] *= torch.linspace(0.5, 6.0, T).view(T, 1, 1)              # later keys give larger logits
    qkv = qkv.reshape(B * T, 3072).to(bf)
    out = ops.attention(qkv.to(dev), B, T)
    q, k, v = qkv.float().view(B, T, 3, 16, 64).permute(2, 0, 3, 1, 4)
    ref = contract_attention(q, k, v, 0.125).transpose(1, 2).reshape(B * T, 1024)
    check_stage(out, ref, "attention (tiled keys) growing max", ulp_exact=False)


def test_attention_sharp_softmax(ops):
    """Large logits (peaked softmax) -- the max subtraction must keep it finite and exact."""
    from oracle.vit import contract_attention
    torch.manual_seed(3)
    B, T = 2, 261
    qkv = (torch.randn(B * T, 3072) * 4).to(bf)
    out = ops.attention(qkv.to(dev), B, T)
    q, k, v = qkv.float().view(B, T, 3, 16, 64).permute(2, 0, 3, 1, 4)
    ref = contract_attention(q, k, v, 0.125).transpose(1, 2).reshape(B * T, 1024)
    check_stage(out, ref, "attention sharp", ulp_exact=False)


# ------------------------------------------------------------------------------------------- preprocessing
def test_normalize_and_im2col_bit_exact(ops):
    from oracle.pipeline import reference_normalize
    torch.manual_seed(0)
    img = torch.rand(3, 3, 56, 56)
    img[0, :, :4, :4] = torch.tensor([0.0, 1.0, 0.5, 1 / 255]).view(1, 1, 4)
    ref = reference_normalize(img.to(bf))                              # torchvision Normalize on the bf16 tensor
    got = ops.normalize_image(img.to(dev)).cpu()
    assert torch.equal(got, ref)
    patches = ops.im2col(img.to(dev)).cpu()                            # fused normalise + gather
    want = torch.nn.functional.unfold(ref.float(), kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert torch.equal(patches[:, :588].float(), want) and torch.all(patches[:, 588:] == 0)
    assert torch.equal(ops.im2col(ref.to(dev)).cpu(), patches)        # bf16 pre-normalised input path


# ------------------------------------------------------------------------------------------- score / top-k / FFA
def _score_case(B, P, seed, near=True):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(1, P, 1024, generator=g)
    t = 0.6 * q + 0.8 * torch.randn(B, P, 1024, generator=g) if near else torch.randn(B, P, 1024, generator=g)
    return t.to(bf), q.to(bf)


@pytest.mark.parametrize("B,P", [(1, 1), (3, 7), (37, 256), (64, 900)])
def test_score_bit_exact(ops, B, P):
    from oracle import score as S
    ft, fq = _score_case(B, P, B * 1000 + P)
    k = min(3, B)
    scores, idx, vals, patch = ops.score_topk(ft.to(dev), fq.to(dev), k=k, return_patch_scores=True)
    want, want_patch = S.engine_order_scores(ft, fq, return_patch=True)
    assert np.array_equal(patch.cpu().numpy(), want_patch)
    assert np.array_equal(scores.cpu().numpy(), want)
    widx, wvals = S.stable_topk(want, k)
    assert idx.cpu().numpy().astype(np.int64).tolist() == widx.tolist()
    assert np.array_equal(vals.cpu().numpy(), wvals)
    # and against the reference's own lines on CPU bf16: identical up to isolated 1-ulp flips, same winner
    ref = S.reference_scores(ft, fq).float().numpy()
    assert np.all(np.abs(ref - want) <= np.abs(ref) * 2 ** -7)
    assert (ref != want).mean() <= 0.05


@pytest.mark.parametrize("D", [256, 512, 768])
def test_score_bit_exact_other_token_widths(ops, D):
    """score_rows_kernel is instantiated per D / 256 (ViT-S/B/L widths); 6 hypotheses = one full work item + a tail of two."""
    from oracle import score as S
    g = torch.Generator().manual_seed(D)
    fq = torch.randn(1, 37, D, generator=g).to(bf)
    ft = (0.6 * fq.float() + 0.8 * torch.randn(6, 37, D, generator=g)).to(bf)
    scores, idx, vals, patch = ops.score_topk(ft.to(dev), fq.to(dev), k=2, return_patch_scores=True)
    want, want_patch = S.engine_order_scores(ft, fq, return_patch=True)
    assert np.array_equal(patch.cpu().numpy(), want_patch) and np.array_equal(scores.cpu().numpy(), want)
    widx, _ = S.stable_topk(want, 2)
    assert idx.cpu().numpy().astype(np.int64).tolist() == widx.tolist()


def test_score_publish_and_topk_after_exchange_single_rank(ops, tmp_path):
    """The peer-memory exchange (fp_score_publish / fp_topk_after_exchange) with a world of one:
    the rank's own buffer is its only peer, so the publishing stores, the completion flag, the epoch parity and the
    device-side wait all run on a single GPU and must reproduce fp_score_topk bit for bit, exchange after exchange."""
    import torch.distributed as dist
    from freepose_b200.distributed import PeerScoreGather
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("gloo", init_method=f"file://{tmp_path}/rdzv", rank=0, world_size=1)
    try:
        B, P = 13, 256
        sg = PeerScoreGather(B, 1, torch.device(dev), rank=0)
        for rep in range(5):                                   # five epochs: both parities of the buffer, flags re-armed
            ft, fq = _score_case(B, P, 77 + rep)
            want_scores, want_idx, want_val, _ = ops.score_topk(ft.to(dev), fq.to(dev), k=3)
            sg.score_into(0, ft.to(dev), fq.to(dev))
            scores, idx, val = sg.gather_topk(0, 3)
            assert torch.equal(scores, want_scores) and torch.equal(idx, want_idx) and torch.equal(val, want_val), rep
        sg.close()
    finally:
        if created:
            dist.destroy_process_group()


def test_score_and_vit_edge_cases(ops, lib):
    """Empty and degenerate inputs at the C boundary: zero hypotheses are a no-op, k larger than the number of hypotheses
    and a token width the row kernels are not instantiated for are rejected with a message (no launch), a single
    hypothesis with a single patch works (the smallest work item)."""
    from freepose_b200 import _lib
    f = torch.zeros(0, 16, 256, dtype=bf, device=dev)
    q = torch.randn(16, 256, device=dev).to(bf)
    scores, idx, vals, _ = ops.score_topk(f, q, k=0)
    assert scores.numel() == 0
    with pytest.raises(RuntimeError, match="k=3 out of range"):
        ops.score_topk(torch.randn(2, 16, 256, device=dev).to(bf), q, k=3)
    with pytest.raises(RuntimeError, match="multiple of 256"):
        ops.score_topk(torch.randn(2, 16, 320, device=dev).to(bf), torch.randn(16, 320, device=dev).to(bf), k=1)
    one = torch.randn(1, 1, 1024, device=dev).to(bf)
    s1, i1, v1, _ = ops.score_topk(one, one[0], k=1)
    assert int(i1[0]) == 0 and abs(float(s1[0]) - 1.0) < 2 ** -7          # cosine of a row with itself, bf16-rounded
    assert _lib.load().fp_last_error() is not None


def test_score_golden_ties_weights_and_raw_query(ops, golden):
    from oracle import score as S
    g = golden["score"]
    ft = torch.from_numpy(g["feats_t"]).view(bf)
    fq = torch.from_numpy(g["feat_q"]).view(bf)
    scores, idx, vals, _ = ops.score_topk(ft.to(dev), fq.to(dev), k=3)
    assert np.array_equal(scores.cpu().numpy(), g["scores"])          # the reference's output, bit for bit
    assert np.array_equal(vals.cpu().numpy(), g["top_scores"])
    assert scores[3] == scores[5]
    order = idx.cpu().tolist()
    assert order == S.stable_topk(g["scores"], 3)[0].tolist()          # ties -> lowest index
    w = torch.from_numpy(g["masks"])
    sw, iw, _, _ = ops.score_topk(ft.to(dev), fq.to(dev), k=1, weights=w.to(dev))
    assert np.array_equal(sw.cpu().numpy(), S.engine_order_scores(ft, fq, weights=w))
    np.testing.assert_allclose(sw.cpu().numpy(), g["weighted"], rtol=2e-6)
    assert int(iw) == int(np.argmax(g["weighted"]))
    # coarse->fine quirk: the coarse query feature is used WITHOUT normalisation (online_pose_estimator.py:41,50)
    sr, _, _, _ = ops.score_topk(ft.to(dev), fq.to(dev), k=1, normalise_query=False)
    assert np.array_equal(sr.cpu().numpy(), S.engine_order_scores(ft, fq, normalise_query=False))


def test_score_fine_stage_masked_and_raw_query_vs_reference_lines(ops, golden):
    """Fine stage (online_pose_estimator.py:67-79): mask-weighted score with the (30, 30) bilinear resize of
    OR(template mask, proposal mask), for a normalised query (frames > 0) and for the RAW coarse query feature (first
    frame).  Expected values: the reference's lines executed verbatim on the CPU (tests/golden/online_fine.npz); the
    inputs are regenerated from the same seeds."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
    from make_golden import online_fine_inputs
    import torch.nn.functional as F
    g = golden["online_fine"]
    feats, query, tmasks, pmask = online_fine_inputs()
    # the estimator's own weight computation (pipeline/estimators/online_pose_estimator.py forward_fine), on the device
    mk = torch.logical_or(tmasks.to(dev), pmask.to(dev)[None]).float()
    weights = F.interpolate(mk[None], size=(30, 30), mode="bilinear")[0].reshape(len(feats), 900).contiguous()
    np.testing.assert_allclose(weights.cpu().numpy(), g["weights"], rtol=0, atol=1e-6)
    for tag, normalise in (("norm", True), ("raw", False)):
        sm, im, _, _ = ops.score_topk(feats.to(dev), query.to(dev), k=1, weights=weights, normalise_query=normalise)
        np.testing.assert_allclose(sm.cpu().numpy(), g[f"masked_{tag}"], rtol=3e-6)
        assert int(im) == int(g[f"argmax_masked_{tag}"])
        su, iu, _, _ = ops.score_topk(feats.to(dev), query.to(dev), k=1, normalise_query=normalise)
        ref = g[f"plain_{tag}"]
        ulp = 2.0 ** (np.floor(np.log2(np.abs(ref))) - 7)
        assert np.all(np.abs(su.cpu().numpy() - ref) <= ulp)                  # bf16-valued: at most one ulp (summation order)
        assert ref[int(iu)] >= ref.max() - ulp[int(iu)]


def test_score_full_size_properties(ops):
    """BASELINE size (520 x 256 x 1024): permutation equivariance, power-of-two scale invariance, determinism."""
    ft, fq = _score_case(520, 256, 5)
    ft, fq = ft.to(dev), fq.to(dev)
    s0, i0, v0, _ = ops.score_topk(ft, fq, k=3)
    s1, i1, _, _ = ops.score_topk(ft, fq, k=3)
    assert torch.equal(s0, s1) and torch.equal(i0, i1)
    perm = torch.randperm(520, generator=torch.Generator().manual_seed(0)).to(dev)
    sp, _, vp, _ = ops.score_topk(ft[perm].contiguous(), fq, k=3)
    assert torch.equal(sp, s0[perm]) and torch.equal(vp, v0)
    s4, _, _, _ = ops.score_topk((ft.float() * 4).to(bf), (fq.float() * 0.5).to(bf), k=3)
    assert torch.equal(s4, s0)  # cosine: exact under power-of-two rescaling
    # a template equal to the query scores bf16(~1) and wins
    ft2 = ft.clone()
    ft2[123] = fq[0]
    s5, i5, _, _ = ops.score_topk(ft2, fq, k=1)
    assert int(i5) == 123 and abs(float(s5[123]) - 1.0) < 0.01
    idx, vals = ops.topk(s0, 5)
    assert idx[:3].tolist() == i0.tolist()


def test_ffa_pool_bit_exact(ops):
    from oracle import score as S
    torch.manual_seed(0)
    feats = torch.randn(6, 256, 1024).to(bf)
    rng = np.random.default_rng(0)
    masks = rng.random((6, 224, 224)) > 0.9995
    masks[4] = False
    masks[5] = True
    out, valid = ops.ffa_pool(feats.to(dev), torch.from_numpy(masks).to(dev))
    want, counts = S.ffa_engine_order(feats, masks)
    assert valid.cpu().tolist() == counts.tolist() and counts[5] == 256 and counts[4] == 0
    got = out.cpu().numpy()
    assert np.isnan(got[4]).all()
    assert np.array_equal(got[[0, 1, 2, 3, 5]], want[[0, 1, 2, 3, 5]])


# ------------------------------------------------------------------------------------------- raster
def _mesh(sub):
    from freepose_b200.synthetic import synthetic_mesh
    return synthetic_mesh(0, subdivisions=sub)


def _poses(n, seed=0, z=1.1):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        p = np.eye(4)
        p[:3, :3] = q
        p[:3, 3] = [rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), z]
        out.append(p)
    return np.array(out)


@pytest.mark.parametrize("sub,res,msaa,cull", [(3, 224, 4, False), (3, 224, 1, False), (0, 224, 4, False),
                                              (4, 420, 4, False), (2, 224, 4, True), (5, 224, 4, False)])
def test_raster_bit_exact(ops, sub, res, msaa, cull):
    from freepose_b200.pipeline.utils import mesh_to_device
    from freepose_b200.synthetic import camera_for
    from oracle import raster as R
    m = _mesh(sub)
    poses = _poses(5, seed=sub + res)
    poses[4, :3, 3] = [0.3, -0.28, 0.9]  # partially outside the frustum: clipping against the viewport
    fx, fy, cx, cy = camera_for(res)
    want_rgb, want_depth = R.render(m.vertices, m.faces, m.vertex_colors, poses, fx, fy, cx, cy, res, msaa, cull)
    v, f, c = mesh_to_device(m, torch.device(dev))
    rgb, depth = ops.rasterize(v, f, c, torch.from_numpy(poses).float().to(dev), fx, fy, cx, cy, res, msaa, cull)
    assert np.array_equal(rgb.cpu().numpy(), want_rgb), "RGB differs"
    assert np.array_equal(depth.cpu().numpy(), want_depth), "depth differs"
    assert (want_depth > 0).sum() > 1000


@pytest.mark.parametrize("tex_size,res,msaa,vcol", [((512, 256), 224, 4, False), ((512, 256), 224, 1, False),
                                                   ((2048, 1024), 224, 4, False),      # strong minification: upper mips
                                                   ((37, 19), 224, 4, True),           # odd sizes, magnified, x COLOR_0
                                                   ((1, 1), 224, 4, False), ((256, 256), 420, 4, False)])
def test_raster_textured_bit_exact(ops, tex_size, res, msaa, vcol):
    """Textured meshes (trimesh TextureVisuals -> pyrender baseColorTexture): CUDA vs the C restatement, bit for bit."""
    from freepose_b200.synthetic import camera_for, synthetic_textured_mesh
    from oracle import raster as R
    m = synthetic_textured_mesh(1, 3, tex_size, with_vertex_colors=vcol)
    poses = _poses(4, seed=res + tex_size[0])
    poses[3, :3, 3] = [0.3, -0.28, 0.9]
    fx, fy, cx, cy = camera_for(res)
    want_rgb, want_depth = R.render_mesh(m, poses, fx, fy, cx, cy, res, msaa)
    rgb, depth = ops.rasterize_mesh(m, torch.from_numpy(poses).float().to(dev), fx, fy, cx, cy, res, msaa)
    assert np.array_equal(depth.cpu().numpy(), want_depth), "depth differs"
    diff = rgb.cpu().numpy().astype(int) - want_rgb.astype(int)
    assert np.array_equal(rgb.cpu().numpy(), want_rgb), f"RGB differs in {np.count_nonzero(diff)} values, max {np.abs(diff).max()}"
    assert (want_depth > 0).sum() > 1000 and want_rgb[want_depth > 0].std() > 5


@pytest.mark.parametrize("n,res,msaa", [(20000, 224, 4), (3000, 224, 1), (50000, 420, 4)])
def test_raster_points_bit_exact(ops, n, res, msaa):
    """trimesh.PointCloud inputs (reference renderer.py:46-51 -> pyrender.Mesh.from_points): 1-pixel point sprites."""
    from freepose_b200.synthetic import camera_for, synthetic_point_cloud
    from oracle import raster as R
    pc = synthetic_point_cloud(2, n)
    poses = _poses(3, seed=n)
    fx, fy, cx, cy = camera_for(res)
    want_rgb, want_depth = R.render_mesh(pc, poses, fx, fy, cx, cy, res, msaa)
    rgb, depth = ops.rasterize_mesh(pc, torch.from_numpy(poses).float().to(dev), fx, fy, cx, cy, res, msaa)
    assert np.array_equal(rgb.cpu().numpy(), want_rgb), "RGB differs"
    assert np.array_equal(depth.cpu().numpy(), want_depth), "depth differs"
    assert (want_depth > 0).sum() > 500


def test_renderer_accepts_trimesh_like_inputs(ops):
    """MeshRenderer.render_from_poses takes what the reference passes: objects with .vertices/.faces/.visual (vertex or
    texture kind) or point clouds with .colors; empty point-cloud colours render white (renderer.py:47-50)."""
    from types import SimpleNamespace as NS
    from freepose_b200.pipeline.retrieval.renderer import MeshRenderer
    from freepose_b200.synthetic import synthetic_point_cloud, synthetic_textured_mesh
    r = MeshRenderer(2, resolution=224)
    tm = synthetic_textured_mesh(0, 3)
    tri = NS(vertices=tm.vertices, faces=tm.faces,
             visual=NS(kind="texture", uv=tm.uv, material=NS(image=tm.texture)))
    a = r.render_from_poses(tri, r.mesh_poses)
    b = r.render_from_poses(tm, r.mesh_poses)
    assert all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(a, b))
    pc = synthetic_point_cloud(0, 5000)
    white = r.render_from_poses(NS(vertices=pc.vertices, colors=np.zeros((0, 4), np.uint8)), r.mesh_poses)
    rgb, depth = white[0][0], white[0][1]
    assert (depth > 0).sum() > 100 and set(np.unique(rgb[depth > 0])) <= {64, 128, 191, 255}   # white x MSAA coverage


def test_raster_long_offscreen_triangles_with_small_clipped_boxes(ops):
    """Triangles whose vertices are thousands of pixels apart but whose viewport-clipped box is a few pixels: they must not
    take the 32-bit small-triangle path (their edge functions do not fit) -- bit-exact against the oracle."""
    from freepose_b200.pipeline.utils import Mesh
    from oracle import raster as R
    px = lambda u, v, z=1.0: [(u - 112.0) / 320.0 * z, (v - 112.0) / 320.0 * z, z - 1.0]
    verts = np.array([px(-4000, -4000), px(5, 2), px(-4000, 3),         # wedge into the top-left corner: clipped box 6 x 3 px
                      px(4224, -4000), px(218, 2), px(4224, 3),         # ... and into the top-right corner
                      px(100, 100), px(9000, 104), px(100, 108),        # long to the right, big on screen too
                      px(50, 50, 1.2), px(52, 50, 1.2), px(51, 52, 1.2)])   # an ordinary tiny one
    faces = np.arange(12).reshape(4, 3)
    colors = np.tile(np.array([[90, 40, 20], [20, 90, 40], [40, 20, 90]], np.uint8), (4, 1))
    m = Mesh(verts, faces, colors)
    pose = np.eye(4)[None].copy()
    pose[0, 2, 3] = 1.0
    for msaa in (4, 1):
        want_rgb, want_depth = R.render_mesh(m, pose, 320.0, 320.0, 112.0, 112.0, 224, msaa)
        rgb, depth = ops.rasterize_mesh(m, torch.from_numpy(pose).float().to(dev), 320.0, 320.0, 112.0, 112.0, 224, msaa)
        assert np.array_equal(rgb.cpu().numpy(), want_rgb) and np.array_equal(depth.cpu().numpy(), want_depth)
        assert (want_depth[0, 0:3, 0:5] > 0).any() and (want_depth[0, 0:3, 219:224] > 0).any()


def test_raster_behind_camera_and_empty(ops):
    from freepose_b200.pipeline.utils import mesh_to_device
    m = _mesh(1)
    v, f, c = mesh_to_device(m, torch.device(dev))
    poses = _poses(2, z=-1.0)
    rgb, depth = ops.rasterize(v, f, c, torch.from_numpy(poses).float().to(dev), 320, 320, 112, 112, 224)
    assert int(rgb.max()) == 0 and float(depth.max()) == 0.0
    with pytest.raises(RuntimeError, match="multiple of 4"):
        ops.rasterize(v, f, c, torch.from_numpy(poses).float().to(dev), 320, 320, 112, 112, 222)


def _clip_scene():
    """A ground quad running from BEHIND the camera (z = -1) to z = 6, a wall whose left vertices project 60 000 px off
    screen, and a small ordinary triangle: everything GL would clip, nothing it would drop."""
    from freepose_b200.pipeline.utils import Mesh
    verts = np.array([[-2, 0.3, -1.0], [2, 0.3, -1.0], [2, 0.3, 6.0], [-2, 0.3, 6.0],           # ground, crosses z = 0 and znear
                      [-400, -0.5, 1.0], [0.2, -0.5, 2.0], [0.2, 0.2, 2.0], [-400, 0.2, 1.0],   # wall far off screen to the left
                      [-0.1, -0.1, 1.5], [0.1, -0.1, 1.5], [0.0, 0.1, 1.5]], np.float32)
    faces = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [8, 9, 10]], np.int32)
    colors = np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0], [90, 40, 20], [20, 90, 40], [40, 20, 90],
                       [120, 120, 30], [10, 100, 100], [100, 10, 100], [100, 100, 10]], np.uint8)
    return Mesh(verts, faces, colors)


@pytest.mark.parametrize("msaa,cull,znear", [(4, False, 0.0), (1, False, 0.0), (4, True, 0.0), (4, False, 1e-4)])
def test_raster_near_plane_and_guard_band_clipping_bit_exact(ops, msaa, cull, znear):
    """ADVICE r1 / VERDICT r1 #6: triangles that straddle the near plane, reach behind the camera or leave the 16 384 px
    guard band are clipped (per sample, homogeneous rasterisation), not dropped -- CUDA == C restatement bit for bit; the
    restatement itself is checked against an analytic ray / plane intersection in tests/test_oracle_golden.py."""
    from oracle import raster as R
    m = _clip_scene()
    poses = np.stack([np.eye(4), np.eye(4)])
    poses[1, :3, 3] = [0.05, -0.1, 0.2]
    kw = dict(znear=znear, zfar=0.0)
    want_rgb, want_depth = R.render_mesh(m, poses, 150.0, 150.0, 64.0, 64.0, 128, msaa, cull, **kw)
    rgb, depth = ops.rasterize_mesh(m, torch.from_numpy(poses).float().to(dev), 150.0, 150.0, 64.0, 64.0, 128, msaa, cull, **kw)
    assert np.array_equal(depth.cpu().numpy(), want_depth), "depth differs"
    assert np.array_equal(rgb.cpu().numpy(), want_rgb), "RGB differs"
    if not cull:
        assert (want_depth[0, 100:, :] > 0).mean() > 0.9        # the ground fills the bottom of the image
        assert (want_depth[0, 40:60, :20] > 0).mean() > 0.9     # ... and the off-screen wall its left edge


def test_raster_camera_inside_a_textured_mesh_bit_exact(ops):
    """The camera INSIDE a closed textured mesh (every triangle around it, many behind it or through the near plane), with
    the refiner's per-view intrinsics and its znear = 1e-4 (reference tracking_refiner.py:30-43)."""
    from freepose_b200.synthetic import synthetic_textured_mesh
    from oracle import raster as R
    m = synthetic_textured_mesh(1, 2, (64, 32), with_vertex_colors=True)
    poses = np.stack([np.eye(4), np.eye(4)])
    poses[0, :3, 3] = [0.02, -0.03, 0.05]          # the object's centre 5 cm in front of the camera: the camera is inside it
    poses[1, :3, 3] = [0.0, 0.0, 0.26]             # just outside, surface through the near plane region
    view_k = np.array([[900.0, 900.0, 259.0, 259.0], [2500.0, 2400.0, 200.0, 300.0]], np.float32)
    kw = dict(ambient=5.0, znear=1e-4, zfar=9999.0)
    want_rgb, want_depth = R.render_mesh(m, poses, 1.0, 1.0, 0.0, 0.0, 520, 4, False, view_k=view_k, **kw)
    rgb, depth = ops.rasterize_mesh(m, torch.from_numpy(poses).float().to(dev), 1.0, 1.0, 0.0, 0.0, 520, 4, False,
                                    view_k=torch.from_numpy(view_k).to(dev), **kw)
    assert np.array_equal(depth.cpu().numpy(), want_depth), "depth differs"
    assert np.array_equal(rgb.cpu().numpy(), want_rgb), "RGB differs"
    assert (want_depth[0] > 0).mean() > 0.99        # from the inside, the mesh covers the whole view


# ------------------------------------------------------------------------------------------- geometry
def test_mask_bbox_and_fallback(ops):
    from oracle import crop as C
    rng = np.random.default_rng(0)
    d = np.zeros((4, 420, 420), np.float32)
    d[0, 100:300, 150:330] = 1.0
    d[1, 7, 9] = 0.5                      # tiny mask -> reference forces the centre square
    d[2, 0, 0] = d[2, 419, 419] = 1.0
    d[3] = (rng.random((420, 420)) > 0.999) * 1.0
    bbox, count, mask = ops.mask_bbox(torch.from_numpy(d).to(dev), fallback=(105, 315), return_mask=True)
    for i in range(4):
        m = d[i] > 0
        if m.sum() < 100:
            m = m.copy()
            m[105:315, 105:315] = True
        assert bbox[i].cpu().tolist() == C.mask_to_bbox(m).tolist()
        assert np.array_equal(mask[i].cpu().numpy().astype(bool), m)
        assert int(count[i]) == int((d[i] > 0).sum())


def test_crop_resize_pad_golden_and_patches(ops, golden):
    from oracle import crop as C
    from oracle.pipeline import reference_normalize
    g = golden["crop"]
    rng = np.random.default_rng(int(g["rng_seed"]))
    i = 0
    while f"c{i}_spec" in g.files:
        H, W, T = [int(v) for v in g[f"c{i}_spec"]]
        ext = float(g[f"c{i}_ext"])
        ext = int(ext) if ext == 0 else ext
        boxes = g[f"c{i}_boxes"]
        imgs = rng.random((len(boxes), 3, H, W)).astype(np.float32)
        want = C.crop_resize_pad(imgs, boxes, T, bbox_extend=ext, orig_size=(H, W))  # == reference (CPU test)
        ext_boxes = np.array([C.extend_box(b, ext, W, H) for b in boxes], dtype=np.int32)
        got, status = ops.crop_resize_pad(torch.from_numpy(imgs).to(dev), torch.from_numpy(ext_boxes).to(dev), T)
        assert int(status) == 0
        assert np.array_equal(got.cpu().numpy(), want), f"crop case {i}"
        i += 1
    # u8 render -> normalised bf16 patch matrix == reference chain (img/255 -> crop -> bf16 -> Normalize -> unfold)
    rgb = rng.integers(0, 256, (3, 224, 224, 3), dtype=np.uint8)
    boxes = np.array([[40, 50, 180, 200], [0, 0, 223, 223], [100, 20, 130, 210]], dtype=np.int32)
    crops = C.crop_resize_pad((rgb / 255).astype(np.float32).transpose(0, 3, 1, 2), boxes, 224)
    norm = reference_normalize(torch.from_numpy(crops).to(bf))
    want = torch.nn.functional.unfold(norm.float(), kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 588)
    patches, status = ops.crop_resize_pad(torch.from_numpy(rgb).to(dev), torch.from_numpy(boxes).to(dev), 224,
                                          to_patches=True)
    assert int(status) == 0
    assert torch.equal(patches.cpu()[:, :588].float(), want) and torch.all(patches[:, 588:] == 0)
    f32, _ = ops.crop_resize_pad(torch.from_numpy(rgb).to(dev), torch.from_numpy(boxes).to(dev), 224)
    assert np.array_equal(f32.cpu().numpy(), crops)
    # degenerate box: the reference raises; the kernel reports which box
    _, status = ops.crop_resize_pad(torch.from_numpy(rgb).to(dev),
                                    torch.tensor([[10, 10, 50, 50], [30, 30, 30, 90], [0, 0, 9, 9]]).to(dev), 224)
    assert int(status) == 2


def test_depth_extents_vs_reference_geometry(ops, golden):
    from freepose_b200.pipeline import utils as U
    g = golden["geometry"]
    for i in range(4):
        c = {k[len(f"c{i}_"):]: g[k] for k in g.files if k.startswith(f"c{i}_")}
        d = torch.from_numpy(c["depth"])[None].to(dev)
        ext = ops.depth_extents(d, c["Kt"]).cpu().numpy()[0]
        assert int(ext[7]) == int(c["n_points"])
        s = float(c["est_scale"])
        for recentre, want in ((True, c["tco_coarse"]), (False, c["tco_fine"])):
            dx, dy = U.rescaled_extents(ext, s, recentre)
            got = U.tco_from_extents(c["bbox"], dx, dy, c["Kq"], c["T0"])
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)  # the reference's TCO (fixture)
    # view selection + empty view
    d = torch.zeros(3, 224, 224, device=dev)
    d[1, 10:20, 30:50] = 1.0
    e = ops.depth_extents(d, np.array([[320.0, 0, 112], [0, 320.0, 112], [0, 0, 1]]),
                          view_idx=torch.tensor([1, 0], device=dev)).cpu().numpy()
    assert e[0, 7] == 200 and e[1, 7] == 0


# ------------------------------------------------------------------------------------------- mesh retrieval
def test_normalize_rows_and_scan_bit_exact(ops):
    from oracle import retrieval as R
    rng = np.random.default_rng(5)
    for M, D, Q in ((1, 1024, 1), (37, 1024, 3), (2500, 1024, 33), (300, 768, 2), (64, 256, 5)):
        db = rng.standard_normal((M, D)).astype(np.float32) * rng.uniform(0.1, 30, size=(M, 1)).astype(np.float32)
        if M > 2:
            db[1] = 0                                                   # zero row: clamped by eps, stays 0
        qs = rng.standard_normal((Q, D)).astype(np.float32)
        dbn, qn = R.engine_normalize(db), R.engine_normalize(qs)
        got_db = ops.normalize_rows(torch.from_numpy(db).to(dev))
        assert torch.equal(got_db.float().cpu(), torch.from_numpy(dbn))
        # bf16 input path == fp32 input path after the cast
        assert torch.equal(ops.normalize_rows(torch.from_numpy(db).to(bf).to(dev)), got_db)
        got_q = ops.normalize_rows(torch.from_numpy(qs).to(dev))
        scores = ops.retrieval_scan(got_db, got_q).cpu().numpy()
        assert np.array_equal(scores, R.engine_scan(dbn, qn))


@pytest.mark.parametrize("M,k", [(1, 1), (100, 100), (1000, 3), (5000, 100), (46037, 100), (46037, 1024), (70000, 517)])
def test_topk_rows_bit_exact_with_ties_and_specials(ops, M, k):
    from oracle.retrieval import engine_topk
    rng = np.random.default_rng(M + k)
    s = torch.from_numpy(rng.standard_normal((3, M)).astype(np.float32)).to(bf).float().numpy()  # bf16-valued: many ties
    s[1] = np.round(s[1] * 4) / 4                                                                # massive ties
    if M >= 1000:
        s[2, 5], s[2, 17], s[2, 400], s[2, 401], s[2, 402] = np.nan, np.inf, -np.inf, 0.0, -0.0
        s[2, 900] = np.nan
    idx, val = ops.topk_rows(torch.from_numpy(s).to(dev), k)
    ridx, rval = engine_topk(s, k)
    assert np.array_equal(idx.cpu().numpy(), ridx)
    assert np.array_equal(val.cpu().numpy(), rval, equal_nan=True)
    with pytest.raises(RuntimeError, match="out of range"):
        ops.topk_rows(torch.zeros(1, 4, device=dev), 5)


def test_retrieval_database_vs_oracle_and_reference_fixture(ops, golden):
    from freepose_b200.pipeline.retrieval.database import RetrievalDatabase, SoftVote
    from oracle import retrieval as R
    g = golden["retrieval"]
    case = R.synthetic_case(0)
    dbn = R.engine_normalize(case["db"])
    qn = R.engine_normalize(case["queries"])
    loads = []

    def loader(m):
        loads.append(m)
        return case["fine"][m]

    db = RetrievalDatabase(case["db"], case["ids"], fine_loader=loader, pool_views=100 * 8 * 40)
    feats = db.normalize(torch.from_numpy(case["queries"]))
    assert torch.equal(feats.float().cpu(), torch.from_numpy(qn))
    for topk in (0, 3, 10):
        meshes, scores, cand, cs = db.retrieve(feats, topk=topk, return_sparse=True)
        best, score, ecand, ecs = R.engine_retrieve(dbn, case["fine"], qn, topk)
        assert np.array_equal(cand.cpu().numpy(), ecand)                       # candidate indices: bit-exact
        assert np.array_equal(cs.cpu().numpy(), ecs)                           # coarse / fine scores: bit-exact
        assert meshes == [case["ids"][m] for m in best] and np.array_equal(np.float32(scores), score)
        assert meshes == [case["ids"][m] for m in g[f"best_{topk}"]]           # = the reference lines' retrieval
        assert np.array_equal(np.float64(scores), g[f"score_{topk}"])
    assert len(loads) == len(set(loads))                                       # every mesh file read at most once
    # preloaded store gives the same answer
    db2 = RetrievalDatabase(case["db"], case["ids"], fine_features=case["fine"])
    assert db2.retrieve(feats, topk=3) == db.retrieve(feats, topk=3)
    with pytest.raises(RuntimeError, match="out of range"):
        db2.retrieve(feats, topk=60)                                           # more than a mesh has views (torch raises)
    # video soft vote
    vote, per_frame = SoftVote(db, 3), []
    for fr in case["video"]:
        f = db.normalize(torch.from_numpy(fr))
        _, _, cand, cs = db.retrieve(f, topk=3, return_sparse=True)
        vote.add_frame(cand, cs)
        per_frame.append((cand.cpu().numpy(), cs.cpu().numpy()))
    vm, vs = vote.result()
    ebest, escore, emean = R.engine_softvote_dense(per_frame, db.M)
    assert vm == [case["ids"][m] for m in ebest] and np.array_equal(np.float32(vs), escore)
    assert vm == [case["ids"][m] for m in g["vote_best"]]
    np.testing.assert_allclose(np.float32(vs), g["vote_score"], rtol=1e-6)


def test_retrieval_full_size_properties(ops):
    """46 037 x 1024 (the reference's table size): planted rows must come out on top, scan is linear in the query count."""
    from freepose_b200.pipeline.retrieval.database import RetrievalDatabase
    g = torch.Generator(device=dev).manual_seed(0)
    M, D, Q = 46037, 1024, 12
    table = torch.randn(M, D, device=dev, generator=g)
    planted = torch.randint(0, M, (Q,), device=dev, generator=g)
    queries = table[planted] + 0.3 * torch.randn(Q, D, device=dev, generator=g)
    db = RetrievalDatabase(table, [str(i) for i in range(M)])
    feats = db.normalize(queries)
    val, idx = db.coarse(feats)
    assert torch.equal(idx[:, 0].long(), planted)
    assert torch.all(val[:, :-1] >= val[:, 1:])                                 # sorted
    full = db.scores(feats)
    assert torch.equal(full[3:4], db.scores(feats[3:4]))                        # batch invariance
    assert torch.equal(full.gather(1, idx.long()), val)
    kth = val[:, -1:]
    assert torch.all((full > kth).sum(1) <= 99) and torch.all((full >= kth).sum(1) >= 100)
    meshes, scores = db.retrieve(feats)
    assert meshes == [str(int(i)) for i in planted]
