"""Whole-path parity on the GPU, through the reference-facing module (which calls the C-ABI):

  * against the CPU oracle on the same seeded inputs (all stage boundaries the oracle exposes)
  * against the committed goldens (outputs of the real reference head, tests/golden/)
  * size-independent properties at the benchmark size (medium, 8 views, batch 32)

Tolerances (north star): vertex coordinates within 1e-3 relative (fp32, metres), MPJPE against the reference
within 0.1 mm.  The path computes its GEMMs with fp16 operands (11-bit significand, as TF32) and fp32 accumulation.

32-NN selection (blocks 1..NB-1) is DISCONTINUOUS in the regressed coordinates: the goldens contain queries whose
32nd / 33rd neighbours are equidistant to ~1e-6, and any implementation that is not bit-identical to the reference
(including the reference itself on another device) picks the other one there.  The comparison is therefore split into
the two statements that can be exact:
  (A) coordinates: against the oracle run on the neighbour sets the kernels actually used (`neighbours=` hook of the
      oracle, `poem_debug_export_neighbours` of the library): EVERY point within 1e-3 relative, no allowance;
  (B) neighbour sets: each exported set is a valid 32-NN set of the coordinates it was computed from (bit-exactness of
      the search itself on identical coordinates is tests/test_kernels_gpu.py::test_knn32_bit_exact);
and against the un-forced oracle / reference golden every query whose sets agree with the reference's is held to the
same 1e-3, while the queries whose set differs somewhere (reported) are bounded separately.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import poem_oracle as orc  # noqa: E402
from golden_util import CASES, load_case  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402
from poem_v2_b200.config import release_dims  # noqa: E402
from poem_v2_b200.head import POEM_Generalized_Head, PtEmbedTRv4  # noqa: E402

MM = 1e-3  # metres
REL_TOL = 1e-3          # north star: relative error of a regressed point, ||ours - ref|| / ||ref||
FLIP_REL_TOL = 5e-3     # queries whose 32-NN set differs from the reference's (discontinuity), measured <= 1.7e-3
FLIP_FRAC_MAX = 0.25    # cumulative over the four searches of blocks 1-2; measured 4-10 % with the stress weights


def build_head(dims, sd):
    head = POEM_Generalized_Head(dims, template_mesh=synth.standin_template())
    head.load_state_dict(sd, strict=True)
    return head.cuda().eval()


def to_cuda(metas):
    m = dict(metas)
    m["cam_intr"] = metas["cam_intr"].cuda()
    m["cam_extr"] = metas["cam_extr"].cuda()
    return m


def run_with_neighbours(head, feat, metas, ref_j):
    """-> (coords (NB,B,799,3) cpu, neighbour sets (NB-1, 2, B, 799, 32) int64 cpu) of one device run."""
    head.debug_export_neighbours = True
    try:
        out = head(mlvl_feat=feat.cuda(), img_metas=to_cuda(metas), reference_joints=ref_j.cuda(), debug_metas=None)
    finally:
        head.debug_export_neighbours = False
    got = out["all_coords_preds"]
    assert got.dtype == torch.float32 and got.is_cuda
    torch.cuda.synchronize()
    return got.cpu(), head.last_neighbours.long().cpu(), out


def d_mpjpe(ours, ref):
    g = torch.Generator().manual_seed(123)
    gt = ref + 5e-3 * torch.randn(ref.shape, generator=g) / 3 ** 0.5
    mpjpe_ours = (ours - gt).norm(dim=-1)[..., :21].mean(dim=-1)
    mpjpe_ref = (ref - gt).norm(dim=-1)[..., :21].mean(dim=-1)
    return (mpjpe_ours - mpjpe_ref).abs().max().item()


def check_coords(ours, ref, label, mean_mm, flipped=None):
    """ours/ref: (NB,B,799,3) metres.  Every point within REL_TOL of the reference (||ours - ref|| <= 1e-3 ||ref||, i.e.
    ~0.6 mm at 0.6 m), no allowance.  `flipped` (NB,B,799) bool marks the queries whose 32-NN set differs from the
    reference's in this or an earlier block: those are held to FLIP_REL_TOL instead and their fraction is bounded.
    Also: mean point error <= `mean_mm`; MPJPE against a ground truth 5 mm away changes by <= 0.1 mm (`MeanEPE`)."""
    err = (ours - ref).norm(dim=-1)                    # per point, metres
    rel = err / ref.norm(dim=-1)
    mean_err = err.mean(dim=-1)                        # (NB,B)
    dm = d_mpjpe(ours, ref)
    if flipped is None:
        flipped = torch.zeros_like(rel, dtype=torch.bool)
    keep = ~flipped
    rel_same = rel[keep].max().item()
    rel_flip = rel[flipped].max().item() if flipped.any() else 0.0
    print(f"{label}: mean |ours-ref| per block {[round(v, 4) for v in (mean_err.max(dim=1).values / MM).tolist()]} mm, "
          f"worst point {err[keep].max().item() / MM:.3f} mm, rel max {rel_same:.2e}"
          + (f" | queries with a different 32-NN set: {flipped[-1].float().mean().item():.4f}, their rel max {rel_flip:.2e}"
             if flipped.any() else "") + f", |dMPJPE vs GT| {dm / MM:.4f} mm")
    assert dm <= 0.1 * MM
    assert rel_same <= REL_TOL
    assert rel_flip <= FLIP_REL_TOL and flipped[-1].float().mean().item() <= FLIP_FRAC_MAX
    assert mean_err.max().item() <= mean_mm * MM


def flipped_mask(nbr, stages, dims):
    """(NB,B,799) bool: query's 32-NN set (self or cross) differs from the oracle's in block <= i."""
    B = nbr.shape[2]
    m = torch.zeros(dims.n_blocks, B, dims.n_query, dtype=torch.bool)
    for i in range(1, dims.n_blocks):
        for k, key in enumerate(("idx_self", "idx_cross")):
            ref_sets = stages[f"b{i}.{key}"].sort(dim=-1).values
            m[i] |= (nbr[i - 1, k].sort(dim=-1).values != ref_sets).any(dim=-1)
        m[i] |= m[i - 1]
    return m


def check_neighbour_sets(nbr, got, ref_j, dims, pt_xyz):
    """(B): every exported set holds the 32 nearest points of the coordinates the search saw — block i searches
    around the coordinates block i-1 regressed (recovered from the metric output, hence the 1e-5 slack on d^2)."""
    centre = ref_j[:, dims.center_idx][:, None]
    for i in range(1, dims.n_blocks):
        xyz = (got[i - 1] - centre) / dims.radius                      # (B,799,3) normalised
        for k, ref in enumerate((xyz, pt_xyz)):
            d = ((xyz[:, :, None, :] - ref[:, None, :, :]) ** 2).sum(-1)         # (B,799,Lr)
            idx = nbr[i - 1, k]
            assert idx.min() >= 0 and idx.max() < ref.shape[1]
            assert (idx.sort(dim=-1).values.diff(dim=-1) > 0).all(), "duplicate neighbour"
            sel = torch.gather(d, 2, idx)
            assert (sel.diff(dim=-1) >= -1e-5).all(), "neighbours not in ascending distance order"
            rest = d.scatter(2, idx, float("inf"))
            assert (sel.max(dim=-1).values <= rest.min(dim=-1).values + 1e-5).all(), "not the 32 nearest"


@pytest.mark.parametrize("name", CASES)
def test_head_matches_oracle_and_golden(name):
    meta, dims, sd, feat, metas, ref_j, gold = load_case(name)
    bps, a_xyz, a_idx = synth.load_assets()
    tmpl = synth.standin_template()
    head = build_head(dims, sd)
    got, nbr, _ = run_with_neighbours(head, feat, metas, ref_j)
    assert torch.isfinite(got).all()
    st, st_f = {}, {}
    with torch.no_grad():
        want = orc.head_forward(sd, dims, feat, metas, ref_j, tmpl, bps, a_xyz, a_idx, stages=st)
        want_f = orc.head_forward(sd, dims, feat, metas, ref_j, tmpl, bps, a_xyz, a_idx, stages=st_f, neighbours=nbr)
    assert got.shape == want.shape
    mean_mm = 0.05
    # (A) same neighbour sets: every point inside the north-star bound
    check_coords(got, want_f, name + " vs oracle on the kernels' 32-NN sets", mean_mm)
    # (B) the sets themselves
    check_neighbour_sets(nbr, got, ref_j, dims, st["pt_xyz"])
    # un-forced oracle and the reference's own output
    flipped = flipped_mask(nbr, st, dims)
    check_coords(got, want, name + " vs oracle", 0.1, flipped)
    check_coords(got, gold["all_coords_preds"], name + " vs reference golden", 0.1, flipped)
    # normalised offsets from the template (what the decoder actually regresses): relative error of the update
    q0 = st_f["q_xyz"]
    for i in range(dims.n_blocks):
        upd_ref = st_f[f"b{i}.xyz"] - q0
        centre = ref_j[:, dims.center_idx][:, None]
        upd_got = (got[i] - centre) / dims.radius - q0
        num = (upd_got - upd_ref).norm(dim=-1).mean().item()
        den = upd_ref.norm(dim=-1).mean().item()
        print(f"{name} block {i}: mean |d_update| / mean |update| = {num / den:.3e}")
        assert num / den <= 5e-3


@pytest.mark.parametrize("size,views", [("small", [10]), ("small", [1]), ("medium", [1, 10, 5, 7]), ("small", [6, 9, 3])])
def test_view_count_edge_cases_match_oracle(size, views):
    """1 view (single-view merge path), the maximum of 10 views and view counts that do not divide 128 (ragged rows
    of the merge GEMMs), checked against the pinned oracle."""
    dims = release_dims(size)
    sd = synth.make_state_dict(dims, 21, "init")
    feat, metas, ref_j = synth.make_inputs(dims, len(views), views, 5)
    bps, a_xyz, a_idx = synth.load_assets()
    head = build_head(dims, sd)
    got, nbr, _ = run_with_neighbours(head, feat, metas, ref_j)
    with torch.no_grad():
        want_f = orc.head_forward(sd, dims, feat, metas, ref_j, synth.standin_template(), bps, a_xyz, a_idx, neighbours=nbr)
    check_coords(got, want_f, f"{size} views={views} (kernels' 32-NN sets)", 0.05)
    st = {}
    with torch.no_grad():
        want = orc.head_forward(sd, dims, feat, metas, ref_j, synth.standin_template(), bps, a_xyz, a_idx, stages=st)
    check_coords(got, want, f"{size} views={views}", 0.05, flipped_mask(nbr, st, dims))


def test_transformer_module_matches_oracle():
    dims = release_dims("small")
    sd = synth.make_state_dict(dims, 11)
    B = 2
    g = torch.Generator().manual_seed(5)
    pt_feats = torch.randn(B, dims.n_sample, dims.embed_dims, generator=g)
    bps, a_xyz, a_idx = synth.load_assets()
    pt_xyz = (bps / dims.radius)[None].repeat(B, 1, 1)
    q_xyz = (synth.standin_template() / dims.radius)[None].repeat(B, 1, 1) + 0.01 * torch.randn(B, 799, 3, generator=g)
    q_feat = sd["query_feat_embedding.weight"][None].expand(B, -1, -1)
    xyz_all = []
    qf, qx = q_feat, q_xyz
    with torch.no_grad():
        for i in range(dims.n_blocks):
            qf, qx = orc.metro_block(sd, i, dims, qx, qf, pt_xyz, pt_feats, (a_xyz, a_idx))
            xyz_all.append(qx)
    want = torch.stack(xyz_all)
    tr = PtEmbedTRv4(dims)
    tr.load_state_dict({k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}, strict=True)
    tr = tr.cuda().eval()
    got, pose, shape = tr(query_xyz=q_xyz.cuda(), query_feat=q_feat.cuda(), pt_xyz=pt_xyz.cuda(), pt_feats=pt_feats.cuda())
    assert pose is None and shape is None
    err = (got.cpu() - want).norm(dim=-1)
    print("transformer-only: mean err (radius units)", err.mean().item(), "max", err.max().item())
    assert err.mean().item() <= 1e-3 and err.max().item() <= 5e-2


def test_host_buffer_entry_point_matches_device_entry_point():
    dims = release_dims("small")
    sd = synth.make_state_dict(dims, 0)
    feat, metas, ref_j = synth.make_inputs(dims, 2, [2, 3], 9)
    head = build_head(dims, sd)
    dev_out = head(mlvl_feat=feat.cuda(), img_metas=to_cuda(metas), reference_joints=ref_j.cuda())["all_coords_preds"]
    pin = lambda t: t.contiguous().pin_memory()  # noqa: E731
    hm = dict(metas)
    hm["cam_intr"], hm["cam_extr"] = pin(metas["cam_intr"]), pin(metas["cam_extr"])
    host_out = head.forward_host(pin(feat), hm, pin(ref_j))
    torch.cuda.synchronize()
    assert torch.equal(host_out, dev_out.cpu())          # same kernels, same order: bit-identical


def test_host_entry_point_pipelines_calls_of_different_shapes():
    """Consecutive host-buffer calls whose batch / view counts differ (last partial batch of an epoch, ragged views)
    while the previous call is still running: the staging slots sit at shape-independent offsets, so no call may
    overwrite the inputs of the one before it (slot 1 of a small call used to land inside slot 0 of a
    large one)."""
    dims = release_dims("small")
    sd = synth.make_state_dict(dims, 0)
    head = build_head(dims, sd)
    pin = lambda t: t.contiguous().pin_memory()  # noqa: E731
    shapes = [[4, 4, 4, 4], [2], [3, 1, 2], [1], [4, 4, 4, 4], [2, 2]]
    calls = []
    for i, views in enumerate(shapes):
        feat, metas, ref_j = synth.make_inputs(dims, len(views), views, 30 + i)
        dev_out = head(mlvl_feat=feat.cuda(), img_metas=to_cuda(metas), reference_joints=ref_j.cuda())["all_coords_preds"].cpu()
        hm = dict(metas)
        hm["cam_intr"], hm["cam_extr"] = pin(metas["cam_intr"]), pin(metas["cam_extr"])
        calls.append((pin(feat), hm, pin(ref_j), dev_out))
    torch.cuda.synchronize()
    head.forward_host(*calls[0][:3])          # sizes the staging buffer for the largest call
    torch.cuda.synchronize()
    for rep in range(3):                      # back to back, no synchronisation in between
        outs = [head.forward_host(f, m, r) for f, m, r, _ in calls]
        torch.cuda.synchronize()
        for o, (_, _, _, want) in zip(outs, calls):
            assert torch.equal(o, want)


@pytest.mark.parametrize("name", ["small_ragged_b3", "medium_v8_b1", "small_v1_b2", "large_v2_b1"])
def test_merged_bps_features_match_oracle(name):
    """Stage boundary a6 on its own (input projection + positional term, projection, bilinear
    sampling, the raw `.view` regroup, merge network): the merged BPS features `pt_feats` the kernels hand to the decoder
    blocks (`poem_debug_export_pt_feats`, fp16) against the oracle's (ptEmb_head.py:926) on the golden inputs —
    fused sampler/merge kernel for D <= 256, the four-kernel chain for POEM-large."""
    from poem_v2_b200 import _native as nat
    meta, dims, sd, feat, metas, ref_j, gold = load_case(name)
    bps, a_xyz, a_idx = synth.load_assets()
    st = {}
    with torch.no_grad():
        orc.head_forward(sd, dims, feat, metas, ref_j, synth.standin_template(), bps, a_xyz, a_idx, stages=st)
    want = st["pt_feats"]                                     # (B, 4096, D) fp32
    head = build_head(dims, sd)
    buf = torch.zeros(want.shape, dtype=torch.float16, device="cuda")
    lib = nat.load()
    lib.poem_debug_export_pt_feats(buf.data_ptr(), buf.numel())
    try:
        head(mlvl_feat=feat.cuda(), img_metas=to_cuda(metas), reference_joints=ref_j.cuda())
        torch.cuda.synchronize()
    finally:
        lib.poem_debug_export_pt_feats(None, 0)
    got = buf.float().cpu()
    err = (got - want).abs()
    scale = want.abs().max().item()
    rel_l2 = ((got - want).norm() / want.norm()).item()
    print(f"{name}: pt_feats max |err| {err.max().item():.4e} ({err.max().item() / scale:.2e} of the range), rel-L2 {rel_l2:.2e}")
    # fp16 operands through input_proj, the sampled rows and the two merge MLPs, fp16 storage of the result (2^-11)
    assert rel_l2 <= 1e-3 and err.max().item() <= 2e-3 * scale        # measured 3.7e-4 ... 6.2e-4 / 5.5e-4 ... 7.4e-4


@pytest.mark.parametrize("size,views", [("small", [1, 3, 7, 10, 8]), ("medium", [8, 8]), ("medium", [5, 1, 9])])
def test_fused_sampler_merge_matches_the_unfused_chain(size, views):
    """sample_merge_kernel (sampler + merge MLP0 + cross-view reduce in one kernel, row tiles of floor(128 / N) tokens,
    tiles that straddle channel planes / views when N does not divide 128) against the four-kernel chain it replaces
    (project_sample -> GEMM -> GEMM -> merge_reduce, `poem_debug_force_unfused`), on the device, whole path."""
    from poem_v2_b200 import _native as nat
    dims = release_dims(size)
    sd = synth.make_state_dict(dims, 3, "stress")
    feat, metas, ref_j = synth.make_inputs(dims, len(views), views, 11)
    head = build_head(dims, sd)
    lib = nat.load()
    fused = head(mlvl_feat=feat.cuda(), img_metas=to_cuda(metas), reference_joints=ref_j.cuda())["all_coords_preds"].cpu()
    lib.poem_debug_force_unfused(1)
    try:
        chain = head(mlvl_feat=feat.cuda(), img_metas=to_cuda(metas), reference_joints=ref_j.cuda())["all_coords_preds"].cpu()
    finally:
        lib.poem_debug_force_unfused(0)
    # block 0 has no 32-NN search: the two paths differ only by rounding (Mm stays fp32 in the fused kernel)
    err0 = (fused[0] - chain[0]).norm(dim=-1)
    print(f"{size} views={views}: block 0 fused vs chain: mean {err0.mean().item() * 1e3:.5f} mm, max {err0.max().item() * 1e3:.5f} mm")
    assert torch.isfinite(fused).all()
    # measured (stress weights): mean 0.012 mm, max 0.125 mm; both paths pass the oracle comparisons above
    assert err0.max().item() <= 0.3 * MM and err0.mean().item() <= 0.03 * MM


def test_deterministic_and_batch_independent():
    """Samples are independent units: a sample's output must not depend on its batch neighbours."""
    dims = release_dims("small")
    sd = synth.make_state_dict(dims, 0)
    head = build_head(dims, sd)
    feat, metas, ref_j = synth.make_inputs(dims, 3, [2, 2, 2], 21)
    full = head(mlvl_feat=feat.cuda(), img_metas=to_cuda(metas), reference_joints=ref_j.cuda())["all_coords_preds"].cpu()
    again = head(mlvl_feat=feat.cuda(), img_metas=to_cuda(metas), reference_joints=ref_j.cuda())["all_coords_preds"].cpu()
    assert torch.equal(full, again)
    m1 = {"inp_img_shape": (256, 256), "cam_intr": metas["cam_intr"][2:4].cuda(), "cam_extr": metas["cam_extr"][2:4].cuda(),
          "master_id": [0], "cam_view_num": np.array([2])}
    one = head(mlvl_feat=feat[2:4].cuda(), img_metas=m1, reference_joints=ref_j[1:2].cuda())["all_coords_preds"].cpu()
    assert torch.equal(one[:, 0], full[:, 1])


def test_benchmark_size_properties():
    """medium, 8 views, batch 32 (the bench workload): finite, deterministic, and equal to the same samples run
    in four shards of 8 (the multi-GPU sharding rule) — a size-independent check the oracle need not run."""
    dims = release_dims("medium")
    sd = synth.make_state_dict(dims, 0)
    head = build_head(dims, sd)
    B, V = 32, 8
    feat, metas, ref_j = synth.make_inputs(dims, B, V, 1)
    full = head(mlvl_feat=feat.cuda(), img_metas=to_cuda(metas), reference_joints=ref_j.cuda())["all_coords_preds"]
    torch.cuda.synchronize()
    assert torch.isfinite(full).all()
    centre = ref_j[:, dims.center_idx].cuda()
    assert ((full - centre[None, :, None]).norm(dim=-1) < 1.0).all()      # stays near the hand (metres)
    for s in range(4):
        sl = slice(8 * s, 8 * s + 8)
        m = {"inp_img_shape": (256, 256), "cam_intr": metas["cam_intr"][64 * s:64 * s + 64].cuda(),
             "cam_extr": metas["cam_extr"][64 * s:64 * s + 64].cuda(), "master_id": [0] * 8,
             "cam_view_num": np.array([V] * 8)}
        part = head(mlvl_feat=feat[64 * s:64 * s + 64].cuda(), img_metas=m, reference_joints=ref_j[sl].cuda())
        assert torch.equal(part["all_coords_preds"], full[:, sl])


def test_errors_are_loud():
    dims = release_dims("small")
    head = POEM_Generalized_Head(dims, template_mesh=synth.standin_template()).cuda()
    feat, metas, ref_j = synth.make_inputs(dims, 1, [2], 1)
    with pytest.raises(Exception):
        head(mlvl_feat=feat, img_metas=metas, reference_joints=ref_j)          # CPU tensors: no fallback
    metas_bad = to_cuda(metas)
    metas_bad["master_id"] = [1]
    with pytest.raises(AssertionError):
        head(mlvl_feat=feat.cuda(), img_metas=metas_bad, reference_joints=ref_j.cuda())
    # POEM-huge (D = 1024, head dim 256) is not built: rejected with a message, never computed approximately
    from poem_v2_b200 import _native as nat
    huge = POEM_Generalized_Head(release_dims("huge"), template_mesh=synth.standin_template()).cuda()
    hf, hm, hr = synth.make_inputs(release_dims("huge"), 1, [2], 1)
    with pytest.raises(nat.PoemError, match="POEM-huge"):
        huge(mlvl_feat=hf.cuda(), img_metas=to_cuda(hm), reference_joints=hr.cuda())
    del huge
    metas11 = to_cuda(metas)
    metas11["cam_view_num"] = np.array([11])
    with pytest.raises(Exception):
        head(mlvl_feat=feat.cuda().repeat(6, 1, 1, 1)[:11], img_metas=metas11, reference_joints=ref_j.cuda())
