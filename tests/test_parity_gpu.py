"""Whole-path parity on the GPU, through the reference-facing module (which calls the C-ABI):

  * against the CPU oracle on the same seeded inputs (all stage boundaries the oracle exposes)
  * against the committed goldens (outputs of the real reference head, tests/golden/)
  * size-independent properties at the benchmark size (medium, 8 views, batch 32)

Tolerances (north star): vertex coordinates within 1e-3 relative (fp32, metres), MPJPE against the reference
within 0.1 mm.  The path computes its GEMMs in bf16 with fp32 accumulation; 32-NN selection is discontinuous, so a
small fraction of queries whose 32nd/33rd neighbours are nearly equidistant may pick the other one — bounded below.
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


def build_head(dims, sd):
    head = POEM_Generalized_Head(dims, template_mesh=synth.standin_template())
    head.load_state_dict(sd, strict=True)
    return head.cuda().eval()


def to_cuda(metas):
    m = dict(metas)
    m["cam_intr"] = metas["cam_intr"].cuda()
    m["cam_extr"] = metas["cam_extr"].cuda()
    return m


def check_coords(ours, ref, label, mean_mm):
    """ours/ref: (NB,B,799,3) metres.

    Tolerances (north star: 1e-3 relative on fp32 vertex coordinates, MPJPE within 0.1 mm of the reference):
      * per point ||ours - ref|| <= 1e-3 * ||ref|| (~0.6 mm at 0.6 m) for >= 95 % of the points (measured: 96.1 % for
        POEM-large, >= 99.3 % for small/medium with stress weights, 100 % with reference-style initialisation); the
        worst point (a query whose 32-NN set or near-one-hot softmax flipped) <= 1e-2 * ||ref||
      * mean point error <= `mean_mm` (bf16 operands, fp32 accumulation; measured 0.05 mm with reference-style
        initialisation, 0.10-0.28 mm with the O(1)-everywhere "stress" weights whose per-block updates are ~8 mm)
      * MPJPE against a ground truth 5 mm away from the reference changes by <= 0.1 mm (what `MeanEPE` reports)
    """
    err = (ours - ref).norm(dim=-1)                    # per point, metres
    rel = err / ref.norm(dim=-1)
    mean_err = err.mean(dim=-1)                        # (NB,B)
    g = torch.Generator().manual_seed(123)
    gt = ref + 5e-3 * torch.randn(ref.shape, generator=g) / 3 ** 0.5
    mpjpe_ours = (ours - gt).norm(dim=-1)[..., :21].mean(dim=-1)
    mpjpe_ref = (ref - gt).norm(dim=-1)[..., :21].mean(dim=-1)
    d_mpjpe = (mpjpe_ours - mpjpe_ref).abs().max().item()
    print(f"{label}: mean |ours-ref| per block {[round(v, 4) for v in (mean_err.max(dim=1).values / MM).tolist()]} mm, "
          f"worst point {err.max().item() / MM:.3f} mm, frac(rel>1e-3) {(rel > 1e-3).float().mean().item():.4f}, "
          f"rel max {rel.max().item():.2e}, |dMPJPE vs GT| {d_mpjpe / MM:.4f} mm")
    assert d_mpjpe <= 0.1 * MM
    assert (rel > 1e-3).float().mean().item() <= 0.05
    assert rel.max().item() <= 1e-2   # worst observed: 5.2e-3 (2.4 mm, POEM-large, one query whose 32-NN set flips)
    assert mean_err.max().item() <= mean_mm * MM


@pytest.mark.parametrize("name", CASES)
def test_head_matches_oracle_and_golden(name):
    meta, dims, sd, feat, metas, ref_j, gold = load_case(name)
    bps, a_xyz, a_idx = synth.load_assets()
    st = {}
    with torch.no_grad():
        want = orc.head_forward(sd, dims, feat, metas, ref_j, synth.standin_template(), bps, a_xyz, a_idx, stages=st)
    head = build_head(dims, sd)
    out = head(mlvl_feat=feat.cuda(), img_metas=to_cuda(metas), reference_joints=ref_j.cuda(), debug_metas=None)
    got = out["all_coords_preds"]
    assert got.shape == want.shape and got.dtype == torch.float32 and got.is_cuda
    got = got.cpu()
    assert torch.isfinite(got).all()
    mean_mm = 0.1 if meta["mode"] == "init" else 0.35
    check_coords(got, want, name + " vs oracle", mean_mm)
    check_coords(got, gold["all_coords_preds"], name + " vs reference golden", mean_mm)
    # normalised offsets from the template (what the decoder actually regresses): relative error of the update
    tmpl = st["q_xyz"]
    for i in range(dims.n_blocks):
        upd_ref = st[f"b{i}.xyz"] - tmpl
        centre = ref_j[:, dims.center_idx][:, None]
        upd_got = (got[i] - centre) / dims.radius - tmpl
        num = (upd_got - upd_ref).norm(dim=-1).mean().item()
        den = upd_ref.norm(dim=-1).mean().item()
        print(f"{name} block {i}: mean |d_update| / mean |update| = {num / den:.3e}")
        assert num / den <= (1e-2 if meta["mode"] == "init" else 5e-2)


@pytest.mark.parametrize("size,views", [("small", [10]), ("small", [1]), ("medium", [1, 10, 5, 7]), ("small", [6, 9, 3])])
def test_view_count_edge_cases_match_oracle(size, views):
    """1 view (single-view merge path), the maximum of 10 views and view counts that do not divide 128 (ragged rows
    of the merge GEMMs), checked against the pinned oracle."""
    dims = release_dims(size)
    sd = synth.make_state_dict(dims, 21, "init")
    feat, metas, ref_j = synth.make_inputs(dims, len(views), views, 5)
    bps, a_xyz, a_idx = synth.load_assets()
    with torch.no_grad():
        want = orc.head_forward(sd, dims, feat, metas, ref_j, synth.standin_template(), bps, a_xyz, a_idx)
    head = build_head(dims, sd)
    got = head(mlvl_feat=feat.cuda(), img_metas=to_cuda(metas), reference_joints=ref_j.cuda())["all_coords_preds"].cpu()
    check_coords(got, want, f"{size} views={views}", 0.1)


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
    metas11 = to_cuda(metas)
    metas11["cam_view_num"] = np.array([11])
    with pytest.raises(Exception):
        head(mlvl_feat=feat.cuda().repeat(6, 1, 1, 1)[:11], img_metas=metas11, reference_joints=ref_j.cuda())
