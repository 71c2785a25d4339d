"""HRNet-W40 stage 4: implicit-GEMM convolutions and the whole stage against the fp32 oracle."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import poem_oracle as orc  # noqa: E402
from poem_v2_b200 import _native as nat  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402
from poem_v2_b200.hrnet import HRNetStage4  # noqa: E402


def _p(t):
    return None if t is None else t.data_ptr()


@pytest.mark.parametrize("live", [False, True])
@pytest.mark.parametrize("N,R,cin,cout,k,stride,relu,res", [
    (2, 64, 40, 40, 3, 1, True, False),
    (2, 32, 80, 80, 3, 1, True, True),
    (3, 16, 160, 160, 3, 1, False, True),
    (3, 8, 320, 320, 3, 1, True, True),      # 64 pixels per image: two images per 128-row tile, odd image count
    (2, 64, 40, 80, 3, 2, False, False),     # fuse-layer downsampling, last conv of a chain
    (2, 32, 40, 40, 3, 2, True, False),
    (2, 16, 160, 320, 3, 2, False, False),
    (2, 8, 320, 40, 1, 1, False, False),     # fuse-layer 1x1 (upsampled later)
    (2, 32, 80, 40, 1, 1, False, False),
    (3, 32, 40, 40, 3, 1, True, True),       # halo kernel, other map sizes / channel splits
    (2, 16, 80, 80, 3, 1, False, True),
    (1, 16, 33, 33, 3, 1, True, False),      # 33 live channels -> 48 (32 + 16) of 64
    (2, 32, 150, 150, 3, 1, True, True),     # 150 -> 160 (64 + 64 + 32) of 192
    (2, 16, 128, 128, 3, 1, True, True),     # no padding at all
    (2, 32, 256, 40, 3, 1, True, False),     # transition1: 256 -> 40 (Cin != Cout on the halo kernel)
    (2, 32, 240, 80, 3, 1, True, False),     # uv_decode: cat(160, 80) -> 80
    (3, 16, 120, 40, 3, 1, True, False),     # uv_decode: cat(80, 40) -> 40
])
def test_conv_nhwc(N, R, cin, cout, k, stride, relu, res, live):
    lib = nat.load()
    g = torch.Generator().manual_seed(R * 7 + cin + cout + k)
    cin_p, cout_p = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64
    x = torch.randn(N, cin, R, R, generator=g).half().float()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).half().float()
    b = 0.1 * torch.randn(cout, generator=g)
    Ro = R // stride
    r = torch.randn(N, cout, Ro, Ro, generator=g).half().float() if res else None
    xp = torch.zeros(N, R, R, cin_p)
    xp[..., :cin] = x.permute(0, 2, 3, 1)
    wp = torch.zeros(cout_p, k, k, cin_p)
    wp[:cout, :, :, :cin] = w.permute(0, 2, 3, 1)
    bp = torch.zeros(cout_p)
    bp[:cout] = b
    rp = None
    if res:
        rp = torch.zeros(N, Ro, Ro, cout_p)
        rp[..., :cout] = r.permute(0, 2, 3, 1)
    d = [t.half().contiguous().cuda() if t is not None else None for t in (xp, wp.reshape(cout_p, -1), rp)]
    bd = bp.cuda()
    out = torch.full((N, Ro, Ro, cout_p), float("nan"), device="cuda", dtype=torch.float16)
    # live=True promises that channels >= cin / cout are zero padding: the 3x3 stride-1 kernel then multiplies only
    # the live channels (mixed 128/64/32-byte swizzle K blocks) and writes the padding as zeros
    nat.check(lib.poem_conv_nhwc(_p(d[0]), N, R, R, cin_p, _p(d[1]), _p(bd), cout_p, k, stride, int(relu), _p(d[2]),
                                 _p(out), cin if live else 0, cout if live else 0,
                                 torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = F.conv2d(x, w, b, stride=stride, padding=k // 2)
    if res:
        ref = ref + r
    if relu:
        ref = F.relu(ref)
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    assert (got[..., cout:] == 0).all()                                  # padded channels stay exactly zero
    err = (got[..., :cout].permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item())                # fp16 output rounding (2^-11), fp32 accumulation


@pytest.mark.parametrize("N", [2, 5])
def test_stage4_matches_oracle(N):
    sd = synth.make_stage4_state_dict(0)
    xs = synth.make_stage4_inputs(N, 64, 1)
    with torch.no_grad():
        want = orc.hrnet_stage4(sd, xs)
    m = HRNetStage4()
    m.load_state_dict(sd, strict=True)
    got = m([x.cuda() for x in xs])
    for b, (g_, w_) in enumerate(zip(got, want)):
        g_ = g_.cpu()
        assert g_.shape == w_.shape and torch.isfinite(g_).all()
        err = (g_ - w_).abs()
        scale = w_.abs().max().item()
        print(f"stage4 N={N} branch {b}: max err {err.max().item():.4f} mean err {err.mean().item():.5f} "
              f"(|ref| max {scale:.2f}, mean {w_.abs().mean().item():.3f})")
        # 3 modules x (8 fp16 convs per branch + fuse) with fp16 activations in between
        assert err.max().item() <= 4e-3 * scale       # measured 1.2e-3 / 1e-4 of the range (fp16 activations between layers)
        assert err.mean().item() <= 4e-4 * scale


def test_stage4_matches_reference_golden():
    import ast
    import os
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hrnet_stage4_n2.npz"))
    meta = ast.literal_eval(str(z["meta"]))
    sd = synth.make_stage4_state_dict(meta["wseed"])
    xs = synth.make_stage4_inputs(meta["n_images"], 64, meta["iseed"])
    m = HRNetStage4()
    m.load_state_dict(sd, strict=True)
    got = [g.cpu() for g in m([x.cuda() for x in xs])]
    want = [torch.from_numpy(z[f"y{b}"]) for b in range(4)]
    sub = [got[0][:, :, ::4, ::4], got[1][:, :, ::2, ::2], got[2], got[3]]
    for g_, w_ in zip(sub, want):
        assert (g_ - w_).abs().max().item() <= 4e-3 * w_.abs().max().item()


# ------------------------------------------------------------------------------------------------ whole backbone (f1)
def _check_maps(got, want, label, max_frac=6e-3, mean_frac=6e-4):
    for b, (g_, w_) in enumerate(zip(got, want)):
        assert g_.shape == w_.shape and torch.isfinite(g_).all()
        err = (g_ - w_).abs()
        scale = w_.abs().max().item()
        rel_l2 = ((g_ - w_).norm() / w_.norm()).item()
        print(f"{label} branch {b}: max err {err.max().item():.4f} mean err {err.mean().item():.5f} rel-L2 {rel_l2:.2e} "
              f"(|ref| max {scale:.2f}, mean {w_.abs().mean().item():.3f})")
        # ~150 fp16 convolutions deep with fp16 activations in between: measured max 2e-3 / mean 2e-4 of the range, rel-L2 1e-3
        assert err.max().item() <= max_frac * scale
        assert err.mean().item() <= mean_frac * scale
        assert rel_l2 <= 3e-3


@pytest.mark.parametrize("N", [1, 3])
def test_backbone_matches_oracle(N):
    from poem_v2_b200.hrnet import HRNetW40
    sd = synth.make_backbone_state_dict(0)
    img = synth.make_images(N, 256, 1)
    with torch.no_grad():
        want = orc.hrnet_forward(sd, img)
    m = HRNetW40()
    m.load_state_dict(sd, strict=True)
    got = [g.cpu() for g in m(img.cuda())]
    _check_maps(got, want, f"backbone N={N}")


def test_backbone_matches_reference_golden():
    import ast
    import os
    import numpy as np
    from poem_v2_b200.hrnet import HRNetW40
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hrnet_w40_n1.npz"))
    meta = ast.literal_eval(str(z["meta"]))
    sd = synth.make_backbone_state_dict(meta["wseed"])
    sd["classifier.weight"] = torch.zeros(1000, 2048)        # dead classification-head key: accepted and dropped
    m = HRNetW40()
    m.load_state_dict(sd, strict=True)
    got = [g.cpu() for g in m(synth.make_images(meta["n_images"], 256, meta["iseed"]).cuda())]
    want = [torch.from_numpy(z[f"y{b}"]) for b in range(4)]
    sub = [got[0][:, :, ::4, ::4], got[1][:, :, ::2, ::2], got[2], got[3]]
    _check_maps(sub, want, "backbone golden")


def test_backbone_deterministic():
    from poem_v2_b200.hrnet import HRNetW40
    sd = synth.make_backbone_state_dict(3)
    m = HRNetW40()
    m.load_state_dict(sd, strict=True)
    img = synth.make_images(2, 256, 5).cuda()
    a = m(img)
    b = m(img)
    for x, y in zip(a, b):
        assert torch.equal(x, y)          # deterministic


def test_backbone_rejects_bad_input():
    from poem_v2_b200.hrnet import HRNetW40
    m = HRNetW40()
    with pytest.raises(nat.PoemError):
        m(torch.zeros(1, 3, 256, 256))                  # CPU tensor: no CPU implementation
    with pytest.raises(nat.PoemError):
        m(torch.zeros(1, 3, 224, 224, device="cuda"))   # only the 256x256 release resolution
    with pytest.raises(RuntimeError):
        m.load_state_dict({"conv1.weight": torch.zeros(64, 3, 3, 3)}, strict=True)


# ------------------------------------------------------------------------------------------------ images -> mlvl_feat
@pytest.mark.parametrize("N", [1, 4])
def test_image_stage_matches_oracle(N):
    from poem_v2_b200.hrnet import ImageStage
    sd = synth.make_image_stage_state_dict(0)
    img = synth.make_images(N, 256, 1)
    with torch.no_grad():
        want, want_maps = orc.image_features(sd, img)
    m = ImageStage()
    m.load_state_dict(sd, strict=True)
    res = m(img.cuda(), return_maps=True)
    got, maps = res["mlvl_feat"], res["img_feats"]
    _check_maps([got.cpu()], [want], f"mlvl_feat N={N}")
    _check_maps([x.cpu() for x in maps], want_maps, f"image stage maps N={N}")
    again = m(img.cuda())                       # without the map export: same features
    assert torch.equal(again, got)


def test_image_stage_matches_reference_golden():
    import ast
    import os
    import numpy as np
    from poem_v2_b200.hrnet import ImageStage
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "image_stage_n3.npz"))
    meta = ast.literal_eval(str(z["meta"]))
    n = meta["n_images"]
    sd = synth.make_image_stage_state_dict(meta["wseed"])
    sd["ptEmb_head.input_proj.weight"] = torch.zeros(4)       # other parts of the full-model checkpoint: ignored
    sd["uv_in.conv.bias"] = torch.zeros(80)
    m = ImageStage()
    m.load_state_dict(sd, strict=True)
    res = m(synth.make_images(n, 256, meta["iseed"]).cuda(), return_uv=True)
    _check_maps([res["mlvl_feat"].cpu()], [torch.from_numpy(z["mlvl_feat"])], "mlvl_feat golden")
    uv = res["pred_joints_uv"].cpu()
    duv = (uv - torch.from_numpy(z["pred_joints_uv"])).abs()
    print(f"pred_joints_uv golden: max |d| {duv.max().item():.4f} px, mean {duv.mean().item():.4f} px")
    assert duv.max().item() <= 0.2 and duv.mean().item() <= 0.03          # 256-px image, fp16 conv stack (measured 0.057 / 0.008 px)
    intr, extr = synth.make_cameras(1, [n], meta["iseed"])
    rj = ImageStage.triangulate(res["pred_joints_uv"], intr, extr, [n]).cpu()
    drj = (rj - torch.from_numpy(z["ref_joints"])).norm(dim=-1)
    print(f"ref_joints golden: max {drj.max().item() * 1e3:.4f} mm")
    assert drj.max().item() <= 1.5e-4                                     # 1 px at 0.6 m and f = 900 is 0.67 mm; measured 0.035 mm


@pytest.mark.parametrize("views", [[2], [1, 3, 8], [10, 2]])
def test_triangulate_dlt_matches_oracle(views):
    """Jacobi-SVD DLT kernel vs the oracle's torch.linalg.svd on projections of known joints (+ pixel noise)."""
    from poem_v2_b200.hrnet import ImageStage
    g = torch.Generator().manual_seed(sum(views))
    B = len(views)
    intr, extr = synth.make_cameras(B, views, 3)
    joints = torch.tensor([0.0, 0.0, 0.6]) + 0.05 * torch.randn(B, 21, 3, generator=g)
    T = torch.linalg.inv(extr)
    uv, start = [], 0
    for b, v in enumerate(views):
        for i in range(start, start + v):
            pc = joints[b] @ T[i, :3, :3].T + T[i, :3, 3]
            px = pc @ intr[i].T
            uv.append(px[:, :2] / px[:, 2:3])
        start += v
    uv = torch.stack(uv) + 0.5 * torch.randn(sum(views), 21, 2, generator=g)
    want = orc.triangulate_dlt(uv, intr, extr, views)
    got = ImageStage.triangulate(uv.cuda(), intr.cuda(), extr.cuda(), views).cpu()
    err = (got - want).norm(dim=-1).max().item()
    print(f"DLT views={views}: max |ours - svd| = {err * 1e3:.5f} mm; |svd - truth| = {(want - joints).norm(dim=-1).max().item() * 1e3:.3f} mm")
    assert err <= 2e-5          # fp32 SVD in the oracle vs fp64 Jacobi here: 0.02 mm


@pytest.mark.parametrize("N", [2])
def test_heatmap_stage_matches_oracle(N):
    from poem_v2_b200.hrnet import ImageStage
    sd = synth.make_image_stage_state_dict(1)
    sd["uv_out.conv.weight"] *= 0.2            # unsaturated heatmaps
    img = synth.make_images(N, 256, 4)
    with torch.no_grad():
        _, maps = orc.image_features(sd, img)
        want_uv, want_h = orc.uv_decode_heatmap(sd, maps)
    m = ImageStage()
    m.load_state_dict(sd, strict=True)
    res = m(img.cuda(), return_uv=True, return_heatmap=True)
    dh = (res["uv_hmap"].cpu() - want_h).abs()
    duv = (res["pred_joints_uv"].cpu() - want_uv).abs()
    print(f"heatmap N={N}: max |dh| {dh.max().item():.4f} mean {dh.mean().item():.5f}; uv max {duv.max().item():.4f} px")
    assert dh.mean().item() <= 1e-3 and dh.max().item() <= 1e-2          # measured 1.5e-4 / 1.1e-3
    assert duv.max().item() <= 0.05                                      # measured 0.0125 px


def test_images_to_mesh_pipeline():
    """ImageStage -> POEM_Generalized_Head on the device, against oracle.image_features -> oracle.head_forward."""
    from poem_v2_b200.config import release_dims
    from poem_v2_b200.head import POEM_Generalized_Head
    from poem_v2_b200.hrnet import ImageStage
    dims = release_dims("small")
    B, V = 1, 2
    sd_img = synth.make_image_stage_state_dict(0)
    sd_img["feat_in.conv.weight"] *= 0.1      # unit-scale mlvl_feat, the range the synthetic head weights are made for
    sd_img["feat_in.conv.bias"] *= 0.1
    sd_head = synth.make_state_dict(dims, 0)
    _, metas, ref_j = synth.make_inputs(dims, B, [V], 1)
    img = synth.make_images(B * V, 256, 2)
    bps, a_xyz, a_idx = synth.load_assets()
    with torch.no_grad():
        feat_ref, _ = orc.image_features(sd_img, img)
        want = orc.head_forward(sd_head, dims, feat_ref, metas, ref_j, synth.standin_template(), bps, a_xyz, a_idx)
    stage = ImageStage()
    stage.load_state_dict(sd_img, strict=True)
    head = POEM_Generalized_Head(dims, template_mesh=synth.standin_template())
    head.load_state_dict(sd_head, strict=True)
    head = head.cuda().eval()
    m = dict(metas)
    m["cam_intr"], m["cam_extr"] = metas["cam_intr"].cuda(), metas["cam_extr"].cuda()
    feat = stage(img.cuda())
    got = head(mlvl_feat=feat, img_metas=m, reference_joints=ref_j.cuda())["all_coords_preds"].cpu()
    err = (got - want).norm(dim=-1)
    print(f"images->mesh: mean |ours - oracle| = {err.mean().item() * 1e3:.4f} mm, max {err.max().item() * 1e3:.3f} mm")
    assert torch.isfinite(got).all() and err.mean().item() * 1e3 <= 0.1   # north star: MPJPE within 0.1 mm (measured 0.06 mm)
