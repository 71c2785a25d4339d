"""TEST INFRASTRUCTURE — writes tests/golden/*.npz by running the UNMODIFIED reference head.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
Inputs and weights are pure functions of seeds (`poem_v2_b200.synth`), so only OUTPUTS of the
reference are stored: final coordinates plus strided slices of the stage boundaries.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_shim  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402
from poem_v2_b200.config import release_dims  # noqa: E402

# name -> (size, views, weight seed, input seed, weight mode)
CASES = {
    "small_v2_b1": ("small", [2], 0, 1, "stress"),
    "small_v1_b2": ("small", [1, 1], 0, 2, "stress"),
    "small_ragged_b3": ("small", [3, 1, 2], 3, 4, "stress"),
    "small_v4_b2_init": ("small", [4, 4], 5, 6, "init"),
    "medium_v8_b1": ("medium", [8], 0, 1, "stress"),
    "large_v2_b1": ("large", [2], 0, 1, "stress"),
}
ROW_STRIDE = 37


def run_reference(size, views, wseed, iseed, mode):
    dims = release_dims(size)
    head, _ = ref_shim.build_reference_head(size, template_fn=synth.standin_template)
    sd = synth.make_state_dict(dims, wseed, mode)
    missing, unexpected = head.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    live = set(sd)
    assert live.isdisjoint(missing)
    feat, metas, ref_j = synth.make_inputs(dims, len(views), views, iseed)
    cap = {}
    tr = head.transformer

    def pre(mod, args, kwargs):
        cap["pt_feats"] = kwargs["pt_feats"].detach().clone()
        cap["pt_xyz"] = kwargs["pt_xyz"].detach().clone()
        cap["q_xyz"] = kwargs["query_xyz"].detach().clone()
    h = [tr.register_forward_pre_hook(pre, with_kwargs=True)]
    for i, blk in enumerate(tr.pt_metro_encoder):
        def post(mod, args, out, i=i):
            cap[f"b{i}.out"] = out[0].detach().clone()
            cap[f"b{i}.xyz"] = out[1].detach().clone()
        h.append(blk.register_forward_hook(post))
    with torch.no_grad():
        res = head(mlvl_feat=feat, img_metas=metas, reference_joints=ref_j, debug_metas=None)
    for x in h:
        x.remove()
    cap["all_coords_preds"] = res["all_coords_preds"].detach().clone()
    return cap


def run_reference_parametric(views, wseed, iseed, mseed=11):
    """The real reference head built from config/release/train_medium_MANO.yaml (PARAMETRIC_OUTPUT): its own
    `get_parametric_output` (pt_metro_transformer.py:139-151), `rot6d_to_aa` (utils/transform.py:448-466) and head
    epilogue (ptEmb_head.py:950-963) run on the stub layer's restated third-party pieces (pytorch3d.transforms,
    manotorch LBS on `synth.synthetic_mano()` stand-in parameters)."""
    dims = release_dims("medium_MANO")
    ref_shim.MANO_PARAMS = synth.synthetic_mano(mseed)
    try:
        head, _ = ref_shim.build_reference_head("medium_MANO", template_fn=synth.standin_template)
        sd = synth.make_state_dict(dims, wseed, "stress")
        ref_sd = dict(sd)
        for i in range(dims.n_blocks - 1):       # the reference owns the tail layers in every block; only the last runs
            for n in ("flat_verts.weight", "flat_verts.bias", "mano_linear.weight", "mano_linear.bias"):
                ref_sd.setdefault(f"transformer.pt_metro_encoder.{i}.{n}",
                                  torch.zeros_like(sd[f"transformer.pt_metro_encoder.{dims.n_blocks - 1}.{n}"]))
        missing, unexpected = head.load_state_dict(ref_sd, strict=False)
        assert not unexpected, unexpected
        assert set(ref_sd).isdisjoint(missing)
        feat, metas, ref_j = synth.make_inputs(dims, len(views), views, iseed)
        cap = {}
        last = head.transformer.pt_metro_encoder[dims.n_blocks - 1]
        orig = last.get_parametric_output

        def spy(verts_feat, verts):
            cap["tail_feats"] = verts_feat.detach().clone()
            cap["tail_xyz_in"] = verts.detach().clone()
            return orig(verts_feat, verts)
        last.get_parametric_output = spy
        with torch.no_grad():
            res = head(mlvl_feat=feat, img_metas=metas, reference_joints=ref_j, debug_metas=None)
        cap.update(all_coords_preds=res["all_coords_preds"].detach().clone(), pred_pose=res["pred_pose"].detach().clone(),
                   pred_shape=res["pred_shape"].detach().clone())
        return cap
    finally:
        ref_shim.MANO_PARAMS = None


def write_parametric_golden():
    views, wseed, iseed, mseed = [2, 3], 2, 5, 11
    cap = run_reference_parametric(views, wseed, iseed, mseed)
    meta = dict(size="medium_MANO", views=views, wseed=wseed, iseed=iseed, mseed=mseed, mode="stress")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mano_medium_b2.npz"), meta=np.array(repr(meta)),
                        **{k: v.numpy() for k, v in cap.items()})
    print("mano_medium_b2", {k: tuple(v.shape) for k, v in cap.items()}, cap["pred_pose"][0, :2], cap["pred_shape"][0, :3])


LOSS_WEIGHTS = {"HEATMAP_JOINTS_WEIGHT": 10.0, "JOINTS_LOSS_WEIGHT": 1.0, "VERTICES_LOSS_WEIGHT": 1.0,
                "JOINTS_2D_LOSS_WEIGHT": 1.0, "VERTICES_2D_LOSS_WEIGHT": 0.5}   # release weights; 2-D vertex term switched on


def run_reference_loss(B, V, seed, parametric):
    """The real `PtEmbedMultiviewStereoV2.compute_loss` (lib/models/POEM.py:363-466) called on a shell object that
    carries exactly the attributes the method reads (criteria as POEM.py:133-144 builds them for the release loss types)."""
    ref_shim.install(synth.standin_template)
    import types
    from lib.utils.builder import BACKBONE, HEAD, build_from_cfg
    sys.modules["lib.models.heads"].build_head = lambda cfg, **kw: build_from_cfg(cfg, HEAD, **kw)
    bb = sys.modules.get("lib.models.backbones")
    if bb is None:
        _import_reference_hrnet()
        bb = sys.modules.get("lib.models.backbones")
    if bb is not None and not hasattr(bb, "build_backbone"):
        bb.build_backbone = lambda cfg, **kw: build_from_cfg(cfg, BACKBONE, **kw)
    import lib.models.POEM as poem_mod
    cls = poem_mod.PtEmbedMultiviewStereoV2
    mano = synth.synthetic_mano(11)
    shell = types.SimpleNamespace(
        heatmap_joints_weights=LOSS_WEIGHTS["HEATMAP_JOINTS_WEIGHT"], joints_weight=LOSS_WEIGHTS["JOINTS_LOSS_WEIGHT"],
        vertices_weight=LOSS_WEIGHTS["VERTICES_LOSS_WEIGHT"], joints_2d_weight=LOSS_WEIGHTS["JOINTS_2D_LOSS_WEIGHT"],
        vertices_2d_weight=LOSS_WEIGHTS["VERTICES_2D_LOSS_WEIGHT"], pose_weight=0.001, shape_weight=0.0005, num_joints=21,
        mano_layer=types.SimpleNamespace(th_J_regressor=mano["J_regressor"]), parametric_output=parametric,
        transformer_center_idx=9, criterion_joints=torch.nn.MSELoss(), criterion_vertices=torch.nn.L1Loss(),
        criterion_parameters=torch.nn.MSELoss(), loss_proj_to_multicam=cls.loss_proj_to_multicam)
    preds, gt = synth.make_loss_case(B, V, seed, parametric)
    with torch.no_grad():
        loss, ld = cls.compute_loss(shell, preds, gt)
    assert torch.equal(loss, ld["loss"])
    return {k: float(v) for k, v in ld.items()}


def write_loss_golden():
    cases = {"plain_ragged": (3, [2, 1, 3], 5, False), "parametric_v4": (2, 4, 6, True)}
    out = {}
    for name, (B, V, seed, par) in cases.items():
        ld = run_reference_loss(B, V, seed, par)
        out[name] = dict(B=B, V=V, seed=seed, parametric=par, losses=ld)
        print("loss", name, ld)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "loss_cases.npz"), meta=np.array(repr(out)))


GRAD_KEYS = ["input_proj.weight", "adapt_pos3d.bias", "merge_net_feature.0.2.weight", "merge_net_feature.1.2.bias",
             "query_feat_embedding.weight", "transformer.pt_metro_encoder.0.embedding.weight",
             "transformer.pt_metro_encoder.0.encoder.attn.self.key.weight",
             "transformer.pt_metro_encoder.1.encoder.cross_attn.output.LayerNorm.weight",
             "transformer.pt_metro_encoder.1.encoder.vec_attn.query_self_attn.fc_gamma.0.weight",
             "transformer.pt_metro_encoder.2.encoder.vec_attn.query_cross_attn.fc_delta.0.weight",
             "transformer.pt_metro_encoder.2.encoder.vec_attn.query_cross_attn.w_vs.weight",
             "transformer.pt_metro_encoder.2.encoder.vec_attn.reg_branch.2.weight",
             "transformer.pt_metro_encoder.0.encoder.output.dense.weight"]


def grad_loss(coords, seed):
    """Scalar used for the gradient golden: squared distance (mm^2) of every block's prediction to a seeded target."""
    g = torch.Generator().manual_seed(seed + 77)
    target = coords.detach() + 0.005 * torch.randn(coords.shape, generator=g)
    return ((coords - target) * 1e3).pow(2).mean()


def run_reference_grads(size, views, wseed, iseed):
    """Autograd of the real reference head (eval mode: dropout off; the 32-NN indices are constants of the graph)."""
    dims = release_dims(size)
    head, _ = ref_shim.build_reference_head(size, template_fn=synth.standin_template)
    sd = synth.make_state_dict(dims, wseed, "stress")
    head.load_state_dict(sd, strict=False)
    feat, metas, ref_j = synth.make_inputs(dims, len(views), views, iseed)
    params = dict(head.named_parameters())
    for p_ in params.values():
        p_.requires_grad_(False)
    for k in GRAD_KEYS:
        params[k].requires_grad_(True)
    res = head(mlvl_feat=feat, img_metas=metas, reference_joints=ref_j, debug_metas=None)
    loss = grad_loss(res["all_coords_preds"], iseed)
    loss.backward()
    out = {"loss": np.array(loss.item())}
    for k in GRAD_KEYS:
        g = params[k].grad.detach().reshape(-1)
        out["norm:" + k] = np.array(float(g.norm()))
        out["head:" + k] = g[:: max(1, g.numel() // 64)][:64].numpy().copy()
    return out


def write_grad_golden():
    size, views, wseed, iseed = "small", [2, 1], 1, 3
    out = run_reference_grads(size, views, wseed, iseed)
    meta = dict(size=size, views=views, wseed=wseed, iseed=iseed, keys=GRAD_KEYS)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "grad_small_b2.npz"), meta=np.array(repr(meta)), **out)
    print("grad_small_b2 loss", float(out["loss"]), {k[5:]: float(out[k]) for k in out if k.startswith("norm:")})


def _import_reference_hrnet():
    ref_shim.install(synth.standin_template)
    import lib.external.metro.hrnet  # noqa: F401  (bare package; the backbone imports its config from there)
    import types
    cfgmod = types.ModuleType("lib.external.metro.hrnet.config")
    cfgmod.config = None
    cfgmod.update_config = lambda *a, **k: None
    sys.modules.setdefault("lib.external.metro.hrnet", types.ModuleType("lib.external.metro.hrnet"))
    sys.modules.setdefault("lib.external.metro.hrnet.config", cfgmod)
    import lib.models.backbones.hrnet as hr
    return hr


def run_reference_backbone(n_images, wseed, iseed):
    """The real reference `HighResolutionNet` (hrnet.py:239-420) built from its own W40 yaml, on CPU, eval mode."""
    import yaml
    hr = _import_reference_hrnet()
    with open(os.path.join(ref_shim.REF_ROOT, "config/backbone/cls_hrnet_w40_sgd_lr5e-2_wd1e-4_bs32_x100.yaml")) as f:
        cfg = yaml.safe_load(f)
    net = hr.HighResolutionNet(cfg).eval()
    sd = synth.make_backbone_state_dict(wseed)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    dead = ("incre_modules.", "downsamp_modules.", "final_layer.", "classifier.")
    assert not unexpected and all(k.startswith(dead) for k in missing), (missing[:5], unexpected[:5])
    with torch.no_grad():
        return net(synth.make_images(n_images, 256, iseed))


def run_reference_image_stage(n_images, wseed, iseed):
    """Real reference backbone (`HighResolutionNet`, W40 yaml) followed by the real `feat_decode` method of
    `PtEmbedMultiviewStereoV2` (lib/models/POEM.py:189-203) bound to real `ConvBlock`s, on CPU in eval mode."""
    import yaml
    hr = _import_reference_hrnet()
    # POEM.py imports the two builders from packages the stub layer keeps bare (their __init__ would pull every head)
    from lib.utils.builder import BACKBONE, HEAD, build_from_cfg
    sys.modules["lib.models.heads"].build_head = lambda cfg, **kw: build_from_cfg(cfg, HEAD, **kw)
    bb = sys.modules.get("lib.models.backbones")
    if bb is not None and not hasattr(bb, "build_backbone"):
        bb.build_backbone = lambda cfg, **kw: build_from_cfg(cfg, BACKBONE, **kw)
    import lib.models.POEM as poem_mod
    from lib.models.bricks.conv import ConvBlock
    with open(os.path.join(ref_shim.REF_ROOT, "config/backbone/cls_hrnet_w40_sgd_lr5e-2_wd1e-4_bs32_x100.yaml")) as f:
        cfg = yaml.safe_load(f)

    class Shell(torch.nn.Module):       # the attributes feat_decode touches, built as POEM.py:169-181 builds them
        def __init__(self):
            super().__init__()
            fs = (40, 80, 160, 320)
            self.img_backbone = hr.HighResolutionNet(cfg)
            self.feat_delayer = torch.nn.ModuleList([ConvBlock(fs[i], fs[i + 1], kernel_size=3, stride=2, relu=True, norm="bn")
                                                     for i in range(3)])
            self.feat_in = ConvBlock(fs[3], fs[2], kernel_size=1, padding=0, relu=False, norm=None)
            self.num_joints = 21
            self.uv_delayer = torch.nn.ModuleList([ConvBlock(fs[3 - i] + fs[2 - i], fs[2 - i], kernel_size=3, relu=True,
                                                             norm="bn") for i in range(3)])
            self.uv_out = ConvBlock(fs[0], 21, kernel_size=1, padding=0, relu=False, norm=None)
            self.uv_in = ConvBlock(21, fs[1], kernel_size=1, padding=0, relu=True, norm="bn")

        def uv_decode(self, feats):
            return poem_mod.PtEmbedMultiviewStereoV2.uv_decode(self, feats)
    shell = Shell().eval()
    sd = synth.make_image_stage_state_dict(wseed)
    missing, unexpected = shell.load_state_dict(sd, strict=False)
    dead = ("img_backbone.incre_modules.", "img_backbone.downsamp_modules.", "img_backbone.final_layer.",
            "img_backbone.classifier.", "uv_in.")
    assert not unexpected and all(k.startswith(dead) for k in missing), (missing[:5], unexpected[:5])
    from lib.utils.triangulation import batch_triangulate_dlt_torch
    with torch.no_grad():
        feats = shell.img_backbone(synth.make_images(n_images, 256, iseed))
        mlvl = poem_mod.PtEmbedMultiviewStereoV2.feat_decode(shell, feats, "HRNet")
        uv = poem_mod.PtEmbedMultiviewStereoV2.heatmap_stage(shell, feats, 256, 256)
        # the per-sample DLT loop of POEM.py:284-299 on the reference function, all images as the views of one sample
        intr, extr = synth.make_cameras(1, [n_images], iseed)
        K, T = intr.view(-1, 3, 3), torch.linalg.inv(extr.view(-1, 4, 4))
        ref_j = batch_triangulate_dlt_torch(uv.unsqueeze(0), K.unsqueeze(0), T.unsqueeze(0))
    return mlvl, feats, uv, ref_j


def run_reference_pa_metrics(seed):
    """The real `PAEval` (lib/metrics/pa_eval.py) fed with seeded joints / vertices; returns inputs and its measures."""
    ref_shim.install(synth.standin_template)
    from lib.metrics.pa_eval import PAEval
    g = torch.Generator().manual_seed(seed)
    gt_j = torch.tensor([0.0, 0.0, 0.6]) + 0.05 * torch.randn(5, 21, 3, generator=g)
    gt_v = torch.tensor([0.0, 0.0, 0.6]) + 0.05 * torch.randn(5, 778, 3, generator=g)
    # predictions: a similarity transform of the ground truth (rotation, scale, shift) plus noise
    ang = 0.3 * torch.randn(5, 3, generator=g)
    Rm = torch.linalg.matrix_exp(torch.stack([torch.tensor([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]]) for a in ang]))
    sc = 1.0 + 0.1 * torch.randn(5, 1, 1, generator=g)
    sh = 0.02 * torch.randn(5, 1, 3, generator=g)
    pr_j = (gt_j - gt_j.mean(1, keepdim=True)) @ Rm.transpose(1, 2) * sc + gt_j.mean(1, keepdim=True) + sh \
        + 0.004 * torch.randn(5, 21, 3, generator=g)
    pr_v = (gt_v - gt_v.mean(1, keepdim=True)) @ Rm.transpose(1, 2) * sc + gt_v.mean(1, keepdim=True) + sh \
        + 0.004 * torch.randn(5, 778, 3, generator=g)
    ev = PAEval(None, mesh_score=True)
    ev.feed(pr_j, gt_j, pr_v, gt_v)
    al_j = np.stack([PAEval.align_w_scale(gt_j[i].numpy().copy(), pr_j[i].numpy().copy())[0] for i in range(5)])
    return dict(gt_j=gt_j.numpy(), gt_v=gt_v.numpy(), pr_j=pr_j.numpy(), pr_v=pr_v.numpy(), aligned_j=al_j,
                measures=np.array([ev.get_measures()[k] for k in ("pa_mpjpe", "mpjpe", "pa_mpvpe", "mpvpe")]))


def run_reference_stage4(n_images, wseed, iseed):
    """The real reference `HighResolutionModule` x3 (= `HighResolutionNet.stage4`, hrnet.py:272-277) on CPU."""
    hr = _import_reference_hrnet()
    ch = [40, 80, 160, 320]
    mods = torch.nn.Sequential(*[hr.HighResolutionModule(4, hr.BasicBlock, [4] * 4, list(ch), list(ch), "SUM", True)
                                 for _ in range(3)]).eval()
    sd = synth.make_stage4_state_dict(wseed)
    mods.load_state_dict(sd, strict=True)
    xs = synth.make_stage4_inputs(n_images, 64, iseed)
    with torch.no_grad():
        ys = mods(list(xs))
    return ys


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    if "--only-mano" in sys.argv:
        return write_parametric_golden()
    if "--only-loss" in sys.argv:
        return write_loss_golden()
    if "--only-grad" in sys.argv:
        return write_grad_golden()
    ys = run_reference_backbone(1, 0, 1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hrnet_w40_n1.npz"),
                        meta=np.array(repr(dict(kind="hrnet_w40", n_images=1, wseed=0, iseed=1, stride=4))),
                        y0=ys[0][:, :, ::4, ::4].numpy(), y1=ys[1][:, :, ::2, ::2].numpy(), y2=ys[2].numpy(),
                        y3=ys[3].numpy())
    print("hrnet_w40_n1", [tuple(y.shape) for y in ys], [float(y.abs().mean()) for y in ys])
    mf, _, uv, rj = run_reference_image_stage(3, 0, 1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "image_stage_n3.npz"),
                        meta=np.array(repr(dict(kind="image_stage", n_images=3, wseed=0, iseed=1))), mlvl_feat=mf.numpy(),
                        pred_joints_uv=uv.numpy(), ref_joints=rj.numpy())
    print("image_stage_n3", tuple(mf.shape), float(mf.abs().mean()), tuple(uv.shape), tuple(rj.shape), rj[0, :2])
    pm = run_reference_pa_metrics(3)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics_pa.npz"), **pm)
    print("metrics_pa", pm["measures"])
    if "--only-hrnet" in sys.argv:
        return
    ys = run_reference_stage4(2, 0, 1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hrnet_stage4_n2.npz"),
                        meta=np.array(repr(dict(kind="hrnet_stage4", n_images=2, wseed=0, iseed=1, stride=4))),
                        y0=ys[0][:, :, ::4, ::4].numpy(), y1=ys[1][:, :, ::2, ::2].numpy(), y2=ys[2].numpy(),
                        y3=ys[3].numpy())
    print("hrnet_stage4_n2", [tuple(y.shape) for y in ys])
    write_parametric_golden()
    write_loss_golden()
    write_grad_golden()
    for name, (size, views, wseed, iseed, mode) in CASES.items():
        cap = run_reference(size, views, wseed, iseed, mode)
        out = {
            "all_coords_preds": cap["all_coords_preds"].numpy(),
            "pt_feats_rows": cap["pt_feats"][:, ::ROW_STRIDE].numpy(),
            "pt_xyz_rows": cap["pt_xyz"][:, ::ROW_STRIDE].numpy(),
            "q_xyz": cap["q_xyz"].numpy(),
        }
        for k in cap:
            if k.endswith(".out"):
                out[k + "_rows"] = cap[k][:, ::ROW_STRIDE].numpy()
            if k.endswith(".xyz"):
                out[k] = cap[k].numpy()
        meta = dict(size=size, views=views, wseed=wseed, iseed=iseed, mode=mode, row_stride=ROW_STRIDE)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), meta=np.array(repr(meta)), **out)
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
