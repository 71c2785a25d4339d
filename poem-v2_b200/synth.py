"""Seeded synthetic inputs, weights and constants for the decoder path.

Everything here is a pure function of integer seeds and `HeadDims`, generated with a CPU
`torch.Generator`, so the container that writes `tests/golden/` and the GPU box regenerate
bit-identical tensors.
"""
import math
import os

import numpy as np
import torch

from .config import HeadDims, N_QUERY
from .params import (MANO_JOINT_ORDER, MANO_PARENTS, MANO_TIP_VERTS, live_param_shapes,  # noqa: F401  (re-exported)
                     load_assets)



def standin_template(seed: int = 7) -> torch.Tensor:
    """(799,3) stand-in for the MANO zero-pose template (joints ‖ verts, centred on joint 9).

    manotorch + the licensed MANO pickles are absent; SURVEY §8d prescribes a seeded 0.05·N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    t = 0.05 * torch.randn(N_QUERY, 3, generator=g)
    return t - t[9:10]


def synthetic_mano(seed: int = 11):
    """Seeded stand-in for the MANO model parameters (the licensed `assets/mano_v1_2` pickles are absent offline):
    same tensor shapes and roles as manotorch's `th_*` buffers, hand-sized magnitudes (metres).  Used by BOTH the
    reference run that writes the parametric golden and by our head, so the parametric tail
    is exercised end to end; parity with the real MANO data is unpinned."""
    g = torch.Generator().manual_seed(seed)
    v = 0.05 * torch.randn(778, 3, generator=g)
    return {
        "v_template": v.contiguous(),
        "shapedirs": (0.004 * torch.randn(778, 3, 10, generator=g)).contiguous(),
        "posedirs": (0.002 * torch.randn(778, 3, 135, generator=g)).contiguous(),
        "J_regressor": torch.softmax(3.0 * torch.randn(16, 778, generator=g), dim=1).contiguous(),
        "weights": torch.softmax(4.0 * torch.randn(778, 16, generator=g), dim=1).contiguous(),
    }


def _view_list(B, V):
    return [int(V)] * B if isinstance(V, (int, np.integer)) else [int(v) for v in V]


def make_cameras(B: int, V, seed: int = 1):
    """cam_intr (ΣV,3,3), cam_extr (ΣV,4,4) cam->master; view 0 of each sample is the master (identity).
    V is an int or a per-sample list (ragged batches, reference `collation_random_n_views`)."""
    g = torch.Generator().manual_seed(seed + 101)
    K = torch.tensor([[900.0, 0.0, 128.0], [0.0, 900.0, 128.0], [0.0, 0.0, 1.0]])
    c0 = torch.tensor([0.0, 0.0, 0.6])
    views = _view_list(B, V)
    mats = []
    for b in range(B):
        V = views[b]
        for v in range(V):
            yaw = v * 2.0 * math.pi / V * 0.25
            pitch = 0.0 if v == 0 else (torch.rand(1, generator=g).item() * 2 - 1) * math.radians(5.0)
            cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
            Ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float32)
            Rx = torch.tensor([[1, 0, 0], [0, cp, -sp], [0, sp, cp]], dtype=torch.float32)
            R = Ry @ Rx
            # rotate the camera about the hand centre c0: p_master = R (p_cam - c0) + c0
            T = torch.eye(4)
            T[:3, :3] = R
            T[:3, 3] = c0 - R @ c0
            mats.append(T)
    extr = torch.stack(mats)
    intr = K[None].repeat(extr.shape[0], 1, 1)
    return intr.contiguous(), extr.contiguous()


def make_inputs(dims: HeadDims, B: int, V, seed: int = 1):
    """Decoder-only inputs: mlvl_feat ~ N(0,1) (ΣV,C,16,16); reference_joints = c0 + 0.03 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    views = _view_list(B, V)
    feat = torch.randn(sum(views), dims.in_channels, dims.feat_hw, dims.feat_hw, generator=g)
    c0 = torch.tensor([0.0, 0.0, 0.6])
    ref_joints = c0 + 0.03 * torch.randn(B, 21, 3, generator=g)
    intr, extr = make_cameras(B, V, seed)
    img_metas = {
        "inp_img_shape": (256, 256),
        "cam_intr": intr,
        "cam_extr": extr,
        "master_id": [0] * B,
        "cam_view_num": np.array(views),
    }
    return feat, img_metas, ref_joints


# per-layer gains that keep every stage O(1) (logit std ~1-2, per-block xyz update ~0.03 radius units)
_STRESS_GAIN = {
    "merge_net_feature.0.2.weight": 0.35,
    "fc_delta.0.weight": 2.0,
    "reg_branch.2.weight": 0.08,
    "vec_attn.query_self_attn.fc2.weight": 0.5,
    "vec_attn.query_cross_attn.fc2.weight": 0.5,
}


def make_state_dict(dims: HeadDims, seed: int = 0, mode: str = "stress"):
    """Seeded live weights under the reference's key names.

    mode "stress": every matrix ~ N(0, 1/fan_in) (biases 0.1·N(0,1), LayerNorm γ = 1 + 0.1·N) so that
    attention logits, softmaxes and residual branches are all O(1) and non-degenerate;
    the coordinate regression layer is scaled so per-block updates are ~0.05 normalised units.
    mode "init": BERT-style N(0, 0.02) matrices, zero biases.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in live_param_shapes(dims).items():
        if "LayerNorm.weight" in name:
            w = 1.0 + (0.1 * torch.randn(shape, generator=g) if mode == "stress" else 0.0)
            w = w if torch.is_tensor(w) else torch.ones(shape)
        elif name.endswith(".bias"):
            w = 0.1 * torch.randn(shape, generator=g) if mode == "stress" else torch.zeros(shape)
        elif name == "query_feat_embedding.weight":
            w = torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            std = (1.0 / math.sqrt(fan_in)) if mode == "stress" else 0.02
            w = std * torch.randn(shape, generator=g)
            if mode == "stress":
                for suffix, gain in _STRESS_GAIN.items():
                    if name.endswith(suffix):
                        w = w * gain
        sd[name] = w.float().contiguous()
    return sd


def make_stage4_state_dict(seed: int = 0, n_modules: int = 3):
    """Seeded weights for HRNet stage 4 under the reference key names: He-style conv weights, BatchNorm statistics
    and affine terms perturbed around identity so that folding is exercised (gamma, var != 1; beta, mean != 0)."""
    from .hrnet import stage4_param_shapes
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in stage4_param_shapes(n_modules).items():
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(100)
        elif name.endswith("running_var"):
            sd[name] = 0.5 + torch.rand(shape, generator=g)
        elif name.endswith("running_mean") or name.endswith(".bias"):
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:                       # BN gamma
            sd[name] = 0.6 + 0.2 * torch.rand(shape, generator=g)
        else:                                       # conv weight
            fan_in = shape[1] * shape[2] * shape[3]
            sd[name] = torch.randn(shape, generator=g) * math.sqrt(1.0 / fan_in)
    return sd


def make_backbone_state_dict(seed: int = 0):
    """Seeded weights for the whole HRNet-W40 backbone under the reference key names (same recipe as stage 4)."""
    from .hrnet import backbone_param_shapes
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in backbone_param_shapes().items():
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(100)
        elif name.endswith("running_var"):
            sd[name] = 0.5 + torch.rand(shape, generator=g)
        elif name.endswith("running_mean") or name.endswith(".bias"):
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            sd[name] = 0.6 + 0.2 * torch.rand(shape, generator=g)
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            sd[name] = torch.randn(shape, generator=g) * math.sqrt(1.0 / fan_in)
    return sd


def make_image_stage_state_dict(seed: int = 0):
    """Seeded weights for backbone + feat_decode under the full-model key names (`img_backbone.*`, `feat_delayer.*`,
    `feat_in.*`)."""
    from .hrnet import image_stage_param_shapes
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in image_stage_param_shapes().items():
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(100)
        elif name.endswith("running_var"):
            sd[name] = 0.5 + torch.rand(shape, generator=g)
        elif name.endswith("running_mean") or name.endswith(".bias"):
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            sd[name] = 0.6 + 0.2 * torch.rand(shape, generator=g)
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            sd[name] = torch.randn(shape, generator=g) * math.sqrt(1.0 / fan_in)
    return sd


def make_images(n_images: int, res: int = 256, seed: int = 1):
    """ImageNet-normalised-looking images: smooth low-frequency content plus pixel noise, roughly unit variance."""
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(n_images, 3, res // 16, res // 16, generator=g)
    img = torch.nn.functional.interpolate(low, size=(res, res), mode="bilinear", align_corners=False)
    return img + 0.3 * torch.randn(n_images, 3, res, res, generator=g)


def make_stage4_inputs(n_images: int, base_res: int = 64, seed: int = 1):
    g = torch.Generator().manual_seed(seed)
    return [torch.relu(torch.randn(n_images, c, base_res >> b, base_res >> b, generator=g))
            for b, c in enumerate((40, 80, 160, 320))]


def make_model_state_dict(dims: HeadDims, seed: int = 0):
    """Full-model checkpoint keys (reference `PtEmbedMultiviewStereoV2.state_dict()` minus dead weights): image half +
    `ptEmb_head.*`.  `feat_in` is scaled so that `mlvl_feat` has unit scale, the range the synthetic head weights are
    made for; `uv_out` is scaled so that the heatmaps are not saturated."""
    sd = make_image_stage_state_dict(seed)
    sd["feat_in.conv.weight"] = sd["feat_in.conv.weight"] * 0.1
    sd["feat_in.conv.bias"] = sd["feat_in.conv.bias"] * 0.1
    sd["uv_out.conv.weight"] = sd["uv_out.conv.weight"] * 0.2
    sd.update({"ptEmb_head." + k: v for k, v in make_state_dict(dims, seed).items()})
    return sd


def make_batch(B: int, V, seed: int = 1):
    """A reference-style evaluation batch (POEM.py:251-315): images, cameras, view counts, master ids (+ the GT joints the
    reference falls back to when every sample is single-view)."""
    views = _view_list(B, V)
    intr, extr = make_cameras(B, V, seed)
    g = torch.Generator().manual_seed(seed + 7)
    return {"image": make_images(sum(views), 256, seed), "cam_view_num": np.array(views), "target_cam_intr": intr,
            "target_cam_extr": extr, "master_id": [0] * B,
            "master_joints_3d": torch.tensor([0.0, 0.0, 0.6]) + 0.03 * torch.randn(B, 21, 3, generator=g)}


def make_loss_case(B: int, V, seed: int = 1, parametric: bool = False, n_blocks: int = 3):
    """Seeded (preds, gt) pair for the training loss (reference POEM.py:363-466): predictions a few mm / px away from
    the ground truth, ragged view counts allowed."""
    views = _view_list(B, V)
    nv = sum(views)
    intr, extr = make_cameras(B, V, seed)
    g = torch.Generator().manual_seed(seed + 31)
    c0 = torch.tensor([0.0, 0.0, 0.6])
    jg = c0 + 0.03 * torch.randn(B, 21, 3, generator=g)
    vg = c0 + 0.03 * torch.randn(B, 778, 3, generator=g)
    gt = {"image": torch.zeros(nv, 3, 256, 256), "cam_view_num": np.array(views), "target_cam_intr": intr,
          "target_cam_extr": extr, "master_joints_3d": jg, "master_verts_3d": vg,
          "target_joints_2d": 128.0 + 60.0 * torch.randn(nv, 21, 2, generator=g)}
    coords = torch.cat([jg, vg], dim=1)[None] + 0.004 * torch.randn(n_blocks, B, 799, 3, generator=g)
    preds = {"all_coords_preds": coords, "pred_joints_uv": gt["target_joints_2d"] + 3.0 * torch.randn(nv, 21, 2, generator=g)}
    if parametric:
        gt["mano_pose"] = 0.3 * torch.randn(nv, 16, 3, generator=g)
        gt["mano_shape"] = torch.randn(nv, 10, generator=g)
        preds["pred_pose"] = 0.3 * torch.randn(B, 16, 3, generator=g)
        preds["pred_shape"] = torch.randn(B, 10, generator=g)
    return preds, gt
