"""Inference mirror of the reference model `PtEmbedMultiviewStereoV2` (lib/models/POEM.py:31-333): images ->
HRNet-W40 -> `feat_decode` -> `mlvl_feat`; `heatmap_stage` -> 2-D joints -> per-sample DLT -> `reference_joints`;
`POEM_Generalized_Head` -> mesh.  Same class name, batch-dict interface, output keys and checkpoint key names
(`img_backbone.*`, `feat_delayer.*`, `feat_in.*`, `uv_delayer.*`, `uv_out.*`, `ptEmb_head.*`), so it can be registered in
the reference's MODEL registry (`register_model_into`).  Evaluation only: losses, metrics and the training branch of
`_forward_impl` (POEM.py:269-278, 363-466) are not part of this path and raise.
All arithmetic happens in libpoem_b200.so; there is no PyTorch/CPU fallback.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _native as nat
from .config import HeadDims
from .head import POEM_Generalized_Head
from .hrnet import ImageStage

_IMAGE_PREFIXES = ("img_backbone.", "feat_delayer.", "feat_in.", "uv_delayer.", "uv_out.", "uv_in.")


class PtEmbedMultiviewStereoV2(nn.Module):
    def __init__(self, cfg, template_mesh=None, mano_params=None, **kwargs):
        super().__init__()
        self.name = type(self).__name__
        head_cfg = cfg if isinstance(cfg, HeadDims) else cfg.HEAD
        if not isinstance(cfg, HeadDims):
            bb = cfg.BACKBONE.TYPE
            if bb != "HRNet":
                raise NotImplementedError(f"backbone {bb}: only the HRNet-W40 release configuration is built")
        self.ptEmb_head = POEM_Generalized_Head(head_cfg, template_mesh=template_mesh, mano_params=mano_params)
        self.image_stage = ImageStage()
        self.num_joints = 21
        # POEM.py:40,328: the root joint of the *_rel outputs is DATA_PRESET.CENTER_IDX (0 in every release config) —
        # NOT the transformer's TRANSFORMER_CENTER_IDX (9), which only centres the BPS / MANO layer (POEM.py:409)
        preset = None if isinstance(cfg, HeadDims) else getattr(cfg, "DATA_PRESET", None)
        if preset is None and kwargs.get("data_preset") is not None:
            preset = kwargs["data_preset"]
        self.center_idx = int(getattr(preset, "CENTER_IDX", 0)) if preset is not None else 0
        self.num_preds = self.ptEmb_head.num_preds

    # ---- checkpoint interface: the reference's flat key space (net_utils.py:200-231) ----
    def state_dict(self, *a, prefix="", **k):
        sd = {prefix + n: v for n, v in self.image_stage.state_dict().items()}
        sd.update({prefix + "ptEmb_head." + n: v for n, v in self.ptEmb_head.state_dict().items()})
        return sd

    def load_state_dict(self, sd, strict=True):
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
        self.image_stage.load_state_dict({k: v for k, v in sd.items() if k.startswith(_IMAGE_PREFIXES)}, strict=strict)
        self.ptEmb_head.load_state_dict({k[len("ptEmb_head."):]: v for k, v in sd.items() if k.startswith("ptEmb_head.")},
                                        strict=strict)
        other = [k for k in sd if not k.startswith(_IMAGE_PREFIXES + ("ptEmb_head.", "mano_layer."))]
        if strict and other:
            raise RuntimeError(f"PtEmbedMultiviewStereoV2: unexpected keys {other[:4]}")

    # ---- reference method names ----
    def extract_img_feat(self, img, backbone="HRNet"):
        return self.image_stage(img, return_maps=True)["img_feats"]

    @torch.no_grad()
    def _forward_impl(self, batch, **kwargs):
        mode = kwargs.get("mode", "test")
        if mode == "train":
            raise NotImplementedError("training through this shell is not built: the image half (HRNet kernels) has no "
                                      "backward.  Train with the head inside the reference's model — "
                                      "POEM_Generalized_Head.train() routes autograd through the hand-written backward "
                                      "(poem_v2_b200/train.py, INTEGRATION.md)")
        img = batch["image"]
        if not img.is_cuda:
            raise nat.PoemError("batch['image'] must be a CUDA tensor: there is no CPU implementation")
        dev = img.device
        img = img.view(-1, img.shape[-3], img.shape[-2], img.shape[-1])
        views = [int(v) for v in batch["cam_view_num"]]
        batch_size = len(views)
        H, W = int(img.shape[-2]), int(img.shape[-1])
        BN = img.shape[0]
        assert BN == sum(views)
        res = self.image_stage(img, return_uv=True)
        mlvl_feat, uv = res["mlvl_feat"], res["pred_joints_uv"]
        intr = batch["target_cam_intr"].reshape(-1, 3, 3).to(dev)
        extr = batch["target_cam_extr"].reshape(-1, 4, 4).to(dev)
        if BN == batch_size:       # every sample single-view: the reference takes the given joints (POEM.py:279-280)
            ref_joints = batch["master_joints_3d"].reshape(-1, 21, 3).to(dev)
        else:
            ref_joints = ImageStage.triangulate(uv, intr, extr, views)
        img_metas = {"inp_img_shape": (H, W), "cam_intr": intr, "cam_extr": extr, "master_id": batch["master_id"],
                     "cam_view_num": np.asarray(views)}
        preds = self.ptEmb_head(mlvl_feat=mlvl_feat, img_metas=img_metas, reference_joints=ref_joints)
        pj = preds["all_coords_preds"][-1, :, :self.num_joints, :]
        pv = preds["all_coords_preds"][-1, :, self.num_joints:, :]
        centre = pj[:, self.center_idx, :].unsqueeze(1)
        preds.update(pred_joints_3d=pj, pred_verts_3d=pv, pred_joints_3d_rel=pj - centre, pred_verts_3d_rel=pv - centre,
                     pred_joints_uv=uv, pred_ref_joints_3d=ref_joints)
        return preds

    def forward(self, inputs, step_idx=0, mode="test", **kwargs):
        if mode == "train":
            raise NotImplementedError("training_step of the whole model is not built (no backward for the image half); the "
                                      "decoder head trains inside the reference's model, see INTEGRATION.md (SURVEY §8f row f3)")
        return self._forward_impl(inputs, mode=mode, **kwargs)


def register_model_into(model_registry):
    """Re-register under the reference's name (`MODEL.register_module(force=True)`, lib/utils/builder.py:237-239)."""
    model_registry.register_module(name="PtEmbedMultiviewStereoV2", force=True, module=PtEmbedMultiviewStereoV2)
