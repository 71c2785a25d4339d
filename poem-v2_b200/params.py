"""Product-side constants of the decoder path: frozen assets, the state-dict keys / shapes the path reads, and the
MANO topology the parametric tail assumes.  (Kept apart from `synth.py`, which only generates seeded test inputs and
weights: the checker — `oracle/` — must not share a module with the product.)"""
import os

import numpy as np
import torch

from .config import HeadDims

_ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def load_assets():
    """Frozen constants of the path: BPS offsets (4096,3) f32, anchor xyz (32,3) f32, anchor idx (32,) i64.

    Same bytes as the reference's `assets/{bps,anchor,anchor_idx}.npy` (data, not code; loaded by the
    reference at `ptEmb_head.py:790-809` and `point_transformers.py:10-32`)."""
    bps = np.load(os.path.join(_ASSET_DIR, "bps.npy")).reshape(-1, 3).astype(np.float32)
    anchor = np.load(os.path.join(_ASSET_DIR, "anchor.npy")).reshape(-1, 3).astype(np.float32)
    anchor_idx = np.load(os.path.join(_ASSET_DIR, "anchor_idx.npy")).reshape(-1).astype(np.int64)
    return torch.from_numpy(bps), torch.from_numpy(anchor), torch.from_numpy(anchor_idx)


# MANO kinematic tree (parent of each of the 16 joints) and the fingertip vertices manotorch appends for a right hand
MANO_PARENTS = (-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14)
MANO_TIP_VERTS = (745, 317, 444, 556, 673)
# manotorch's 16 joints + 5 tips -> the 21-joint hand order
MANO_JOINT_ORDER = (0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20)



def live_param_shapes(dims: HeadDims):
    """name -> shape of every parameter the path reads (reference state-dict names, relative to the head)."""
    D, C = dims.embed_dims, dims.in_channels
    H = D // 2
    s = {
        "input_proj.weight": (D, C, 1, 1), "input_proj.bias": (D,),
        "adapt_pos3d.weight": (D, 3 * dims.pos_feats, 1, 1), "adapt_pos3d.bias": (D,),
        "merge_net_feature.0.0.weight": (D, D), "merge_net_feature.0.0.bias": (D,),
        "merge_net_feature.0.2.weight": (H, D), "merge_net_feature.0.2.bias": (H,),
        "merge_net_feature.1.0.weight": (H, H), "merge_net_feature.1.0.bias": (H,),
        "merge_net_feature.1.2.weight": (D, H), "merge_net_feature.1.2.bias": (D,),
        "query_feat_embedding.weight": (dims.n_query, D),
    }
    for i in range(dims.n_blocks):
        p = f"transformer.pt_metro_encoder.{i}."
        s[p + "embedding.weight"] = (D, D)
        s[p + "embedding.bias"] = (D,)
        for a in ("attn", "cross_attn"):
            for n in ("self.query", "self.key", "self.value", "output.dense"):
                s[p + f"encoder.{a}.{n}.weight"] = (D, D)
                s[p + f"encoder.{a}.{n}.bias"] = (D,)
            s[p + f"encoder.{a}.output.LayerNorm.weight"] = (D,)
            s[p + f"encoder.{a}.output.LayerNorm.bias"] = (D,)
        for a in ("query_self_attn", "query_cross_attn"):
            q = p + f"encoder.vec_attn.{a}."
            for n in ("fc1", "fc2", "fc_delta.2", "fc_gamma.0", "fc_gamma.2"):
                s[q + n + ".weight"] = (D, D)
                s[q + n + ".bias"] = (D,)
            s[q + "fc_delta.0.weight"] = (D, 3)
            s[q + "fc_delta.0.bias"] = (D,)
            for n in ("w_qs", "w_ks", "w_vs"):
                s[q + n + ".weight"] = (D, D)
        s[p + "encoder.vec_attn.reg_branch.0.weight"] = (D, D)
        s[p + "encoder.vec_attn.reg_branch.0.bias"] = (D,)
        s[p + "encoder.vec_attn.reg_branch.2.weight"] = (3, D)
        s[p + "encoder.vec_attn.reg_branch.2.bias"] = (3,)
        s[p + "encoder.intermediate.dense.weight"] = (4 * D, D)
        s[p + "encoder.intermediate.dense.bias"] = (4 * D,)
        s[p + "encoder.output.dense.weight"] = (D, 4 * D)
        s[p + "encoder.output.dense.bias"] = (D,)
        s[p + "encoder.output.LayerNorm.weight"] = (D,)
        s[p + "encoder.output.LayerNorm.bias"] = (D,)
        if dims.parametric:
            s[p + "flat_verts.weight"] = (1, dims.n_query)
            s[p + "flat_verts.bias"] = (1,)
            s[p + "mano_linear.weight"] = (106, D)
            s[p + "mano_linear.bias"] = (106,)
    return s
