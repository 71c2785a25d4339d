"""Dimensions of the point-embedded transformer decoder path.

Mirrors what the reference head reads from ``cfg.MODEL.HEAD`` (reference
`lib/models/heads/ptEmb_head.py:57-76,686-695`, `lib/models/layers/ptEmb_transformer.py:312-324`)
and the per-release widths of `config/release/train_{small,medium,medium_MANO,large,huge}.yaml`.
"""
from dataclasses import dataclass, asdict

N_JOINTS = 21
N_VERTS = 778
N_QUERY = N_JOINTS + N_VERTS  # 799


@dataclass(frozen=True)
class HeadDims:
    embed_dims: int = 256        # D   (EMBED_DIMS == POINTS_FEAT_DIM == TRANSFORMER.INPUT_FEAT_DIM)
    in_channels: int = 160       # C   (IN_CHANNELS)
    n_sample: int = 4096         # P   (N_SAMPLE == BPS_FEAT_DIM)
    n_query: int = N_QUERY       # Q
    n_blocks: int = 3            # NB  (TRANSFORMER.N_BLOCKS == NUM_PREDS)
    n_heads: int = 4             # h   (NUM_ATTENTION_HEADS)
    n_neighbor: int = 32         # K   (N_NEIGHBOR == N_NEIGHBOR_QUERY)
    radius: float = 0.1          # r   (RADIUS_SAMPLE), metres
    center_idx: int = 9          # TRANSFORMER_CENTER_IDX
    pos_feats: int = 128         # POSITIONAL_ENCODING.NUM_FEATS (== D/2 in every release config)
    pos_normalize: bool = True
    feat_hw: int = 16            # HRNet stride-16 feature map (16x16 for 256x256 input)
    parametric: bool = False     # TRANSFORMER.PARAMETRIC_OUTPUT (medium_MANO)
    dropout: float = 0.0         # TRANSFORMER.DROPOUT (training mode only: hidden + attention-probability dropout of the BERT layers)

    def as_dict(self):
        return asdict(self)


_WIDTH = {"small": 128, "medium": 256, "medium_MANO": 256, "large": 512, "huge": 1024}


def release_dims(size: str) -> HeadDims:
    d = _WIDTH[size]
    return HeadDims(embed_dims=d, pos_feats=d // 2, parametric=(size == "medium_MANO"))


def dims_from_cfg(cfg) -> HeadDims:
    """Build HeadDims from a reference-style config node (yacs CN or any mapping with attribute access)."""
    def g(node, key, default=None):
        if hasattr(node, "get"):
            v = node.get(key, default)
        else:
            v = getattr(node, key, default)
        return v
    tr = g(cfg, "TRANSFORMER")
    pe = g(cfg, "POSITIONAL_ENCODING")
    d = int(g(cfg, "EMBED_DIMS"))
    assert int(g(cfg, "POINTS_FEAT_DIM")) == d and int(g(tr, "INPUT_FEAT_DIM")) == d, \
        "self.pt_feat_dim should be equal to feat_dim"
    assert g(cfg, "CAM_FEAT_MERGE", "attn") == "attn"
    assert g(cfg, "QUERY_TYPE", "POEM") == "KPT"
    nn_ = int(g(tr, "N_NEIGHBOR"))
    assert nn_ == int(g(tr, "N_NEIGHBOR_QUERY")) == 32, "kernels are specialised for 32 neighbours"
    return HeadDims(embed_dims=d, in_channels=int(g(cfg, "IN_CHANNELS")), n_sample=int(g(cfg, "N_SAMPLE")),
                    n_query=int(g(cfg, "NUM_QUERY")), n_blocks=int(g(tr, "N_BLOCKS")),
                    n_heads=int(g(tr, "NUM_ATTENTION_HEADS")), n_neighbor=nn_,
                    radius=float(g(cfg, "RADIUS_SAMPLE")), center_idx=int(g(tr, "TRANSFORMER_CENTER_IDX", 9)),
                    pos_feats=int(g(pe, "NUM_FEATS")), pos_normalize=bool(g(pe, "NORMALIZE")),
                    parametric=bool(g(tr, "PARAMETRIC_OUTPUT", False)), dropout=float(g(tr, "DROPOUT", 0.0) or 0.0))
