"""poem-v2_b200 — B200-native point-embedded transformer decoder (POEM-v2 hot path).

Host side mirrors the reference plug-in boundary (`POEM_Generalized_Head` / `PtEmbedTRv4`,
reference `lib/models/heads/ptEmb_head.py:683`, `lib/models/layers/ptEmb_transformer.py:303`);
all arithmetic runs in hand-written sm_100a kernels behind the C-ABI in `include/poem_b200.h`.
"""
from .config import HeadDims, release_dims, dims_from_cfg  # noqa: F401
