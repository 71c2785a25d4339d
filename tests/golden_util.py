"""Helpers shared by the parity tests: load a golden case and regenerate its seeded inputs."""
import ast
import os

import numpy as np
import torch

from poem_v2_b200 import synth
from poem_v2_b200.config import release_dims

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith(("hrnet", "image_stage", "metrics", "mano", "loss", "grad")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = ast.literal_eval(str(z["meta"]))
    gold = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    dims = release_dims(meta["size"])
    sd = synth.make_state_dict(dims, meta["wseed"], meta["mode"])
    feat, metas, ref_j = synth.make_inputs(dims, len(meta["views"]), meta["views"], meta["iseed"])
    return meta, dims, sd, feat, metas, ref_j, gold
