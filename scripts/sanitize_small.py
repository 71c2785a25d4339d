"""Tiny decoder forwards for compute-sanitizer (memcheck): POEM-small head (ragged views), the medium_MANO head, and view counts that
do not divide 128 (generic tiles of the fused sampler / merge kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poem_v2_b200 import synth
from poem_v2_b200.config import release_dims
from poem_v2_b200.head import POEM_Generalized_Head

for size, mano, views in (("small", None, [2, 1]), ("medium_MANO", synth.synthetic_mano(11), [2, 1]),
                          ("small", None, [3, 10, 7]), ("medium", None, [5, 9])):   # last two: generic (ragged) tiles of the fused sampler
    dims = release_dims(size)
    head = POEM_Generalized_Head(dims, template_mesh=None if mano else synth.standin_template(), mano_params=mano)
    head.load_state_dict(synth.make_state_dict(dims, 0), strict=True)
    head = head.cuda().eval()
    feat, metas, ref_j = synth.make_inputs(dims, len(views), views, 1)
    m = dict(metas)
    m["cam_intr"], m["cam_extr"] = metas["cam_intr"].cuda(), metas["cam_extr"].cuda()
    out = head(mlvl_feat=feat.cuda(), img_metas=m, reference_joints=ref_j.cuda())
    torch.cuda.synchronize()
    print(size, views, {k: tuple(v.shape) for k, v in out.items()}, "finite", bool(torch.isfinite(out["all_coords_preds"]).all()))
