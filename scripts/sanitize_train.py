"""Tiny training steps for compute-sanitizer (memcheck): POEM-small head with ragged views, with dropout, and the parametric
(MANO tail) variant — forward with saved activations, 3-D loss, backward, per-tensor clip, Adam."""
import os
import sys
from dataclasses import replace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from poem_v2_b200 import synth  # noqa: E402
from poem_v2_b200.config import release_dims  # noqa: E402
from poem_v2_b200.pack import mano_zero_pose_template  # noqa: E402
from poem_v2_b200.train import HeadTrainer, TrainStep  # noqa: E402

for name, dims, views, p_drop in (("small", release_dims("small"), [2, 1], 0.0), ("small+dropout", release_dims("small"), [1, 3], 0.1),
                                  ("small parametric", replace(release_dims("small"), parametric=True), [2], 0.1)):
    mano = synth.synthetic_mano(11) if dims.parametric else None
    tmpl = mano_zero_pose_template(mano, dims.center_idx) if dims.parametric else synth.standin_template()
    tr = HeadTrainer(dims, synth.make_state_dict(dims, 0), tmpl, dropout=p_drop, mano=mano)
    step = TrainStep(tr, lr=1e-4, max_norm=1.0)
    feat, metas, ref_j = synth.make_inputs(dims, len(views), views, 1)
    m = dict(metas)
    m["cam_intr"], m["cam_extr"] = metas["cam_intr"].cuda(), metas["cam_extr"].cuda()
    gt_v = ref_j[:, 9:10] + 0.05 * torch.randn(len(views), 778, 3)
    losses = [float(step(feat.cuda(), m, ref_j.cuda(), ref_j, gt_v).item()) for _ in range(2)]
    torch.cuda.synchronize()
    print(name, views, "loss", losses, "finite", bool(torch.isfinite(tr.p_flat).all()))
