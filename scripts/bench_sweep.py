"""Device-resident decoder timing over BASELINE.json's other configurations (reported context, not the bench line):
POEM-large view sweep 2..10 at 8 samples per GPU (configs[4]) and POEM-medium_MANO, 8 views, batch 32 (configs[3], per-GPU
evaluation forward incl. the MANO tail).  Eager launches, CUDA events, inputs rotated over 4 sets; prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import analytic_roofline, measured_peaks  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402
from poem_v2_b200.config import release_dims  # noqa: E402
from poem_v2_b200.head import POEM_Generalized_Head  # noqa: E402


def run(size, V, B, steps=20, warmup=5):
    dims = release_dims(size)
    mano = synth.synthetic_mano(11) if dims.parametric else None
    head = POEM_Generalized_Head(dims, template_mesh=None if dims.parametric else synth.standin_template(), mano_params=mano)
    head.load_state_dict(synth.make_state_dict(dims, 0), strict=True)
    head = head.cuda().eval()
    sets = []
    for r in range(4):
        feat, metas, ref_j = synth.make_inputs(dims, B, V, seed=1 + r)
        m = dict(metas)
        m["cam_intr"], m["cam_extr"] = metas["cam_intr"].cuda(), metas["cam_extr"].cuda()
        sets.append((feat.cuda(), m, ref_j.cuda()))
    for i in range(warmup):
        f, m, r = sets[i % 4]
        out = head(mlvl_feat=f, img_metas=m, reference_joints=r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        f, m, r = sets[i % 4]
        out = head(mlvl_feat=f, img_metas=m, reference_joints=r)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    peaks = measured_peaks()
    flops, byts = analytic_roofline(dims.embed_dims, V)
    t_roof = max(flops / (peaks["bf16_tflops"] * 1e12), byts / (peaks["hbm_gbs"] * 1e9))
    sps = B / ms * 1e3
    res = {"size": size, "views": V, "batch": B, "ms_per_step": round(ms, 3), "samples_per_s": round(sps, 1),
           "frac_of_path_roofline": round(sps * t_roof, 4), "finite": bool(torch.isfinite(out["all_coords_preds"]).all()),
           "outputs": sorted(out)}
    del head, sets
    torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    rows = [run("medium_MANO", 8, 32), run("medium", 8, 32)]
    for V in (2, 4, 6, 8, 10):
        rows.append(run("large", V, 8))
    print(json.dumps({"launch": "eager, device-resident inputs, CUDA events, 20 steps after 5 warm-ups", "rows": rows}))
