"""images -> mesh on B200 (SURVEY §8d metric ii): HRNet-W40 backbone + feat_decode + POEM decoder head, reference joints
given (the heatmap / DLT stage, §8f row f2, is outside this measurement).  samples/s with images resident in HBM."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poem_v2_b200 import _native as nat  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402
from poem_v2_b200.config import release_dims  # noqa: E402
from poem_v2_b200.head import POEM_Generalized_Head  # noqa: E402
from poem_v2_b200.hrnet import ImageStage, backbone_flops_per_image  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "medium"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
V = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dims = release_dims(size)
stage = ImageStage()
sd_img = synth.make_image_stage_state_dict(0)
sd_img["feat_in.conv.weight"] *= 0.1          # unit-scale mlvl_feat for the synthetic head weights
sd_img["feat_in.conv.bias"] *= 0.1
stage.load_state_dict(sd_img)
head = POEM_Generalized_Head(dims, template_mesh=synth.standin_template())
head.load_state_dict(synth.make_state_dict(dims, 0), strict=True)
head = head.cuda().eval()
_, metas, ref_j = synth.make_inputs(dims, B, [V] * B, 1)
m = dict(metas)
m["cam_intr"], m["cam_extr"] = metas["cam_intr"].cuda(), metas["cam_extr"].cuda()
ref_j = ref_j.cuda()
imgs = [synth.make_images(B * V, 256, s).cuda() for s in range(2)]   # 2 x 201 MB, alternated (>> L2)


def step(i):
    feat = stage(imgs[i & 1])
    return head(mlvl_feat=feat, img_metas=m, reference_joints=ref_j)["all_coords_preds"]


for i in range(3):
    step(i)
torch.cuda.synchronize()
K = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K):
    out = step(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
t_img = []
for i in range(K):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    stage(imgs[i & 1])
    b.record()
    torch.cuda.synchronize()
    t_img.append(a.elapsed_time(b))
lib = nat.load()
lib.poem_profile_enable(1)
stage(imgs[0])
torch.cuda.synchronize()
prof = nat.profile_summary()
lib.poem_profile_enable(0)
ms_img = sorted(t_img)[len(t_img) // 2]
gflop_img = sum(backbone_flops_per_image().values()) / 1e9
print(json.dumps({
    "workload": f"images -> mesh, POEM-{size}, {B} samples x {V} views (3x256x256), reference joints given",
    "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "images_per_s": B * V / ms * 1e3,
    "image_stage_ms": ms_img, "decoder_ms": ms - ms_img, "image_stage_images_per_s": B * V / ms_img * 1e3,
    "backbone_nominal_tflops": gflop_img * B * V / ms_img,
    "finite": bool(torch.isfinite(out).all()),
    "image_stage_kernels_ms": {k: [round(v["ms"], 3), v["n"]] for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]}}))
