"""images -> mesh on B200 (SURVEY §8d metric ii): the whole evaluation forward of the reference model
(`PtEmbedMultiviewStereoV2._forward_impl(mode="test")`, POEM.py:251-333): HRNet-W40 + feat_decode + heatmap stage + DLT
+ decoder head.  samples/s with the images resident in HBM (two image sets alternated, each >> L2)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poem_v2_b200 import _native as nat  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402
from poem_v2_b200.config import release_dims  # noqa: E402
from poem_v2_b200.hrnet import backbone_flops_per_image  # noqa: E402
from poem_v2_b200.model import PtEmbedMultiviewStereoV2  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "medium"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
V = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dims = release_dims(size)
model = PtEmbedMultiviewStereoV2(dims, template_mesh=synth.standin_template())
model.load_state_dict(synth.make_model_state_dict(dims, 0), strict=True)
model = model.cuda().eval()
batches = []
for s in range(2):
    b = synth.make_batch(B, V, s + 1)
    batches.append({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()})


def step(i):
    return model(batches[i & 1], mode="test")["all_coords_preds"]


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
    model.image_stage(batches[i & 1]["image"], return_uv=True)
    b.record()
    torch.cuda.synchronize()
    t_img.append(a.elapsed_time(b))
lib = nat.load()
lib.poem_profile_enable(1)
step(0)
torch.cuda.synchronize()
prof = nat.profile_summary()
lib.poem_profile_enable(0)
from poem_v2_b200.graph import graph_model  # noqa: E402
g = graph_model(model, batches[0])
for i in range(3):
    g(batches[i & 1])
torch.cuda.synchronize()
e0.record()
for i in range(K):
    g(batches[i & 1])
e1.record()
torch.cuda.synchronize()
ms_graph = e0.elapsed_time(e1) / K
ms_img = sorted(t_img)[len(t_img) // 2]
gflop_img = sum(backbone_flops_per_image().values()) / 1e9
groups = {}
for k, v in prof.items():
    name = k.split(":")[0] if "conv" not in k else ("conv3x3_halo" if "halo" in k else "conv_generic")
    g = groups.setdefault(name, [0.0, 0])
    g[0] += v["ms"]
    g[1] += v["n"]
print(json.dumps({
    "workload": f"images -> mesh, POEM-{size}, {B} samples x {V} views (3x256x256), heatmap + DLT reference joints",
    "ms_per_step": ms, "ms_per_step_cuda_graph": ms_graph, "samples_per_s": B / ms * 1e3,
    "samples_per_s_cuda_graph": B / ms_graph * 1e3, "images_per_s": B * V / ms * 1e3,
    "image_stage_ms": ms_img, "decoder_and_dlt_ms": ms - ms_img, "image_stage_images_per_s": B * V / ms_img * 1e3,
    "backbone_nominal_tflops": gflop_img * B * V / ms_img, "finite": bool(torch.isfinite(out).all()),
    "kernel_groups_ms": {k: [round(v[0], 3), v[1]] for k, v in sorted(groups.items(), key=lambda kv: -kv[1][0])}}))
