"""Whole HRNet-W40 backbone on B200: images/s and achieved TFLOP/s against the nominal FLOP count."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poem_v2_b200 import _native as nat  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402
from poem_v2_b200.hrnet import HRNetW40, backbone_flops_per_image  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = HRNetW40()
m.load_state_dict(synth.make_backbone_state_dict(0))
img = synth.make_images(N, 256, 1).cuda()
for _ in range(3):
    m(img)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    m(img)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
lib = nat.load()
lib.poem_profile_enable(1)
m(img)
torch.cuda.synchronize()
prof = nat.profile_summary()
lib.poem_profile_enable(0)
fl = backbone_flops_per_image()
flops = sum(fl.values()) * N
print(json.dumps({"workload": f"HRNet-W40 backbone, {N} images 3x256x256 -> maps 64/32/16/8", "ms": ms,
                  "images_per_s": N / ms * 1e3, "gflop_per_image": {k: round(v / 1e9, 3) for k, v in fl.items()},
                  "nominal_tflops": flops / ms / 1e9, "workspace_gb": m._ws.numel() / 1e9,
                  "kernels_ms": {k: [round(v["ms"], 3), v["n"]] for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}))
