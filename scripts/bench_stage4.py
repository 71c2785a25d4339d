"""HRNet-W40 stage 4 on B200: images/s and achieved TFLOP/s (12.45 GFLOP/image nominal)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poem_v2_b200 import _native as nat  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402
from poem_v2_b200.hrnet import HRNetStage4  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = HRNetStage4()
m.load_state_dict(synth.make_stage4_state_dict(0))
xs = [x.cuda() for x in synth.make_stage4_inputs(N, 64, 1)]
for _ in range(3):
    m(xs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    m(xs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
lib = nat.load()
lib.poem_profile_enable(1)
m(xs)
torch.cuda.synchronize()
prof = nat.profile_summary()
lib.poem_profile_enable(0)
flops = 12.45e9 * N
print(json.dumps({"workload": f"HRNet-W40 stage 4, {N} images 256x256 (maps 64/32/16/8)", "ms": ms,
                  "images_per_s": N / ms * 1e3, "nominal_tflops": flops / ms / 1e9,
                  "kernels_ms": {k: [round(v["ms"], 3), v["n"]] for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}))
