"""Decoder forward at the per-GPU batch sizes of BASELINE configs[3] under strong scaling (global batch 32 on 1..8 GPUs
= 32 / 16 / 8 / 4 samples per GPU): graph-replay time per step and the per-kernel breakdown at 4 samples."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from poem_v2_b200 import _native as nat  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
size = sys.argv[1] if len(sys.argv) > 1 else "medium_MANO"
rows = {}
for B in (4, 8, 16, 32):
    ms, steps, ok = bench.timed_head(size, 8, B, dev, 0, torch.cuda.synchronize)
    rows[B] = {"ms_per_step": round(ms, 4), "samples_per_s": round(B / ms * 1e3, 1), "finite": ok}
lib = nat.load()
dims, head = bench.make_head(size, dev)
feat, metas, ref_j = synth.make_inputs(dims, 4, 8, seed=3)
m = dict(metas)
m["cam_intr"], m["cam_extr"] = metas["cam_intr"].to(dev), metas["cam_extr"].to(dev)
f, r = feat.to(dev), ref_j.to(dev)
for _ in range(3):
    head(mlvl_feat=f, img_metas=m, reference_joints=r)
torch.cuda.synchronize()
lib.poem_profile_enable(1)
for _ in range(5):
    head(mlvl_feat=f, img_metas=m, reference_joints=r)
torch.cuda.synchronize()
prof = nat.profile_summary()
lib.poem_profile_enable(0)
kb = {k: [round(v["ms"] / 5, 4), v["n"] // 5] for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
print(json.dumps({"size": size, "views": 8, "by_batch": rows, "kernels_ms_per_step_at_batch_4": kb,
                  "sum_kernels_ms": round(sum(v[0] for v in kb.values()), 4)}))
