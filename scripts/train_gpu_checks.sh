#!/bin/bash
# training path on 1 GPU: step timings eager vs graph, then the bench sub-line alone
python - <<'PY'
import json, sys, time
sys.path.insert(0, ".")
import torch
from poem_v2_b200 import synth, _train_native as tn
from poem_v2_b200.config import release_dims
from poem_v2_b200.train import HeadTrainer, TrainStep
dims = release_dims("medium")
sd = synth.make_state_dict(dims, 0, "init")
out = []
for B in (4, 32):
    feat, metas, ref_j = synth.make_inputs(dims, B, [8] * B, 1)
    m = dict(metas); m["cam_intr"], m["cam_extr"] = metas["cam_intr"].cuda(), metas["cam_extr"].cuda()
    feat, ref_j = feat.cuda(), ref_j.cuda()
    gt_j, gt_v = ref_j.clone(), ref_j[:, 9:10] + 0.05 * torch.randn(B, 778, 3, device="cuda")
    for graph in (False, True):
        step = TrainStep(HeadTrainer(dims, sd, synth.standin_template()), graph=graph)
        for _ in range(3): step(feat, m, ref_j, gt_j, gt_v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(5): l = step(feat, m, ref_j, gt_j, gt_v)
        e1.record(); torch.cuda.synchronize()
        rec = dict(batch=B, graph=graph, ms_per_step=e0.elapsed_time(e1) / 5, wall_ms=(time.perf_counter() - t0) * 200, loss=float(l.item()))
        rec["samples_per_s"] = B / rec["ms_per_step"] * 1e3
        print(json.dumps(rec), flush=True); out.append(rec)
        del step; torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/train_step_eager_vs_graph.json", "w"), indent=1)
PY
