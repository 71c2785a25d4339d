"""Experiment: CPU time needed to ENQUEUE one decoder forward (eager launches through the C-ABI) vs the GPU time it takes.
If enqueue time >= GPU time the eager / host-buffer path is CPU-bound."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poem_v2_b200 import synth
from poem_v2_b200.config import release_dims
from poem_v2_b200.head import POEM_Generalized_Head

dims = release_dims("medium")
head = POEM_Generalized_Head(dims, template_mesh=synth.standin_template())
head.load_state_dict(synth.make_state_dict(dims, 0), strict=True)
head = head.cuda().eval()
feat, metas, ref_j = synth.make_inputs(dims, 32, 8, 1)
hm = dict(metas); hm["cam_intr"], hm["cam_extr"] = metas["cam_intr"].pin_memory(), metas["cam_extr"].pin_memory()
hf, hr = feat.pin_memory(), ref_j.pin_memory()
dm = dict(metas); dm["cam_intr"], dm["cam_extr"] = metas["cam_intr"].cuda(), metas["cam_extr"].cuda()
df, dr = feat.cuda(), ref_j.cuda()
out = torch.empty(3, 32, 799, 3).pin_memory()
for name, fn in (("device entry (head.forward)", lambda: head(mlvl_feat=df, img_metas=dm, reference_joints=dr)),
                 ("host entry (head.forward_host)", lambda: head.forward_host(hf, hm, hr, out=out))):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: enqueue {1e3 * (t1 - t0) / 20:.3f} ms per call, wall incl. GPU drain {1e3 * (t2 - t0) / 20:.3f} ms per call")
