"""Training step of the decoder head (forward with saved activations + backward) timed with CUDA events.
    python scripts/bench_train.py [size] [views] [batch ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from poem_v2_b200 import _train_native as tn  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402
from poem_v2_b200.config import release_dims  # noqa: E402
from poem_v2_b200.train import HeadTrainer  # noqa: E402


def main():
    size = sys.argv[1] if len(sys.argv) > 1 else "medium"
    views = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    batches = [int(a) for a in sys.argv[3:]] or [4, 32]
    dims = release_dims(size)
    sd = synth.make_state_dict(dims, 0, "init")
    tr = HeadTrainer(dims, sd, synth.standin_template())
    lib = tn.load()
    out = []
    for B in batches:
        feat, metas, ref_j = synth.make_inputs(dims, B, [views] * B, 1)
        m = dict(metas)
        m["cam_intr"], m["cam_extr"] = metas["cam_intr"].cuda(), metas["cam_extr"].cuda()
        feat, ref_j = feat.cuda(), ref_j.cuda()
        dco = torch.randn(dims.n_blocks, B, dims.n_query, 3, device="cuda") * 1e-3
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        fw, bw, wall = [], [], []
        torch.cuda.reset_peak_memory_stats()
        n_warm = int(os.environ.get("POEM_TRAIN_WARM", "2"))
        n_it = int(os.environ.get("POEM_TRAIN_ITERS", "3"))
        for it in range(n_warm + n_it):
            tr.zero_grad()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            l0 = lib.poem_tr_kernel_launches()
            ev[0].record()
            tr.forward(feat, m, ref_j)
            ev[1].record()
            tr.backward(dco)
            ev[2].record()
            torch.cuda.synchronize()
            if it >= n_warm:
                fw.append(ev[0].elapsed_time(ev[1]))
                bw.append(ev[1].elapsed_time(ev[2]))
                wall.append((time.perf_counter() - t0) * 1e3)
            launches = lib.poem_tr_kernel_launches() - l0
        rec = dict(size=size, views=views, batch=B, fwd_ms=min(fw), bwd_ms=min(bw), step_ms=min(fw) + min(bw), wall_ms=min(wall),
                   samples_per_s=B / (min(fw) + min(bw)) * 1e3, launches=int(launches),
                   peak_mem_gb=torch.cuda.max_memory_allocated() / 2**30,
                   grad_finite=bool(all(torch.isfinite(g).all().item() for g in tr.g.values())))
        print(json.dumps(rec), flush=True)
        out.append(rec)
        if os.environ.get("POEM_TRAIN_PROF"):
            tn.profile(True)
            tr.zero_grad()
            tr.forward(feat, m, ref_j)
            tr.backward(dco)
            summ = tn.profile_summary()
            tn.profile(False)
            tot = sum(v[1] for v in summ.values())
            rec["profile_ms"] = {k: [v[0], round(v[1], 3)] for k, v in sorted(summ.items(), key=lambda kv: -kv[1][1])}
            print(f"-- per-primitive CUDA-event times, batch {B}: total {tot:.1f} ms")
            for k, v in list(rec["profile_ms"].items())[:40]:
                print(f"   {v[1]:8.2f} ms  n={v[0]:3d}  {k}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"train_step_{size}_v{views}.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
