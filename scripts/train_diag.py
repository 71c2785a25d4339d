"""Diagnostic (GPU): where do the training forward's activations leave the oracle's?  Prints, per block, the relative
L2 distance of the device's saved activations to the fp32 oracle and to the TF32-operand oracle on the same 32-NN sets."""
import ast
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import poem_oracle as orc  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402
from poem_v2_b200.config import release_dims  # noqa: E402
from poem_v2_b200.train import HeadTrainer  # noqa: E402
import test_train_gpu as ttg  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "grad_small_b2.npz"))
meta = ast.literal_eval(str(z["meta"]))
dims = release_dims(meta["size"])
mode = sys.argv[1] if len(sys.argv) > 1 else "stress"
sd = synth.make_state_dict(dims, meta["wseed"], mode)
feat, metas, ref_j = synth.make_inputs(dims, len(meta["views"]), meta["views"], meta["iseed"])
bps, a_xyz, a_idx = synth.load_assets()
tr = HeadTrainer(dims, sd, synth.standin_template())
coords = tr.forward(feat.cuda(), ttg._cuda_metas(metas), ref_j.cuda())
torch.cuda.synchronize()
nbr = tr.last_neighbours.long().cpu()
st32, sttf = {}, {}
with torch.no_grad():
    w32 = orc.head_forward(sd, dims, feat, metas, ref_j, synth.standin_template(), bps, a_xyz, a_idx, neighbours=nbr, stages=st32)
    with ttg._tf32_oracle(orc):
        wtf = orc.head_forward(sd, dims, feat, metas, ref_j, synth.standin_template(), bps, a_xyz, a_idx, neighbours=nbr, stages=sttf)


def rl2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


print("coords: dev-fp32 %.2e  dev-tf32oracle %.2e  tf32oracle-fp32 %.2e" % (rl2(coords.cpu(), w32), rl2(coords.cpu(), wtf), rl2(wtf, w32)))
B = len(meta["views"])
pt = tr.tape["blocks"][0]["pt_feats"].cpu().view(B, dims.n_sample, -1)
print("pt_feats: dev-fp32 %.2e  dev-tf32oracle %.2e  tf32oracle-fp32 %.2e" % (rl2(pt, st32["pt_feats"]), rl2(pt, sttf["pt_feats"]), rl2(sttf["pt_feats"], st32["pt_feats"])))
X = tr.tape["head"]["X"].cpu()
print("sampled: dev-fp32 %.2e" % rl2(X.view(-1), st32["sampled"].reshape(-1)))
for i in range(dims.n_blocks):
    t = tr.tape["blocks"][i]
    for name in ("a1", "a2", "f1", "f2"):
        dv = t[name].cpu().view(B, dims.n_query, -1)
        print(f"b{i}.{name}: dev-fp32 %.2e  dev-tf32oracle %.2e  tf32oracle-fp32 %.2e" % (rl2(dv, st32[f"b{i}.{name}"]), rl2(dv, sttf[f"b{i}.{name}"]), rl2(sttf[f"b{i}.{name}"], st32[f"b{i}.{name}"])))
    r_dev = t["r"].cpu() > 0
    p = f"transformer.pt_metro_encoder.{i}.encoder.vec_attn.reg_branch.0."
    for nm, stg in (("fp32", st32), ("tf32oracle", sttf)):
        r_or = torch.relu(torch.nn.functional.linear(stg[f"b{i}.f2"], sd[p + "weight"], sd[p + "bias"])).view(-1, r_dev.shape[1]) > 0
        print(f"b{i}.reg0 relu mask flips vs {nm}: %.4f %%" % (100.0 * (r_dev != r_or).float().mean().item()))
