"""CPU study for round 2's precision work: which GEMM operand roundings of the decoder cost how much parity?

The oracle (fp32) is re-run with the operands of its `F.linear` calls rounded per STREAM CLASS, which emulates the
operand formats of the tensor-core GEMMs (accumulation stays fp32 like on the device):
  merge : merge-net MLPs on the sampled (token, view) rows        point : projections of the 4096 BPS tokens
  query : everything on the (B, 799, D) query stream              token : the (B, 799, 32, D) vector-attention MLPs
Formats: fp32 (exact), tf32 (10-bit mantissa), bf16 (7-bit mantissa); "<class>_out" additionally rounds what the GEMM
stores (the K / V / kt / v tables of the BPS tokens are bf16 in HBM).  Attention probabilities are not rounded here, so
"all bf16" is a LOWER bound of the device error.  Prints mean |x - fp32| in mm per block and the fraction of
points outside the 1e-3 relative bound, for POEM-medium, 8 views, one sample, "stress" weights (the golden case).

    python scripts/precision_study.py [size views]
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import torch.nn.functional as TF  # noqa: E402

import poem_oracle as orc  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402
from poem_v2_b200.config import release_dims  # noqa: E402


def rnd(t, fmt):
    if fmt == "fp32":
        return t
    if fmt == "bf16":
        return t.bfloat16().float()
    keep = 13                                               # tf32: drop the 13 low mantissa bits, round to nearest even
    i = t.contiguous().view(torch.int32)
    i = i + (((i >> keep) & 1) + ((1 << (keep - 1)) - 1))
    return (i & ~((1 << keep) - 1)).view(torch.float32)


def stream_class(x):
    if x.dim() == 4 and x.shape[1] == 799:
        return "token"
    if x.dim() == 3 and x.shape[1] == 799:
        return "query"
    if x.dim() == 3 and x.shape[1] == 4096:
        return "point"
    return "merge"


def run(policy, case):
    dims, sd, feat, metas, ref_j = case
    shim = types.SimpleNamespace(**{k: getattr(TF, k) for k in dir(TF) if not k.startswith("_")})

    def linear(x, w, b=None):
        cls = stream_class(x)
        fmt = policy.get(cls, "fp32")
        return rnd(TF.linear(rnd(x, fmt), rnd(w, fmt), b), policy.get(cls + "_out", "fp32"))   # "<class>_out": stored format
    shim.linear = linear
    old = orc.F
    orc.F = shim
    try:
        with torch.no_grad():
            return orc.head_forward(sd, dims, feat, metas, ref_j, synth.standin_template(), *synth.load_assets())
    finally:
        orc.F = old


def main():
    size = sys.argv[1] if len(sys.argv) > 1 else "medium"
    V = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    dims = release_dims(size)
    case = (dims, synth.make_state_dict(dims, 0, "stress"), *synth.make_inputs(dims, 1, [V], 1))
    ref = run({}, case)
    policies = {
        "all bf16": dict(merge="bf16", point="bf16", query="bf16", token="bf16"),
        "query tf32, rest bf16": dict(merge="bf16", point="bf16", query="tf32", token="bf16"),
        "query fp32, rest bf16": dict(merge="bf16", point="bf16", query="fp32", token="bf16"),
        "query + point tf32, rest bf16": dict(merge="bf16", point="tf32", query="tf32", token="bf16"),
        "only query bf16": dict(query="bf16"),
        "only point bf16": dict(point="bf16"),
        "only merge bf16": dict(merge="bf16"),
        "only token bf16": dict(token="bf16"),
        "all tf32": dict(merge="tf32", point="tf32", query="tf32", token="tf32"),
        "only point outputs stored bf16": dict(point_out="bf16"),
        "query + point tf32, point stored bf16, rest bf16": dict(merge="bf16", point="tf32", point_out="bf16", query="tf32",
                                                                 token="bf16"),
        "query + point + merge tf32, point stored bf16": dict(merge="tf32", point="tf32", point_out="bf16", query="tf32",
                                                              token="bf16"),
    }
    print(f"POEM-{size}, {V} views, 1 sample, stress weights; error vs the fp32 oracle")
    print(f"{'operand formats':50s} {'mean mm per block':28s} {'worst mm':>9s} {'frac rel>1e-3':>14s}")
    for name, pol in policies.items():
        out = run(pol, case)
        err = (out - ref).norm(dim=-1)
        rel = err / ref.norm(dim=-1)
        per_block = " / ".join(f"{e.mean().item() * 1e3:.4f}" for e in err)
        print(f"{name:50s} {per_block:28s} {err.max().item() * 1e3:9.3f} {(rel > 1e-3).float().mean().item():14.4f}")


if __name__ == "__main__":
    main()
