"""Round-2 CPU precision study: which operand formats meet max ||ours - ref|| / ||ref|| <= 1e-3 on the stress goldens?

Finer than precision_study.py: a policy maps every Linear of the state dict (by key) to (operand format, stored format),
plus the formats inside the attention core (Q.K^T operands, the probabilities P, V) and of the sampled features X.
Formats: fp32, tf32 (10-bit mantissa, rna), bf16, bf16x2 (hi + lo split: ~16 bits), bf16x3 (~24 bits).

    python scripts/precision_study2.py <policy-set> [case ...]
"""
import ast
import math
import os
import re
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as TF  # noqa: E402

import poem_oracle as orc  # noqa: E402
from golden_util import load_case  # noqa: E402
from poem_v2_b200 import synth  # noqa: E402


def rnd(t, fmt):
    if fmt == "fp32" or fmt is None:
        return t
    if fmt == "bf16":
        return t.bfloat16().float()
    if fmt == "bf16x2":
        hi = t.bfloat16().float()
        return hi + (t - hi).bfloat16().float()
    if fmt == "bf16x3":
        hi = t.bfloat16().float()
        mid = (t - hi).bfloat16().float()
        return hi + mid + (t - hi - mid).bfloat16().float()
    if fmt == "fp16":
        return t.half().float()
    assert fmt == "tf32", fmt
    keep = 13
    i = t.contiguous().view(torch.int32)
    i = i + (((i >> keep) & 1) + ((1 << (keep - 1)) - 1))
    return (i & ~((1 << keep) - 1)).view(torch.float32)


GROUPS = [   # (group name, regex on the weight key, optional predicate on the input)
    ("merge", r"^merge_net_feature\.", None),
    ("embed_pt", r"\.embedding\.weight$", lambda x: x.shape[-2] == 4096),
    ("embed_q", r"\.embedding\.weight$", lambda x: x.shape[-2] != 4096),
    ("mha_q", r"encoder\.(attn|cross_attn)\.self\.query", None),
    ("mha_kv", r"encoder\.(attn|cross_attn)\.self\.(key|value)", None),
    ("mha_o", r"encoder\.(attn|cross_attn)\.output\.dense", None),
    ("self_lin", r"query_self_attn\.(fc1|w_qs|w_ks|w_vs)\.", None),
    ("cross_q", r"query_cross_attn\.w_qs\.", None),
    ("cross_kv", r"query_cross_attn\.(fc1|w_ks|w_vs)\.", None),
    ("delta0", r"fc_delta\.0\.", None),
    ("delta2", r"fc_delta\.2\.", None),
    ("gamma0", r"fc_gamma\.0\.", None),
    ("gamma2", r"fc_gamma\.2\.", None),
    ("fc2", r"_attn\.fc2\.", None),
    ("reg0", r"reg_branch\.0\.", None),
    ("reg2", r"reg_branch\.2\.", None),
    ("ffn", r"encoder\.(intermediate|output)\.dense", None),
]


def run(policy, case):
    """policy: {group: fmt | (fmt_in, fmt_out)}, plus 'inproj', 'sampled', 'mha_qk', 'mha_p', 'mha_v', 'mha_ctx'."""
    meta, dims, sd, feat, metas, ref_j, gold = case
    names = {id(v): k for k, v in sd.items()}
    shim = types.SimpleNamespace(**{k: getattr(TF, k) for k in dir(TF) if not k.startswith("_")})

    def group_of(w, x):
        key = names.get(id(w), "?")
        for g, rx, pred in GROUPS:
            if re.search(rx, key) and (pred is None or pred(x)):
                return g
        raise KeyError(key)

    def linear(x, w, b=None):
        f = policy.get(group_of(w, x), "fp32")
        fi, fo = f if isinstance(f, tuple) else (f, "fp32")
        return rnd(TF.linear(rnd(x, fi), rnd(w, fi), b), fo)

    def conv2d(x, w, b=None, **kw):
        f = policy.get("inproj", "fp32")
        return TF.conv2d(rnd(x, f), rnd(w, f), b, **kw)

    def grid_sample(x, g, **kw):
        return rnd(TF.grid_sample(x, g, **kw), policy.get("sampled", "fp32"))

    shim.linear, shim.conv2d, shim.grid_sample = linear, conv2d, grid_sample

    def bert_cross_attention(sd_, prefix, hidden, enc, n_heads):
        B, Lq, D = hidden.shape
        hd = D // n_heads
        split = lambda t: t.view(B, -1, n_heads, hd).transpose(1, 2)  # noqa: E731
        q = split(linear(hidden, sd_[prefix + ".self.query.weight"], sd_[prefix + ".self.query.bias"]))
        k = split(linear(enc, sd_[prefix + ".self.key.weight"], sd_[prefix + ".self.key.bias"]))
        v = split(linear(enc, sd_[prefix + ".self.value.weight"], sd_[prefix + ".self.value.bias"]))
        fqk = policy.get("mha_qk", "fp32")
        s = rnd(q, fqk) @ rnd(k, fqk).transpose(-1, -2) / math.sqrt(hd)
        m = s.max(dim=-1, keepdim=True).values
        e = torch.exp(s - m)
        den = e.sum(dim=-1, keepdim=True)             # device: row sum of the un-rounded fp32 exponentials
        ctx = (rnd(e, policy.get("mha_p", "fp32")) @ rnd(v, policy.get("mha_v", "fp32"))) / den
        ctx = rnd(ctx.transpose(1, 2).reshape(B, Lq, D), policy.get("mha_ctx", "fp32"))
        o = linear(ctx, sd_[prefix + ".output.dense.weight"], sd_[prefix + ".output.dense.bias"])
        return TF.layer_norm(o + hidden, (D,), sd_[prefix + ".output.LayerNorm.weight"],
                             sd_[prefix + ".output.LayerNorm.bias"], eps=1e-12)

    old_f, old_b = orc.F, orc.bert_cross_attention
    orc.F, orc.bert_cross_attention = shim, bert_cross_attention
    try:
        with torch.no_grad():
            return orc.head_forward(sd, dims, feat, metas, ref_j, synth.standin_template(), *synth.load_assets())
    finally:
        orc.F, orc.bert_cross_attention = old_f, old_b


QSTREAM = ["embed_q", "mha_q", "mha_o", "self_lin", "cross_q", "fc2", "reg0", "ffn"]
POINT = ["embed_pt", "mha_kv", "cross_kv"]
TOKEN = ["delta2", "gamma0", "gamma2"]


def pol(**kw):
    p = {}
    for k, v in kw.items():
        if k == "q":
            p.update({g: v for g in QSTREAM})
        elif k == "pt":
            p.update({g: v for g in POINT})
        elif k == "tok":
            p.update({g: v for g in TOKEN})
        elif k == "mha":
            p.update(mha_qk=v, mha_p=v, mha_v=v)
        else:
            p[k] = v
    return p


SETS = {
    "base": {
        "all bf16 (r1 device model)": pol(inproj="bf16", sampled="bf16", merge=("bf16", "bf16"), q="bf16",
                                          pt=("bf16", "bf16"), tok="bf16", mha="bf16", mha_ctx="bf16"),
        "all tf32, fp32 stores": pol(inproj="tf32", merge="tf32", q="tf32", pt="tf32", tok="tf32", mha="tf32"),
        "tf32 except tok bf16": pol(inproj="tf32", merge="tf32", q="tf32", pt="tf32", tok="bf16", mha="tf32"),
        "only tok bf16": pol(tok="bf16"),
        "only tok tf32": pol(tok="tf32"),
        "only mha tf32": pol(mha="tf32"),
        "only mha bf16": pol(mha="bf16"),
        "only merge tf32": pol(merge="tf32"),
        "only q tf32": pol(q="tf32"),
        "only pt tf32": pol(pt="tf32"),
        "only inproj tf32": pol(inproj="tf32"),
        "only inproj bf16": pol(inproj="bf16"),
    },
    "x2": {
        "all bf16x2": pol(inproj="bf16x2", merge="bf16x2", q="bf16x2", pt="bf16x2", tok="bf16x2", mha="bf16x2"),
        "bf16x2 except tok tf32": pol(inproj="bf16x2", merge="bf16x2", q="bf16x2", pt="bf16x2", tok="tf32", mha="bf16x2"),
        "bf16x2 except tok bf16": pol(inproj="bf16x2", merge="bf16x2", q="bf16x2", pt="bf16x2", tok="bf16", mha="bf16x2"),
        "bf16x2, tok tf32, mha tf32": pol(inproj="bf16x2", merge="bf16x2", q="bf16x2", pt="bf16x2", tok="tf32", mha="tf32"),
        "bf16x2, tok tf32, merge tf32": pol(inproj="bf16x2", merge="tf32", q="bf16x2", pt="bf16x2", tok="tf32", mha="bf16x2"),
    },
}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "base"
    cases = sys.argv[2:] or ["medium_v8_b1", "large_v2_b1", "small_v2_b1", "small_ragged_b3"]
    torch.set_num_threads(8)
    for cname in cases:
        case = load_case(cname)
        ref = run({}, case)
        print(f"== {cname}: error vs the fp32 oracle", flush=True)
        print(f"{'policy':44s} {'mean mm per block':26s} {'worst mm':>9s} {'rel max':>9s} {'frac>1e-3':>10s}")
        for name, p in SETS[which].items():
            out = run(p, case)
            err = (out - ref).norm(dim=-1)
            rel = err / ref.norm(dim=-1)
            per_block = " / ".join(f"{e.mean().item() * 1e3:.4f}" for e in err)
            print(f"{name:44s} {per_block:26s} {err.max().item() * 1e3:9.3f} {rel.max().item():9.2e} "
                  f"{(rel > 1e-3).float().mean().item():10.4f}", flush=True)


if __name__ == "__main__":
    main()
