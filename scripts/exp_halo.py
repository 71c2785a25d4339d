"""Probe: 3x3 halo-reuse convolution, descriptor base-offset semantics and timing against the generic path."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poem_v2_b200 import _native as nat  # noqa: E402

lib = nat.load()


def run(N, R, c, cp, relu, res, mode, timing=False):
    g = torch.Generator().manual_seed(R + c)
    x = torch.randn(N, c, R, R, generator=g).half().float()
    w = (torch.randn(c, c, 3, 3, generator=g) / (c * 9) ** 0.5).half().float()
    b = 0.1 * torch.randn(c, generator=g)
    r = torch.randn(N, c, R, R, generator=g).half().float() if res else None
    xp = torch.zeros(N, R, R, cp)
    xp[..., :c] = x.permute(0, 2, 3, 1)
    wp = torch.zeros(cp, 3, 3, cp)
    wp[:c, :, :, :c] = w.permute(0, 2, 3, 1)
    bp = torch.zeros(cp)
    bp[:c] = b
    rp = None
    if res:
        rp = torch.zeros(N, R, R, cp)
        rp[..., :c] = r.permute(0, 2, 3, 1)
    xd, wd = xp.half().cuda(), wp.reshape(cp, -1).half().contiguous().cuda()
    rd = rp.half().cuda() if res else None
    bd = bp.cuda()
    out = torch.full((N, R, R, cp), float("nan"), device="cuda", dtype=torch.float16)
    lib.poem_debug_conv_mode(mode)
    st = torch.cuda.current_stream().cuda_stream

    def call():
        nat.check(lib.poem_conv_nhwc(xd.data_ptr(), N, R, R, cp, wd.data_ptr(), bd.data_ptr(), cp, 3, 1, int(relu),
                                     rd.data_ptr() if res else None, out.data_ptr(), c, c, st))
    call()
    torch.cuda.synchronize()
    ms = None
    if timing:
        for _ in range(3):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
    lib.poem_debug_conv_mode(0)
    if N > 8:
        return None, ms
    ref = F.conv2d(x, w, b, padding=1)
    if res:
        ref = ref + r
    if relu:
        ref = F.relu(ref)
    got = out.float().cpu()
    err = (got[..., :c].permute(0, 3, 1, 2) - ref).abs().max().item()
    return err / max(1.0, ref.abs().max().item()), ms


for (R, c, cp) in [(64, 40, 64), (32, 80, 128), (16, 160, 192), (32, 40, 64), (16, 80, 128)]:
    for mode in (0, 1, 2):
        e, _ = run(3, R, c, cp, True, True, mode)
        print(f"R={R} C={c}->{cp} mode {mode}: rel err {e:.3e}", flush=True)
for (R, c, cp) in [(64, 40, 64), (32, 80, 128), (16, 160, 192)]:
    for mode in (0, 1, 2):
        for res in (False, True):
            _, ms = run(256, R, c, cp, True, res, mode, timing=True)
            print(f"N=256 R={R} Cp={cp} mode {mode} res={res}: {ms * 1e3:.1f} us", flush=True)
