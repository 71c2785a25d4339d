"""Experiment: mha_fwd_tc_kernel launch time vs query count (wave quantisation of the 128-query tiles, 2 CTAs/SM).
B = 32 samples, 4 heads of 64, 4096 keys; times by CUDA events around 20 launches."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poem_v2_b200 import _native as nat

lib = nat.load()
B, D, h, Lk = 32, 256, 4, 4096
g = torch.Generator(device="cuda").manual_seed(0)
K = torch.randn(B * Lk, D, device="cuda", generator=g).half()
V = torch.randn(B * Lk, D, device="cuda", generator=g).half()
st = torch.cuda.current_stream().cuda_stream
for Lq in (512, 640, 768, 799, 896, 1024):
    Q = torch.randn(B * Lq, D, device="cuda", generator=g).half()
    ctx = torch.zeros(B * Lq, D, device="cuda", dtype=torch.float16)
    def run():
        nat.check(lib.poem_mha(Q.data_ptr(), D, K.data_ptr(), D, V.data_ptr(), D, ctx.data_ptr(), D, B, Lq, Lk, D, h, st))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    tiles = (Lq + 127) // 128
    ctas = tiles * h * B
    print(f"Lq={Lq}: {ctas} CTAs = {ctas / 296:.2f} waves of 296, {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
