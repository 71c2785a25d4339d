"""One MHA shape (B=32, 799 queries, 4 heads of 64, 4096 keys): time 20 launches (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poem_v2_b200 import _native as nat
lib = nat.load()
B, D, h, Lk, Lq = 32, 256, 4, 4096, 799
g = torch.Generator(device="cuda").manual_seed(0)
K = torch.randn(B * Lk, 6 * D, device="cuda", generator=g).half()      # strided like the KK table of the decoder
Q = torch.randn(B * Lq, D, device="cuda", generator=g).half()
ctx = torch.zeros(B * Lq, D, device="cuda", dtype=torch.float16)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run():
    nat.check(lib.poem_mha(Q.data_ptr(), D, K.data_ptr(), 6 * D, K.data_ptr() + 8 * D, 6 * D, ctx.data_ptr(), D, B, Lq, Lk, D, h, st))
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for _ in range(int(os.environ.get("N", "20"))):
    flush.zero_()
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print(f"{tot / int(os.environ.get('N', '20')) * 1e3:.1f} us per launch (L2 flushed before each)")
