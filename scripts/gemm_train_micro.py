"""One per-edge-sized TF32 GEMM of the training path (818176 x 256 x 256), NN and NT, timed with CUDA events or run once
under ncu:  python scripts/gemm_train_micro.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from poem_v2_b200 import _train_native as tn  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
M, N, K = 818176, 256, 256
A = torch.randn(M, K, device="cuda")
W = torch.randn(N, K, device="cuda")
C = torch.empty(M, N, device="cuda")
b = torch.randn(N, device="cuda")
for name, kw in (("NN fwd", dict()), ("NT dgrad", dict(b_mn=True)), ("NN fwd no-round", dict(round_ops=0)),
                 ("NT dgrad no-round", dict(b_mn=True, round_ops=0))):
    for _ in range(2):
        tn.gemm(A, W, C, M, N, K, bias=b, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        tn.gemm(A, W, C, M, N, K, bias=b, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:20s} {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s  A+C {(M * K + M * N) * 4 / ms / 1e6:.0f} GB/s")
dy = torch.randn(M, N, device="cuda")
dW = torch.zeros(N, K, device="cuda")
for name, kw in (("TT wgrad", dict()), ("TT wgrad no-round", dict(round_ops=0))):
    for _ in range(2):
        tn.gemm(dy, A, dW, N, K, M, a_mn=True, b_mn=True, accumulate=True, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        tn.gemm(dy, A, dW, N, K, M, a_mn=True, b_mn=True, accumulate=True, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:20s} {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s  dy+x {(M * K + M * N) * 4 / ms / 1e6:.0f} GB/s")
