"""Phase timeline of the training GEMM's CTAs (debug build with -DPOEM_TG_TRACE=1, POEM_TRAIN_LIB pointing at it):
per CTA, ns from kernel start: setup done, first stage landed, last stage landed, accumulator complete, epilogue done."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from poem_v2_b200 import _train_native as tn  # noqa: E402

M, N, K = 818176, 256, 256
A, W, Cc = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.empty(M, N, device="cuda")
for _ in range(2):
    tn.gemm(A, W, Cc, M, N, K, round_ops=0)
torch.cuda.synchronize()
lib = tn.load()
buf = np.zeros(8192 * 16, dtype=np.uint64)
lib.poem_tr_debug_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.poem_tr_debug_trace(buf.ctypes.data, buf.size) == 0
t = buf.reshape(8192, 16).astype(np.int64)
t = t[t[:, 0] > 0]                      # CTAs of the first n-tile column only write the chunk-0 marks
d = lambda a, b: (t[:, a] - t[:, b]) / 1e3   # noqa: E731  (us)
rows = {"setup (start -> barriers + TMEM ready)": d(1, 0), "first stage landed after setup": d(2, 1), "first -> last stage landed": d(6, 2),
        "last stage landed -> accumulator complete": d(3, 6), "epilogue": d(4, 3), "epilogue end -> TMEM freed": d(7, 4), "whole CTA": d(7, 0),
        "  chunk 0: tmem load + wait": d(8, 3), "  chunk 0: transpose into the pad": d(9, 8), "  chunk 0: bias loads": d(10, 9),
        "  chunk 0: read back + stores issued": d(11, 10)}
for k, v in rows.items():
    print(f"{k:45s} median {np.median(v):7.2f} us   p10 {np.percentile(v, 10):7.2f}   p90 {np.percentile(v, 90):7.2f}")
