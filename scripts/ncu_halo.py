"""One profiled launch of each halo-reuse convolution shape of HRNet-W40 at 256 images (run under
`ncu --profile-from-start off --set full`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poem_v2_b200 import _native as nat  # noqa: E402

lib = nat.load()
N = 256
st = torch.cuda.current_stream().cuda_stream
for (R, c, cp) in [(64, 40, 64), (32, 80, 128), (16, 160, 192)]:
    g = torch.Generator().manual_seed(R)
    x = torch.randn(N, R, R, cp, generator=g).half().cuda()
    x[..., c:] = 0
    r = torch.randn(N, R, R, cp, generator=g).half().cuda()
    r[..., c:] = 0
    w = torch.zeros(cp, 3, 3, cp)
    w[:c, :, :, :c] = torch.randn(c, 3, 3, c, generator=g) / (9 * c) ** 0.5
    w = w.reshape(cp, -1).half().cuda()
    b = torch.zeros(cp, device="cuda")
    out = torch.empty_like(x)

    def call():
        nat.check(lib.poem_conv_nhwc(x.data_ptr(), N, R, R, cp, w.data_ptr(), b.data_ptr(), cp, 3, 1, 1, r.data_ptr(),
                                     out.data_ptr(), c, c, st))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    call()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
