"""Host-side mirror of HRNet-W40 **stage 4** (reference lib/models/backbones/hrnet.py:272-277): the three
`HighResolutionModule`s (`hrnet.py:108-234`) that turn [40@64², 80@32², 160@16², 320@8²] into the same four maps.

`HRNetStage4` keeps the reference's parameter names (`{m}.branches.{b}.{k}.conv1.weight`, `...bn1.running_mean`,
`{m}.fuse_layers.{i}.{j}...` — the `stage4.` prefix of `HighResolutionNet`) so a backbone checkpoint loads unchanged;
`forward(x_list)` has the signature of `self.stage4(x_list)` (`hrnet.py:417`).  BatchNorm runs in eval mode (the
release configs freeze it: `FREEZE_BATCHNORM: true`) and is folded into the bf16 implicit-GEMM weights at pack time.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _native as nat

CHANNELS = (40, 80, 160, 320)
BN_EPS = 1e-5


def _pad64(c):
    return (c + 63) // 64 * 64


def stage4_param_shapes(n_modules=3, channels=CHANNELS):
    """name -> shape of every parameter/buffer of reference `stage4` (relative to the `stage4.` prefix)."""
    s = {}

    def bn(prefix, c):
        s[prefix + ".weight"] = (c,)
        s[prefix + ".bias"] = (c,)
        s[prefix + ".running_mean"] = (c,)
        s[prefix + ".running_var"] = (c,)
        s[prefix + ".num_batches_tracked"] = ()
    for m in range(n_modules):
        for b, c in enumerate(channels):
            for k in range(4):
                p = f"{m}.branches.{b}.{k}."
                s[p + "conv1.weight"] = (c, c, 3, 3)
                bn(p + "bn1", c)
                s[p + "conv2.weight"] = (c, c, 3, 3)
                bn(p + "bn2", c)
        for i, ci in enumerate(channels):
            for j, cj in enumerate(channels):
                p = f"{m}.fuse_layers.{i}.{j}."
                if j > i:
                    s[p + "0.weight"] = (ci, cj, 1, 1)
                    bn(p + "1", ci)
                elif j < i:
                    for k in range(i - j):
                        co = ci if k == i - j - 1 else cj
                        s[p + f"{k}.0.weight"] = (co, cj, 3, 3)
                        bn(p + f"{k}.1", co)
    return s


def _fold(sd, conv_key, bn_key, cin_p, cout_p):
    """conv (no bias) + eval BatchNorm -> bf16 [Cout_p, k*k*Cin_p] (K ordered ky, kx, c) and fp32 bias [Cout_p]."""
    w = sd[conv_key + ".weight"].double()
    g, b = sd[bn_key + ".weight"].double(), sd[bn_key + ".bias"].double()
    mu, var = sd[bn_key + ".running_mean"].double(), sd[bn_key + ".running_var"].double()
    scale = g / torch.sqrt(var + BN_EPS)
    w = w * scale[:, None, None, None]
    bias = b - mu * scale
    co, ci, kh, kw = w.shape
    wp = torch.zeros(cout_p, kh, kw, cin_p, dtype=torch.float64)
    wp[:co, :, :, :ci] = w.permute(0, 2, 3, 1)
    bp = torch.zeros(cout_p, dtype=torch.float64)
    bp[:co] = bias
    return wp.reshape(cout_p, kh * kw * cin_p), bp


class HRNetStage4(nn.Module):
    def __init__(self, n_modules=3, channels=CHANNELS):
        super().__init__()
        self.n_modules, self.channels = n_modules, tuple(channels)
        self._names = []
        for name, shape in stage4_param_shapes(n_modules, channels).items():
            t = torch.zeros(shape, dtype=torch.long if name.endswith("num_batches_tracked") else torch.float32)
            if name.endswith("running_var") or (name.endswith(".weight") and len(shape) == 1):
                t = torch.ones(shape)
            self.register_buffer(name.replace(".", "__"), t)
            self._names.append(name)
        self._packed = None
        self._ws = None

    # reference key names contain dots: expose them through state_dict()/load_state_dict()
    def state_dict(self, *a, prefix="", **k):
        return {prefix + n: getattr(self, n.replace(".", "__")) for n in self._names}

    def load_state_dict(self, sd, strict=True):
        missing = [n for n in self._names if n not in sd]
        unexpected = [n for n in sd if n not in self._names]
        if strict and (missing or unexpected):
            raise RuntimeError(f"HRNetStage4: missing keys {missing[:4]}, unexpected keys {unexpected[:4]}")
        for n in self._names:
            if n in sd:
                getattr(self, n.replace(".", "__")).copy_(sd[n])
        self._packed = None

    def _pack(self, device):
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        st = nat.PoemHRStage4()
        st.n_modules = self.n_modules
        keep = []

        def lin(w, b):
            wt = w.to(torch.bfloat16).contiguous().to(device)
            bt = b.to(torch.float32).contiguous().to(device)
            keep.extend([wt, bt])
            return nat.PoemLinear(wt.data_ptr(), bt.data_ptr())
        ch = self.channels
        for i, c in enumerate(ch):
            st.channels[i] = c
        for m in range(self.n_modules):
            mod = st.modules[m]
            for b, c in enumerate(ch):
                cp = _pad64(c)
                for k in range(4):
                    p = f"{m}.branches.{b}.{k}."
                    mod.branch[b][k][0] = lin(*_fold(sd, p + "conv1", p + "bn1", cp, cp))
                    mod.branch[b][k][1] = lin(*_fold(sd, p + "conv2", p + "bn2", cp, cp))
            for i, ci in enumerate(ch):
                for j, cj in enumerate(ch):
                    p = f"{m}.fuse_layers.{i}.{j}."
                    if j > i:
                        mod.fuse[i][j][0] = lin(*_fold(sd, p + "0", p + "1", _pad64(cj), _pad64(ci)))
                    elif j < i:
                        for k in range(i - j):
                            co = ci if k == i - j - 1 else cj
                            mod.fuse[i][j][k] = lin(*_fold(sd, p + f"{k}.0", p + f"{k}.1", _pad64(cj), _pad64(co)))
        self._packed = (st, keep, str(device))
        return st

    @torch.no_grad()
    def forward(self, x_list):
        assert len(x_list) == 4
        x0 = x_list[0]
        if not x0.is_cuda:
            raise nat.PoemError("HRNetStage4 inputs must be CUDA tensors: there is no CPU implementation")
        dev = x0.device
        n, base = x0.shape[0], x0.shape[-1]
        xs = [x.contiguous().float() for x in x_list]
        for b, x in enumerate(xs):
            assert tuple(x.shape) == (n, self.channels[b], base >> b, base >> b), tuple(x.shape)
        lib = nat.load()
        st = self._packed[0] if self._packed is not None and self._packed[2] == str(dev) else self._pack(dev)
        need = lib.poem_hrnet_stage4_workspace_bytes(C.byref(st), n, base)
        if self._ws is None or self._ws.numel() < need + 1024 or self._ws.device != dev:
            self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
        off = (-self._ws.data_ptr()) % 1024
        outs = [torch.empty_like(x) for x in xs]
        ins_p = (C.c_void_p * 4)(*[x.data_ptr() for x in xs])
        outs_p = (C.c_void_p * 4)(*[o.data_ptr() for o in outs])
        nat.check(lib.poem_hrnet_stage4_forward(C.byref(st), n, base, ins_p, outs_p, self._ws.data_ptr() + off,
                                                self._ws.numel() - off, torch.cuda.current_stream(dev).cuda_stream))
        return outs
