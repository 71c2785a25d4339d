"""Host-side mirror of the HRNet-W40 backbone (reference lib/models/backbones/hrnet.py).

`HRNetStage4` is **stage 4** alone (`hrnet.py:272-277`): the three `HighResolutionModule`s (`hrnet.py:108-234`) that
turn [40@64², 80@32², 160@16², 320@8²] into the same four maps.  `HRNetW40` is the whole `HighResolutionNet.forward`
(`hrnet.py:385-420`): image (N,3,256,256) -> the same four maps, with the reference's parameter names from `conv1.` to
`stage4.` (the dead classification-head keys `incre_modules.* / downsamp_modules.* / final_layer.* / classifier.*`
are accepted and dropped).

`HRNetStage4` keeps the reference's parameter names (`{m}.branches.{b}.{k}.conv1.weight`, `...bn1.running_mean`,
`{m}.fuse_layers.{i}.{j}...` — the `stage4.` prefix of `HighResolutionNet`) so a backbone checkpoint loads unchanged;
`forward(x_list)` has the signature of `self.stage4(x_list)` (`hrnet.py:417`).  BatchNorm runs in eval mode (the
release configs freeze it: `FREEZE_BATCHNORM: true`) and is folded into the fp16 implicit-GEMM weights at pack time.
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
    """conv (+ optional bias) + optional eval BatchNorm -> [Cout_p, k*k*Cin_p] (K ordered ky, kx, c), bias [Cout_p]."""
    w = sd[conv_key + ".weight"].double()
    bias = sd[conv_key + ".bias"].double() if (conv_key + ".bias") in sd else torch.zeros(w.shape[0], dtype=torch.float64)
    if bn_key is not None:
        g, b = sd[bn_key + ".weight"].double(), sd[bn_key + ".bias"].double()
        mu, var = sd[bn_key + ".running_mean"].double(), sd[bn_key + ".running_var"].double()
        scale = g / torch.sqrt(var + BN_EPS)
        w = w * scale[:, None, None, None]
        bias = (bias - mu) * scale + b
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
            wt = nat.to_op16(w).contiguous().to(device)
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


# ------------------------------------------------------------------------------------------------ whole backbone
STAGE_MODULES = {2: 1, 3: 4, 4: 3}          # config/backbone/cls_hrnet_w40_*.yaml: NUM_MODULES of stages 2-4
_DEAD_PREFIXES = ("incre_modules.", "downsamp_modules.", "final_layer.", "classifier.")


def backbone_param_shapes(channels=CHANNELS):
    """name -> shape of every live parameter/buffer of reference `HighResolutionNet` (W40 yaml)."""
    s = {}

    def bn(prefix, c):
        s[prefix + ".weight"] = (c,)
        s[prefix + ".bias"] = (c,)
        s[prefix + ".running_mean"] = (c,)
        s[prefix + ".running_var"] = (c,)
        s[prefix + ".num_batches_tracked"] = ()
    s["conv1.weight"] = (64, 3, 3, 3)
    bn("bn1", 64)
    s["conv2.weight"] = (64, 64, 3, 3)
    bn("bn2", 64)
    for k in range(4):                       # layer1: Bottleneck(inplanes, 64) x4, expansion 4 (hrnet.py:70-104)
        cin = 64 if k == 0 else 256
        p = f"layer1.{k}."
        s[p + "conv1.weight"] = (64, cin, 1, 1)
        bn(p + "bn1", 64)
        s[p + "conv2.weight"] = (64, 64, 3, 3)
        bn(p + "bn2", 64)
        s[p + "conv3.weight"] = (256, 64, 1, 1)
        bn(p + "bn3", 256)
        if k == 0:
            s[p + "downsample.0.weight"] = (256, 64, 1, 1)
            bn(p + "downsample.1", 256)
    # transitions (hrnet.py:318-342)
    s["transition1.0.0.weight"] = (channels[0], 256, 3, 3)
    bn("transition1.0.1", channels[0])
    s["transition1.1.0.0.weight"] = (channels[1], 256, 3, 3)
    bn("transition1.1.0.1", channels[1])
    s["transition2.2.0.0.weight"] = (channels[2], channels[1], 3, 3)
    bn("transition2.2.0.1", channels[2])
    s["transition3.3.0.0.weight"] = (channels[3], channels[2], 3, 3)
    bn("transition3.3.0.1", channels[3])
    for stage, n_mod in STAGE_MODULES.items():
        for k, v in stage4_param_shapes(n_mod, channels[:stage]).items():
            s[f"stage{stage}.{k}"] = v
    return s


def _fold_stem(sd):
    """conv1 + bn1 -> fp32 [64, 27] with k = (ky*3 + kx)*3 + c, and fp32 bias [64]."""
    w = sd["conv1.weight"].double()
    scale = sd["bn1.weight"].double() / torch.sqrt(sd["bn1.running_var"].double() + BN_EPS)
    w = (w * scale[:, None, None, None]).permute(0, 2, 3, 1).reshape(64, 27)
    b = sd["bn1.bias"].double() - sd["bn1.running_mean"].double() * scale
    return w.float(), b.float()


class HRNetW40(nn.Module):
    """Drop-in for `HighResolutionNet` / `HRNet` (`hrnet.py:239-420,439-449`): `forward(img) -> [y0, y1, y2, y3]`."""

    def __init__(self, cfg=None, channels=CHANNELS):
        super().__init__()
        self.channels = tuple(channels)
        self.name = "HRNet"
        self._names = []
        for name, shape in backbone_param_shapes(channels).items():
            t = torch.zeros(shape, dtype=torch.long if name.endswith("num_batches_tracked") else torch.float32)
            if name.endswith("running_var") or (name.endswith(".weight") and len(shape) == 1):
                t = torch.ones(shape)
            self.register_buffer(name.replace(".", "__"), t)
            self._names.append(name)
        self._packed = None
        self._ws = None

    def state_dict(self, *a, prefix="", **k):
        return {prefix + n: getattr(self, n.replace(".", "__")) for n in self._names}

    def load_state_dict(self, sd, strict=True):
        live = set(self._names)
        missing = [n for n in self._names if n not in sd]
        unexpected = [n for n in sd if n not in live and not n.startswith(_DEAD_PREFIXES)]
        if strict and (missing or unexpected):
            raise RuntimeError(f"HRNetW40: missing keys {missing[:4]}, unexpected keys {unexpected[:4]}")
        for n in self._names:
            if n in sd:
                getattr(self, n.replace(".", "__")).copy_(sd[n])
        self._packed = None

    def _pack(self, device):
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        net = nat.PoemHRNet()
        keep = []

        def lin(w, b):
            wt = nat.to_op16(w).contiguous().to(device)
            bt = b.to(torch.float32).contiguous().to(device)
            keep.extend([wt, bt])
            return nat.PoemLinear(wt.data_ptr(), bt.data_ptr())
        ch = self.channels
        cp = [_pad64(c) for c in ch]
        for i, c in enumerate(ch):
            net.channels[i] = c
        w1, b1 = _fold_stem(sd)
        w1, b1 = w1.contiguous().to(device), b1.contiguous().to(device)
        keep.extend([w1, b1])
        net.stem1_w, net.stem1_b = w1.data_ptr(), b1.data_ptr()
        net.stem2 = lin(*_fold(sd, "conv2", "bn2", 64, 64))
        for k in range(4):
            p = f"layer1.{k}."
            cin = 64 if k == 0 else 256
            bt = net.layer1[k]
            bt.c1 = lin(*_fold(sd, p + "conv1", p + "bn1", cin, 64))
            bt.c2 = lin(*_fold(sd, p + "conv2", p + "bn2", 64, 64))
            bt.c3 = lin(*_fold(sd, p + "conv3", p + "bn3", 64, 256))
            if k == 0:
                bt.ds = lin(*_fold(sd, p + "downsample.0", p + "downsample.1", 64, 256))
        net.trans1[0] = lin(*_fold(sd, "transition1.0.0", "transition1.0.1", 256, cp[0]))
        net.trans1[1] = lin(*_fold(sd, "transition1.1.0.0", "transition1.1.0.1", 256, cp[1]))
        net.trans2 = lin(*_fold(sd, "transition2.2.0.0", "transition2.2.0.1", cp[1], cp[2]))
        net.trans3 = lin(*_fold(sd, "transition3.3.0.0", "transition3.3.0.1", cp[2], cp[3]))
        for stage, n_mod in STAGE_MODULES.items():
            mods = getattr(net, f"stage{stage}")
            for m in range(n_mod):
                mod = mods[m]
                for b in range(stage):
                    for k in range(4):
                        p = f"stage{stage}.{m}.branches.{b}.{k}."
                        mod.branch[b][k][0] = lin(*_fold(sd, p + "conv1", p + "bn1", cp[b], cp[b]))
                        mod.branch[b][k][1] = lin(*_fold(sd, p + "conv2", p + "bn2", cp[b], cp[b]))
                for i in range(stage):
                    for j in range(stage):
                        p = f"stage{stage}.{m}.fuse_layers.{i}.{j}."
                        if j > i:
                            mod.fuse[i][j][0] = lin(*_fold(sd, p + "0", p + "1", cp[j], cp[i]))
                        elif j < i:
                            for k in range(i - j):
                                co = cp[i] if k == i - j - 1 else cp[j]
                                mod.fuse[i][j][k] = lin(*_fold(sd, p + f"{k}.0", p + f"{k}.1", cp[j], co))
        self._packed = (net, keep, str(device))
        return net

    @torch.no_grad()
    def forward(self, x):
        if not x.is_cuda:
            raise nat.PoemError("HRNetW40 input must be a CUDA tensor: there is no CPU implementation")
        dev = x.device
        x = x.contiguous().float()
        n, c, h, w = x.shape
        assert c == 3 and h == w, tuple(x.shape)
        lib = nat.load()
        net = self._packed[0] if self._packed is not None and self._packed[2] == str(dev) else self._pack(dev)
        need = lib.poem_hrnet_workspace_bytes(C.byref(net), n, h)
        if need == 0:
            raise nat.PoemError(f"HRNetW40: unsupported image size {h}")
        if self._ws is None or self._ws.numel() < need + 1024 or self._ws.device != dev:
            self._ws = None
            self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
        off = (-self._ws.data_ptr()) % 1024
        outs = [torch.empty(n, ch, (h // 4) >> b, (h // 4) >> b, device=dev) for b, ch in enumerate(self.channels)]
        outs_p = (C.c_void_p * 4)(*[o.data_ptr() for o in outs])
        nat.check(lib.poem_hrnet_forward(C.byref(net), n, h, x.data_ptr(), outs_p, self._ws.data_ptr() + off,
                                         self._ws.numel() - off, torch.cuda.current_stream(dev).cuda_stream))
        return outs


def backbone_flops_per_image(img_res=256, channels=CHANNELS):
    """Nominal multiply-add FLOPs (2 * Cout * Cin * k^2 * Hout * Wout, real channel counts) of one image, by section."""
    base = img_res // 4
    out = {}

    def res_of(name):
        t = name.split(".")
        if t[0] == "conv1":
            return img_res // 2
        if t[0] in ("conv2", "layer1"):
            return base
        if t[0].startswith("transition"):
            return base >> int(t[1])
        if t[2] == "branches":
            return base >> int(t[3])
        i, j = int(t[3]), int(t[4])
        return base >> j if j > i else base >> (j + int(t[5]) + 1)
    for name, shape in backbone_param_shapes(channels).items():
        if len(shape) != 4:
            continue
        r = res_of(name)
        sec = name.split(".")[0]
        sec = "stem" if sec in ("conv1", "conv2") else ("transitions" if sec.startswith("transition") else sec)
        out[sec] = out.get(sec, 0.0) + 2.0 * shape[0] * shape[1] * shape[2] * shape[3] * r * r
    return out


# ------------------------------------------------------------------------------------------------ images -> mlvl_feat
FEAT_OUT = 160
N_JOINTS = 21
_OTHER_MODEL_PREFIXES = ("uv_in.", "ptEmb_head.", "mano_layer.", "img_backbone.incre_modules.",
                         "img_backbone.downsamp_modules.", "img_backbone.final_layer.", "img_backbone.classifier.")


def image_stage_param_shapes(channels=CHANNELS, out_channels=FEAT_OUT):
    """Live keys of the image half of reference `PtEmbedMultiviewStereoV2` (lib/models/POEM.py:100-105,169-194):
    `img_backbone.*` (HRNet), `feat_delayer.{0,1,2}.{conv,norm}.*`, `feat_in.conv.*`, `uv_delayer.*`, `uv_out.conv.*`
    (`uv_in.*` only produces `uv_feat`, which inference discards: dead)."""
    s = {"img_backbone." + k: v for k, v in backbone_param_shapes(channels).items()}
    for i in range(3):
        p = f"feat_delayer.{i}."
        s[p + "conv.weight"] = (channels[i + 1], channels[i], 3, 3)
        s[p + "conv.bias"] = (channels[i + 1],)
        for k, shp in (("weight", (channels[i + 1],)), ("bias", (channels[i + 1],)), ("running_mean", (channels[i + 1],)),
                       ("running_var", (channels[i + 1],)), ("num_batches_tracked", ())):
            s[p + "norm." + k] = shp
    s["feat_in.conv.weight"] = (out_channels, channels[3], 1, 1)
    s["feat_in.conv.bias"] = (out_channels,)
    for i in range(3):                       # uv_delayer (POEM.py:184-191): cat(up(x), skip) -> skip's channel count
        cin, cout = channels[3 - i] + channels[2 - i], channels[2 - i]
        p = f"uv_delayer.{i}."
        s[p + "conv.weight"] = (cout, cin, 3, 3)
        s[p + "conv.bias"] = (cout,)
        for k, shp in (("weight", (cout,)), ("bias", (cout,)), ("running_mean", (cout,)), ("running_var", (cout,)),
                       ("num_batches_tracked", ())):
            s[p + "norm." + k] = shp
    s["uv_out.conv.weight"] = (N_JOINTS, channels[0], 1, 1)
    s["uv_out.conv.bias"] = (N_JOINTS,)
    return s


_COUNTS_CACHE = {}


class ImageStage(nn.Module):
    """`extract_img_feat` + `feat_decode` + `heatmap_stage` of the reference model (POEM.py:189-229, 255-268): images
    (BN,3,256,256) -> `mlvl_feat` (BN,160,16,16), the tensor `POEM_Generalized_Head.forward` takes, and the 2-D joint
    estimates `pred_joints_uv` (BN,21,2); `triangulate` turns those into `reference_joints` (POEM.py:284-299).  Keys
    as in the full-model checkpoint; keys of the other parts of the model (`ptEmb_head.*`, `uv_in.*`, `mano_layer.*`)
    are accepted and ignored."""

    def __init__(self, channels=CHANNELS, out_channels=FEAT_OUT):
        super().__init__()
        self.channels, self.out_channels = tuple(channels), out_channels
        self._names = []
        for name, shape in image_stage_param_shapes(channels, out_channels).items():
            t = torch.zeros(shape, dtype=torch.long if name.endswith("num_batches_tracked") else torch.float32)
            if name.endswith("running_var") or (name.endswith(".weight") and len(shape) == 1):
                t = torch.ones(shape)
            self.register_buffer(name.replace(".", "__"), t)
            self._names.append(name)
        self._backbone = HRNetW40(channels=channels)
        self._packed = None
        self._ws = None

    def state_dict(self, *a, prefix="", **k):
        return {prefix + n: getattr(self, n.replace(".", "__")) for n in self._names}

    def load_state_dict(self, sd, strict=True):
        live = set(self._names)
        missing = [n for n in self._names if n not in sd]
        unexpected = [n for n in sd if n not in live and not n.startswith(_OTHER_MODEL_PREFIXES)]
        if strict and (missing or unexpected):
            raise RuntimeError(f"ImageStage: missing keys {missing[:4]}, unexpected keys {unexpected[:4]}")
        for n in self._names:
            if n in sd:
                getattr(self, n.replace(".", "__")).copy_(sd[n])
        self._packed = None

    def _pack(self, device):
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        self._backbone.load_state_dict({k[len("img_backbone."):]: v for k, v in sd.items()
                                        if k.startswith("img_backbone.")}, strict=True)
        net = self._backbone._pack(device)
        fd = nat.PoemFeatDecode()
        keep = []

        def lin(w, b):
            wt = nat.to_op16(w).contiguous().to(device)
            bt = b.to(torch.float32).contiguous().to(device)
            keep.extend([wt, bt])
            return nat.PoemLinear(wt.data_ptr(), bt.data_ptr())
        cp = [_pad64(c) for c in self.channels]
        for i in range(3):
            p = f"feat_delayer.{i}."
            fd.delayer[i] = lin(*_fold(sd, p + "conv", p + "norm", cp[i], cp[i + 1]))
        fd.feat_in = lin(*_fold(sd, "feat_in.conv", None, cp[3], _pad64(self.out_channels)))
        fd.out_channels = self.out_channels
        uv = nat.PoemUVDecode()
        ch = self.channels
        for i in range(3):
            p = f"uv_delayer.{i}."
            uv.delayer[i] = lin(*_fold(sd, p + "conv", p + "norm", _pad64(ch[3 - i] + ch[2 - i]), cp[2 - i]))
        ow = sd["uv_out.conv.weight"].reshape(N_JOINTS, ch[0]).float().contiguous().to(device)
        ob = sd["uv_out.conv.bias"].float().contiguous().to(device)
        keep.extend([ow, ob])
        uv.out_w, uv.out_b, uv.n_joints = ow.data_ptr(), ob.data_ptr(), N_JOINTS
        self._packed = (net, fd, keep, str(device), uv)
        return net, fd, uv

    @staticmethod
    @torch.no_grad()
    def triangulate(uv_px, cam_intr, cam_extr, cam_view_num):
        """`reference_joints` (B,21,3) from the 2-D estimates of each sample's views (POEM.py:284-299)."""
        if not uv_px.is_cuda:
            raise nat.PoemError("triangulate inputs must be CUDA tensors: there is no CPU implementation")
        dev = uv_px.device
        key = (tuple(int(v) for v in cam_view_num), str(dev))
        counts = _COUNTS_CACHE.get(key)
        if counts is None:          # cached so that a captured forward (poem_v2_b200.graph) performs no host->device copy
            counts = _COUNTS_CACHE[key] = torch.tensor(key[0], dtype=torch.int32, device=dev)
        n = sum(key[0])
        uv = uv_px.reshape(n, -1, 2).contiguous().float()
        k = cam_intr.reshape(n, 3, 3).to(dev).contiguous().float()
        e = cam_extr.reshape(n, 4, 4).to(dev).contiguous().float()
        out = torch.empty(len(counts), uv.shape[1], 3, device=dev)
        nat.check(nat.load().poem_triangulate_dlt(uv.data_ptr(), k.data_ptr(), e.data_ptr(), counts.data_ptr(),
                                                  len(counts), uv.shape[1], out.data_ptr(),
                                                  torch.cuda.current_stream(dev).cuda_stream))
        return out

    @torch.no_grad()
    def forward(self, img, return_maps=False, return_uv=False, return_heatmap=False):
        if not img.is_cuda:
            raise nat.PoemError("ImageStage input must be a CUDA tensor: there is no CPU implementation")
        dev = img.device
        img = img.reshape(-1, *img.shape[-3:]).contiguous().float()
        n, c, h, w = img.shape
        assert c == 3 and h == w, tuple(img.shape)
        lib = nat.load()
        if self._packed is not None and self._packed[3] == str(dev):
            net, fd, uv = self._packed[0], self._packed[1], self._packed[4]
        else:
            net, fd, uv = self._pack(dev)
        want_uv = return_uv or return_heatmap
        uv_ref = C.byref(uv) if want_uv else None
        need = lib.poem_image_features_workspace_bytes(C.byref(net), C.byref(fd), uv_ref, n, h)
        if need == 0:
            raise nat.PoemError(f"ImageStage: unsupported image size {h}")
        if self._ws is None or self._ws.numel() < need + 1024 or self._ws.device != dev:
            self._ws = None
            self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
        off = (-self._ws.data_ptr()) % 1024
        feat = torch.empty(n, self.out_channels, h // 16, h // 16, device=dev)
        maps, maps_p = None, None
        if return_maps:
            maps = [torch.empty(n, ch, (h // 4) >> b, (h // 4) >> b, device=dev) for b, ch in enumerate(self.channels)]
            maps_p = (C.c_void_p * 4)(*[o.data_ptr() for o in maps])
        uv_px = torch.empty(n, N_JOINTS, 2, device=dev) if want_uv else None
        heat = torch.empty(n, N_JOINTS, h // 8, h // 8, device=dev) if return_heatmap else None
        nat.check(lib.poem_image_features(C.byref(net), C.byref(fd), uv_ref, n, h, img.data_ptr(), feat.data_ptr(),
                                          uv_px.data_ptr() if want_uv else None,
                                          heat.data_ptr() if return_heatmap else None, maps_p,
                                          self._ws.data_ptr() + off, self._ws.numel() - off,
                                          torch.cuda.current_stream(dev).cuda_stream))
        if not (return_maps or want_uv):
            return feat
        out = {"mlvl_feat": feat}
        if return_maps:
            out["img_feats"] = maps
        if want_uv:
            out["pred_joints_uv"] = uv_px
        if return_heatmap:
            out["uv_hmap"] = heat
        return out
