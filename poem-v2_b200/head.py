"""Host-side mirror of the reference plug-in interface for the decoder path.

`POEM_Generalized_Head(cfg)` / `PtEmbedTRv4(cfg)` keep the reference's class names, `cfg` constructor,
`forward()` signatures, return values, error behaviour and state-dict keys
(reference lib/models/heads/ptEmb_head.py:683-964, lib/models/layers/ptEmb_transformer.py:303-376), so they can be
registered under the same names in the reference registries (lib/utils/builder.py:252-304) — see INTEGRATION.md.
All arithmetic happens in libpoem_b200.so; there is no PyTorch/CPU fallback.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _native as nat
from . import params as _params
from .config import HeadDims, dims_from_cfg
from .pack import MANO_KEYS, PackedManoTail, PackedWeights, mano_zero_pose_template

MAX_VIEWS = 10   # reference README: evaluated with 1..8 (up to 10 in the view sweep) cameras per sample


class _Tree(nn.Module):
    """Container whose children/parameters are created from dotted reference key names."""

    def add(self, dotted, tensor):
        head, _, rest = dotted.partition(".")
        if not rest:
            self.register_parameter(head, nn.Parameter(tensor, requires_grad=False))
            return
        if head not in self._modules:
            self.add_module(head, _Tree())
        self._modules[head].add(rest, tensor)


def _manotorch_layer(dims):
    from manotorch.manolayer import ManoLayer  # type: ignore
    return ManoLayer(joint_rot_mode="axisang", use_pca=False, mano_assets_root="assets/mano_v1_2",
                     center_idx=dims.center_idx, flat_hand_mean=True)


def _resolve_template(dims, template):
    if template is not None:
        t = torch.as_tensor(template, dtype=torch.float32).reshape(dims.n_query, 3)
        return t
    try:  # the reference builds it from manotorch on every forward (ptEmb_head.py:886-894)
        out = _manotorch_layer(dims)(torch.zeros(1, 48), torch.zeros(1, 10))
        return torch.cat([out.joints, out.verts], dim=1)[0].float()
    except Exception as e:  # noqa: BLE001
        raise RuntimeError("MANO template unavailable (manotorch / assets/mano_v1_2 missing): pass "
                           "`template_mesh=(799,3)` to the head or call `set_template()`") from e


def _resolve_mano(dims, mano):
    """MANO model parameters for the parametric tail: a dict with MANO_KEYS (roles of manotorch's `th_*` buffers),
    else the buffers of an installed manotorch layer."""
    if mano is None:
        try:
            layer = _manotorch_layer(dims)
            mano = {k: getattr(layer, "th_" + k) for k in MANO_KEYS}
        except Exception as e:  # noqa: BLE001
            raise RuntimeError("MANO parameters unavailable (manotorch / assets/mano_v1_2 missing): pass `mano_params=` "
                               "(dict with v_template, shapedirs, posedirs, J_regressor, weights), call `set_mano()`, "
                               "or load a checkpoint that carries the `mano_layer.th_*` buffers") from e
    missing = [k for k in MANO_KEYS if k not in mano]
    if missing:
        raise KeyError(f"mano_params lacks {missing}")
    out = {k: torch.as_tensor(mano[k], dtype=torch.float32) for k in MANO_KEYS}
    assert out["v_template"].numel() == 778 * 3 and out["shapedirs"].numel() == 778 * 3 * 10
    assert out["posedirs"].numel() == 778 * 3 * 135 and out["J_regressor"].numel() == 16 * 778
    assert out["weights"].numel() == 778 * 16
    return out


class _NativeDecoder(nn.Module):
    """Parameters under the reference's names + lazily packed kernel weights + workspace cache."""

    def __init__(self, dims: HeadDims, key_prefix: str, template_mesh=None, mano_params=None):
        super().__init__()
        self.dims = dims
        self._mano = None if mano_params is None else _resolve_mano(dims, mano_params)
        self._packed_mano = None
        self._key_prefix = key_prefix           # "" for the head, "transformer." stripped for PtEmbedTRv4
        shapes = _params.live_param_shapes(dims)
        self._live = []
        for name, shape in shapes.items():
            if not name.startswith(key_prefix):
                continue
            local = name[len(key_prefix):]
            self._live.append(local)
            self.add_param(local, torch.zeros(shape))
        bps, a_xyz, a_idx = _params.load_assets()
        self.register_buffer("bps_points", bps, persistent=False)
        self.register_buffer("anchor_xyz", a_xyz, persistent=False)
        self.register_buffer("anchor_idx", a_idx, persistent=False)
        self._template = None if template_mesh is None else _resolve_template(dims, template_mesh)
        self._packed = None
        self._packed_key = None
        self._ws = None
        self.debug_export_neighbours = False    # test hook (poem_debug_export_neighbours): fills `last_neighbours`
        self.last_neighbours = None

    def add_param(self, dotted, tensor):
        head, _, rest = dotted.partition(".")
        if not rest:
            self.register_parameter(head, nn.Parameter(tensor, requires_grad=False))
            return
        if head not in self._modules:
            self.add_module(head, _Tree())
        self._modules[head].add(rest, tensor)

    # dead reference keys (BERT word tables, pooler, unused head layers) are accepted and ignored
    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        live = {prefix + k for k in self._live}
        if self.dims.parametric and self._mano is None:    # a reference checkpoint carries the MANO layer's buffers
            tag = f"pt_metro_encoder.{self.dims.n_blocks - 1}.mano_layer.th_"
            found = {k.rsplit("th_", 1)[1]: v for k, v in state_dict.items() if k.startswith(prefix) and tag in k}
            if all(k in found for k in MANO_KEYS):
                self.set_mano(found)
        for k in [k for k in state_dict if k.startswith(prefix) and k not in live]:
            del state_dict[k]
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    def set_template(self, template_mesh):
        self._template = _resolve_template(self.dims, template_mesh)
        self._packed = None

    def set_mano(self, mano_params):
        """MANO model parameters of the parametric tail (roles of manotorch's `th_*` buffers)."""
        self._mano = _resolve_mano(self.dims, mano_params)
        self._packed = self._packed_mano = None

    def packed_mano(self, device):
        """Call after `packed(device)`: a re-pack of the weights drops this cache too."""
        if self._packed is None:
            self.packed(device)
        if self._packed_mano is None:
            if self._mano is None:
                self._mano = _resolve_mano(self.dims, None)
            sd = {self._key_prefix + k: v for k, v in self.state_dict().items()}
            self._packed_mano = PackedManoTail(sd, self.dims, self._mano, device)
        return self._packed_mano

    def live_state(self):
        sd = self.state_dict()
        return {self._key_prefix + k: sd[k] for k in self._live}

    def packed(self, device):
        params = [p for p in self.parameters()]
        key = (str(device), tuple((p.data_ptr(), p._version) for p in params))
        if self._packed is None or self._packed_key != key:
            self._packed_mano = None
            if self._template is None:
                if self.dims.parametric and self._mano is not None:
                    self._template = mano_zero_pose_template(self._mano, self.dims.center_idx)
                else:
                    self._template = _resolve_template(self.dims, None)
            full = {}
            for name, shape in _params.live_param_shapes(self.dims).items():   # head-only keys absent for the TR class
                full[name] = torch.zeros(shape)
            full.update(self.live_state())
            self._packed = PackedWeights(full, self.dims, device, self._template, self.bps_points, self.anchor_xyz,
                                         self.anchor_idx, MAX_VIEWS)
            self._packed_key = key
        return self._packed

    def workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != torch.device(device):
            self._ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        off = (-self._ws.data_ptr()) % 1024
        return self._ws.data_ptr() + off, self._ws.numel() - off


def _require_cuda(t, what):
    if not t.is_cuda:
        raise nat.PoemError(f"{what} must be a CUDA tensor: the decoder path has no CPU implementation")


class POEM_Generalized_Head(_NativeDecoder):
    """Drop-in for the reference head: `forward(mlvl_feat, img_metas, reference_joints, **kwargs)` ->
    {"all_coords_preds": (NB,B,799,3) metres}."""

    def __init__(self, cfg, template_mesh=None, mano_params=None):
        dims = cfg if isinstance(cfg, HeadDims) else dims_from_cfg(cfg)
        super().__init__(dims, "", template_mesh, mano_params)
        self.num_preds = dims.n_blocks            # read by the model shell (reference POEM.py:115)
        self.parametric_output = dims.parametric

    # ---- training: autograd flows through the hand-written backward of poem_v2_b200/train.py ----
    def train(self, mode=True):
        """`head.train()` makes the live parameters trainable (the reference's are by default); `eval()` leaves them."""
        super().train(mode)
        if mode:
            self._train_requested = True       # a freshly built module is `training` too, but only an explicit train() opts in
            for p in self.parameters():
                p.requires_grad_(True)
        return self

    def trainer(self):
        """The HeadTrainer behind the training-mode forward.  On first use the module's parameters are re-homed into the
        trainer's flat fp32 buffer (`param.data` becomes a view of it): optimisers, DDP and `state_dict()` keep working on
        the same nn.Parameters, the kernels read the same memory, and the inference path re-packs when they change."""
        dev = next(self.parameters()).device
        tr = getattr(self, "_trainer", None)
        if tr is None or tr.dev != dev:
            from .train import HeadTrainer
            if dev.type != "cuda":
                raise nat.PoemError("training needs the module on a CUDA device: there is no CPU implementation")
            if self.dims.parametric and self._mano is None:
                self._mano = _resolve_mano(self.dims, None)
            if self._template is None:
                if self.dims.parametric:
                    self._template = mano_zero_pose_template(self._mano, self.dims.center_idx)
                else:
                    self._template = _resolve_template(self.dims, None)
            tr = HeadTrainer(self.dims, self.live_state(), self._template,
                             assets=(self.bps_points, self.anchor_xyz, self.anchor_idx), device=dev,
                             mano=self._mano if self.dims.parametric else None)
            named = dict(self.named_parameters())
            for k in tr.p:
                named[k[len(self._key_prefix):]].data = tr.p[k]
            self._trainer = tr
            self._trainer_params = [named[k[len(self._key_prefix):]] for k in tr.p]
        return tr

    def forward(self, mlvl_feat, img_metas, reference_joints, **kwargs):
        if self.training and torch.is_grad_enabled() and getattr(self, "_train_requested", False):
            from .train import HeadFunction
            assert int(np.sum(np.asarray(img_metas["master_id"]))) == 0, "only support master_id is 0"
            _require_cuda(mlvl_feat, "mlvl_feat")
            tr = self.trainer()
            out = HeadFunction.apply(tr, mlvl_feat, img_metas, reference_joints, *self._trainer_params)
            if self.dims.parametric:
                return {"all_coords_preds": out[0], "pred_pose": out[1].view(-1, 16, 3), "pred_shape": out[2]}
            return {"all_coords_preds": out}
        return self._forward_eval(mlvl_feat, img_metas, reference_joints, **kwargs)

    @torch.no_grad()
    def _forward_eval(self, mlvl_feat, img_metas, reference_joints, **kwargs):
        d = self.dims
        _require_cuda(mlvl_feat, "mlvl_feat")
        dev = mlvl_feat.device
        views = np.asarray(img_metas["cam_view_num"]).astype(np.int32).reshape(-1)
        B = len(img_metas["master_id"])
        assert int(np.sum(np.asarray(img_metas["master_id"]))) == 0, "only support master_id is 0"
        assert len(views) == B and int(views.sum()) == mlvl_feat.shape[0]
        assert mlvl_feat.shape[1] == d.in_channels and tuple(mlvl_feat.shape[-2:]) == (d.feat_hw, d.feat_hw)
        inp_w, inp_h = img_metas["inp_img_shape"]   # the reference unpacks (H,W) as (w,h) (ptEmb_head.py:831)
        # mirrors the reference's mutation of img_metas (ptEmb_head.py:833).  The tensor is cached: building it per call
        # is a pageable host->device copy that synchronises the stream and serialises consecutive forwards
        # (measured: 6.3 ms of CPU time per call instead of 1.9 ms, scripts/exp_host_enqueue.py).
        key = (str(dev), float(inp_w), float(inp_h))
        if getattr(self, "_inp_res_key", None) != key and not torch.cuda.is_current_stream_capturing():
            self._inp_res, self._inp_res_key = torch.tensor([inp_w, inp_h], dtype=torch.float32, device=dev), key
        if getattr(self, "_inp_res_key", None) == key:
            img_metas["inp_res"] = self._inp_res
        feat = mlvl_feat.contiguous().float()
        intr = img_metas["cam_intr"].to(dev, torch.float32).contiguous()
        extr = img_metas["cam_extr"].to(dev, torch.float32).contiguous()
        refj = reference_joints.to(dev, torch.float32).contiguous()
        assert refj.shape == (B, 21, 3)
        lib = nat.load()
        pw = self.packed(dev)
        cd = nat.make_dims(d, MAX_VIEWS)
        NV = int(feat.shape[0])
        need = lib.poem_workspace_bytes(C.byref(cd), B, NV)
        if need == 0:
            raise nat.PoemError("unsupported dimensions: " + lib.poem_last_error().decode())
        ws_ptr, ws_bytes = self.workspace(need, dev)
        out = torch.empty(d.n_blocks, B, d.n_query, 3, dtype=torch.float32, device=dev)
        inp = nat.PoemInputs(B, NV, views.ctypes.data, feat.data_ptr(), intr.data_ptr(), extr.data_ptr(),
                             refj.data_ptr(), float(inp_w), float(inp_h))
        stream = torch.cuda.current_stream(dev).cuda_stream
        nbr = None
        if self.debug_export_neighbours and d.n_blocks > 1:      # test hook: the 32-NN sets the kernels used
            nbr = torch.zeros(d.n_blocks - 1, 2, B, d.n_query, 32, dtype=torch.int32, device=dev)
            lib.poem_debug_export_neighbours(nbr.data_ptr(), nbr.numel())
        try:
            if not d.parametric:
                nat.check(lib.poem_head_forward(C.byref(cd), C.byref(pw.struct), C.byref(inp), out.data_ptr(), None,
                                                ws_ptr, ws_bytes, stream))
                return {"all_coords_preds": out}
            # medium_MANO: the last block's joints / vertices come from the MANO tail (ptEmb_head.py:950-963)
            pm = self.packed_mano(dev)
            pose = torch.empty(B, 16, 3, dtype=torch.float32, device=dev)
            shape = torch.empty(B, 10, dtype=torch.float32, device=dev)
            nat.check(lib.poem_head_forward_parametric(C.byref(cd), C.byref(pw.struct), C.byref(pm.struct),
                                                       C.byref(inp), out.data_ptr(), pose.data_ptr(), shape.data_ptr(),
                                                       ws_ptr, ws_bytes, stream))
            return {"all_coords_preds": out, "pred_pose": pose, "pred_shape": shape}
        finally:
            if nbr is not None:
                lib.poem_debug_export_neighbours(None, 0)
                self.last_neighbours = nbr

    @torch.no_grad()
    def forward_host(self, mlvl_feat, img_metas, reference_joints, out=None):
        """Same path through `poem_head_forward_host`: HOST (pinned) inputs, HOST output; H2D/D2H inside the call."""
        d = self.dims
        dev = next(self.parameters()).device
        views = np.asarray(img_metas["cam_view_num"]).astype(np.int32).reshape(-1)
        B, NV = len(views), int(mlvl_feat.shape[0])
        lib = nat.load()
        pw = self.packed(dev)
        cd = nat.make_dims(d, MAX_VIEWS)
        ws_ptr, ws_bytes = self.workspace(lib.poem_workspace_bytes(C.byref(cd), B, NV), dev)
        # staging is private to the host entry point (its copy stream writes it while earlier calls still compute)
        st_need = lib.poem_staging_bytes(C.byref(cd), B, NV)
        if getattr(self, "_stage", None) is None or self._stage.numel() < st_need + 1024 or self._stage.device != torch.device(dev):
            if getattr(self, "_stage", None) is not None:
                torch.cuda.synchronize(dev)    # the library's copy stream may still be writing the old buffer
            self._stage = torch.empty(st_need + 1024, dtype=torch.uint8, device=dev)
        st_off = (-self._stage.data_ptr()) % 1024
        if out is None:
            out = torch.empty(d.n_blocks, B, d.n_query, 3, dtype=torch.float32).pin_memory()
        inp_w, inp_h = img_metas["inp_img_shape"]
        for t in (mlvl_feat, img_metas["cam_intr"], img_metas["cam_extr"], reference_joints):
            assert not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        inp = nat.PoemInputs(B, NV, views.ctypes.data, mlvl_feat.data_ptr(), img_metas["cam_intr"].data_ptr(),
                             img_metas["cam_extr"].data_ptr(), reference_joints.data_ptr(), float(inp_w), float(inp_h))
        stream = torch.cuda.current_stream(dev).cuda_stream
        if d.parametric:      # returns (coords, pred_pose (B,16,3), pred_shape (B,10)), all in pinned host memory
            pm = self.packed_mano(dev)
            if getattr(self, "_host_pose", None) is None or self._host_pose.shape[0] != B:
                self._host_pose = torch.empty(B, 16, 3, dtype=torch.float32).pin_memory()
                self._host_shape = torch.empty(B, 10, dtype=torch.float32).pin_memory()
            nat.check(lib.poem_head_forward_parametric_host(C.byref(cd), C.byref(pw.struct), C.byref(pm.struct),
                                                            C.byref(inp), out.data_ptr(), self._host_pose.data_ptr(),
                                                            self._host_shape.data_ptr(), self._stage.data_ptr() + st_off,
                                                            self._stage.numel() - st_off, ws_ptr, ws_bytes, stream))
            return out, self._host_pose, self._host_shape
        nat.check(lib.poem_head_forward_host(C.byref(cd), C.byref(pw.struct), C.byref(inp), out.data_ptr(),
                                             self._stage.data_ptr() + st_off, self._stage.numel() - st_off, ws_ptr,
                                             ws_bytes, stream))
        return out


class PtEmbedTRv4(_NativeDecoder):
    """Drop-in for the reference transformer: `forward(query_xyz, query_feat, pt_xyz, pt_feats)` ->
    (xyz (NB,B,799,3) normalised, pred_pose None, pred_shape None)."""

    def __init__(self, cfg, mano_params=None):
        if isinstance(cfg, HeadDims):
            dims = cfg
        else:
            g = cfg.get if hasattr(cfg, "get") else (lambda k, dflt=None: getattr(cfg, k, dflt))
            dims = HeadDims(embed_dims=int(g("INPUT_FEAT_DIM")), n_blocks=int(g("N_BLOCKS")),
                            n_heads=int(g("NUM_ATTENTION_HEADS")), n_sample=int(g("BPS_FEAT_DIM")),
                            n_neighbor=int(g("N_NEIGHBOR")), parametric=bool(g("PARAMETRIC_OUTPUT", False)),
                            center_idx=int(g("TRANSFORMER_CENTER_IDX", 9)))
        super().__init__(dims, "transformer.", template_mesh=torch.zeros(dims.n_query, 3), mano_params=mano_params)
        self.name = type(self).__name__

    @torch.no_grad()
    def forward(self, query_xyz, query_feat, pt_xyz, pt_feats):
        d = self.dims
        _require_cuda(query_feat, "query_feat")
        dev = query_feat.device
        B = query_feat.shape[0]
        lib = nat.load()
        pw = self.packed(dev)
        cd = nat.make_dims(d, MAX_VIEWS, run_last_ffn=d.parametric)
        need = lib.poem_transformer_workspace_bytes(C.byref(cd), B)
        ws_ptr, ws_bytes = self.workspace(need, dev)
        args = [t.to(dev, torch.float32).contiguous() for t in (query_xyz, query_feat.expand(B, -1, -1), pt_xyz, pt_feats)]
        out = torch.empty(d.n_blocks, B, d.n_query, 3, dtype=torch.float32, device=dev)
        feats = torch.empty(B, d.n_query, d.embed_dims, dtype=torch.float32, device=dev) if d.parametric else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        nat.check(lib.poem_transformer_forward(C.byref(cd), C.byref(pw.struct), B, args[0].data_ptr(),
                                               args[1].data_ptr(), args[2].data_ptr(), args[3].data_ptr(),
                                               out.data_ptr(), None if feats is None else feats.data_ptr(), ws_ptr,
                                               ws_bytes, stream))
        if not d.parametric:
            return out, None, None
        # pt_metro_transformer.py:194-195: the last block's xyz is overwritten by the (root-centred, metric) MANO output
        pm = self.packed_mano(dev)
        pose = torch.empty(B, 48, dtype=torch.float32, device=dev)
        shape = torch.empty(B, 10, dtype=torch.float32, device=dev)
        nat.check(lib.poem_parametric_tail(C.byref(cd), C.byref(pm.struct), B, feats.data_ptr(), None,
                                           out[-1].data_ptr(), pose.data_ptr(), shape.data_ptr(), ws_ptr, ws_bytes,
                                           stream))
        return out, pose, shape


def register_into(head_registry=None, transformer_registry=None):
    """Re-register the native classes under the reference's names (`force=True`, builder.py:237-239,252-304)."""
    if head_registry is not None:
        head_registry.register_module(name="POEM_Generalized_Head", force=True, module=POEM_Generalized_Head)
    if transformer_registry is not None:
        transformer_registry.register_module(name="PtEmbedTRv4", force=True, module=PtEmbedTRv4)
