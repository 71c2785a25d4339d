"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Stub layer that makes the real reference modules (``/root/reference``) importable in THIS
container on CPU, so that ``oracle/make_golden.py`` can run the unmodified reference decoder
(`POEM_Generalized_Head` + `PtEmbedTRv4`, reference `lib/models/heads/ptEmb_head.py:683-964`,
`lib/models/layers/ptEmb_transformer.py:303-376`) and pin `oracle/poem_oracle.py` against it.

What is stubbed (all absent from this image, none of it arithmetic on the hot path except
where stated):
  * yacs / termcolor / imageio / matplotlib / webdataset / trimesh ... : permissive dummies
  * pytorch3d.ops.knn_points  : squared-L2 + topk(smallest, sorted) restatement (third-party,
    pytorch3d v0.7.2 `knn_points`, not vendored by the reference) — ARITHMETIC, unpinned
  * manotorch.ManoLayer       : returns the seeded stand-in template (MANO assets are licensed
    and absent) — parity with real MANO is unpinned; for the parametric tail (medium_MANO) it runs
    the restated LBS forward on seeded stand-in parameters (`MANO_PARAMS`)
  * pytorch3d.transforms      : rotation_6d_to_matrix / matrix_to_quaternion / quaternion_to_axis_angle
    restated (v0.7.2, third-party) — ARITHMETIC, unpinned; cross-checked against SciPy in tests
  * transformers 5.x -> 4.x   : BertAttention adapter that restores 4.x cross-attention
    semantics (`encoder_hidden_states` => K/V source, no mask)

Nothing here runs on the GPU box: `/root/reference` does not exist there.
"""
import copy
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("POEM_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "models"))


# ----------------------------------------------------------------------------- yacs stand-in
class _CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        for k, v in (init_dict or {}).items():
            self[k] = type(self)(v) if isinstance(v, dict) and not isinstance(v, _CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def defrost(self):
        pass

    def freeze(self):
        pass

    def set_new_allowed(self, _):
        pass

    def is_frozen(self):
        return False

    def merge_from_other_cfg(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge_from_other_cfg(v)
            else:
                self[k] = copy.deepcopy(v)

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self.merge_from_other_cfg(type(self)(yaml.safe_load(f)))

    def merge_from_list(self, lst):
        pass

    def dump(self, *a, **k):
        return str(dict(self))


class _Dummy(types.ModuleType):
    """module whose every attribute is another dummy / a no-op callable class"""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None,
                              "__call__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


class _DummyFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("matplotlib", "open3d", "opendr", "trimesh", "chumpy", "webdataset", "braceexpand",
             "skimage", "sklearn_stub", "pyrender", "neural_renderer", "imageio", "termcolor_stub",
             "manotorch", "pytorch3d", "yacs", "termcolor", "dex_ycb_toolkit", "oikit", "pycocotools",
             "tensorboardX", "roma", "smplx", "lmdb")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.ROOTS and name not in sys.modules:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Dummy(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _knn_points(p1, p2, K=1, return_nn=False, **kw):
    """pytorch3d.ops.knn_points restated: squared L2, K smallest ascending, lowest index wins ties."""
    d = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
    # stable sort => lower index first among equal distances
    order = torch.sort(d, dim=-1, stable=True)
    idx = order.indices[..., :K]
    dist = order.values[..., :K]
    nn_ = None
    if return_nn:
        nn_ = torch.gather(p2[:, None].expand(-1, p1.shape[1], -1, -1), 2, idx[..., None].expand(-1, -1, -1, 3))
    return dist, idx, nn_


_installed = False
MANO_PARAMS = None   # set to a `synth.synthetic_mano()` dict to make the stub ManoLayer a real LBS forward


def install(template_fn=None):
    """Install the stub layer (idempotent) and chdir to the reference root."""
    global _installed
    if _installed:
        return
    assert reference_available(), f"reference not mounted at {REF_ROOT}"
    sys.meta_path.append(_DummyFinder())
    sys.path.insert(0, REF_ROOT)
    os.chdir(REF_ROOT)  # relative 'assets/...', 'config/backbone/...' paths

    # yacs
    import yacs.config as yc  # dummy
    yc.CfgNode = _CfgNode
    # termcolor
    import termcolor
    termcolor.colored = lambda s, *a, **k: s
    termcolor.cprint = lambda *a, **k: print(*a)
    # pytorch3d
    import pytorch3d.ops as p3o
    p3o.knn_points = _knn_points

    def _raise(*a, **k):
        raise NotImplementedError("stubbed pytorch3d op")
    p3o.sample_farthest_points = _raise
    p3o.ball_query = _raise
    import pytorch3d.transforms as p3t
    for n in ["axis_angle_to_matrix", "axis_angle_to_quaternion", "euler_angles_to_matrix", "matrix_to_euler_angles",
              "matrix_to_quaternion", "matrix_to_rotation_6d", "quaternion_to_axis_angle", "quaternion_to_matrix",
              "rotation_6d_to_matrix"]:
        setattr(p3t, n, _raise)
    # the three the parametric tail composes (lib/utils/transform.py:448-466): restated in poem_oracle (third-party
    # arithmetic, pytorch3d v0.7.2, unpinned) — the reference's own `rot6d_to_aa` / `get_parametric_output` run on them
    import poem_oracle as _orc
    p3t.rotation_6d_to_matrix = _orc.rotation_6d_to_matrix
    p3t.matrix_to_quaternion = _orc.matrix_to_quaternion
    p3t.quaternion_to_axis_angle = _orc.quaternion_to_axis_angle

    # manotorch
    import manotorch.manolayer as ml

    class ManoLayer(torch.nn.Module):
        """Without MANO_PARAMS: returns the fixed seeded template.  With MANO_PARAMS (a `synth.synthetic_mano()`
        dict): the restated manotorch forward (`poem_oracle.mano_forward`) on those stand-in parameters, so the
        reference's parametric tail runs end to end."""

        def __init__(self, *a, **k):
            super().__init__()
            t = template_fn() if template_fn is not None else torch.zeros(799, 3)
            self.register_buffer("_joints", t[None, :21].clone(), persistent=False)
            self.register_buffer("_verts", t[None, 21:].clone(), persistent=False)
            self.th_faces = torch.zeros(1538, 3, dtype=torch.long)
            self.th_J_regressor = torch.zeros(16, 778)
            self.center_idx = k.get("center_idx")

        def forward(self, pose, betas=None, **k):
            n = pose.shape[0]
            if MANO_PARAMS is not None:
                if betas is None:
                    betas = torch.zeros(n, 10)
                v, j = _orc.mano_forward(MANO_PARAMS, pose, betas, self.center_idx)
                return types.SimpleNamespace(verts=v, joints=j)
            return types.SimpleNamespace(verts=self._verts.repeat(n, 1, 1), joints=self._joints.repeat(n, 1, 1))
    ml.ManoLayer = ManoLayer

    # bare packages so heavy __init__ files never run
    for pkg in ["lib.models", "lib.models.heads", "lib.models.layers", "lib.models.bricks", "lib.external",
                "lib.external.metro", "lib.viztools"]:
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF_ROOT, *pkg.split("."))]
        sys.modules[pkg] = m
    import lib  # noqa: F401  (real package; lib/__init__ is light)
    for pkg in ["lib.models", "lib.external", "lib.viztools"]:
        setattr(sys.modules["lib"], pkg.split(".")[1], sys.modules[pkg])
    draw = types.ModuleType("lib.viztools.draw")
    for n in ["draw_batch_joint_images", "draw_batch_verts_images", "draw_batch_mesh_images_pred", "plot_hand",
              "draw_batch_hm_images", "draw_2d_skeleton", "plot_image_joints_mask", "plot_image_heatmap_mask"]:
        setattr(draw, n, lambda *a, **k: None)
    sys.modules["lib.viztools.draw"] = draw

    # transformers 5.x -> 4.x semantics
    from transformers.models.bert import modeling_bert as mb
    import transformers.pytorch_utils as pu
    if not hasattr(mb, "apply_chunking_to_forward"):
        mb.apply_chunking_to_forward = pu.apply_chunking_to_forward
    _Base = mb.BertAttention

    class BertAttention4x(_Base):
        """4.x behaviour: passing encoder_hidden_states makes K/V come from them; mask replaced by
        encoder_attention_mask (None here)."""

        def __init__(self, config, position_embedding_type=None, **kw):
            try:
                super().__init__(config, is_cross_attention=True)
            except TypeError:
                super().__init__(config)

        def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, output_attentions=False,
                    **kw):
            out = super().forward(hidden_states, attention_mask=None, encoder_hidden_states=encoder_hidden_states,
                                  encoder_attention_mask=None)
            return out if isinstance(out, tuple) else (out,)
    mb.BertAttention = BertAttention4x
    _orig_init_weights = mb.BertPreTrainedModel.init_weights
    _guard = {"on": False}

    def init_weights(self):
        if _guard["on"]:
            return _orig_init_weights(self)
        _guard["on"] = True
        try:
            if hasattr(self, "post_init"):
                self.post_init()
            else:
                _orig_init_weights(self)
        finally:
            _guard["on"] = False
    mb.BertPreTrainedModel.init_weights = init_weights
    _installed = True


def build_reference_head(size="medium", template_fn=None):
    """Real reference `POEM_Generalized_Head` built from config/release/train_<size>.yaml."""
    install(template_fn)
    import yaml
    from lib.utils.config import CN
    import lib.models.layers.ptEmb_transformer  # noqa: F401  registers PtEmbedTRv4
    import lib.models.heads.ptEmb_head as hd
    from lib.utils.builder import build_from_cfg, HEAD
    with open(os.path.join(REF_ROOT, "config", "release", f"train_{size}.yaml")) as f:
        cfg = CN(yaml.safe_load(f))
    head = build_from_cfg(cfg.MODEL.HEAD, HEAD, data_preset=cfg.DATA_PRESET)
    return head.eval(), cfg
