"""CPU-side checks: C-ABI surface, weight packing/folding, module boundary, sharding (gloo, world_size 2)."""
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import poem_oracle as orc
from poem_v2_b200 import _native as nat
from poem_v2_b200 import pack, shard, synth
from poem_v2_b200.config import release_dims, dims_from_cfg
from poem_v2_b200.head import POEM_Generalized_Head, PtEmbedTRv4

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    nat.build()
    header = open(os.path.join(ROOT, "include", "poem_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(poem_[a-z0-9_]+)\s*\(", header))
    assert declared == set(nat.EXPORTS), declared ^ set(nat.EXPORTS)
    lib = nat.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.poem_abi_version() == 2


def test_train_library_exports_every_declared_symbol():
    """include/poem_train.h <-> libpoem_train.so <-> the ctypes table (training-path primitives)."""
    from poem_v2_b200 import _train_native as tnat
    tnat.build()
    header = open(os.path.join(ROOT, "include", "poem_train.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(poem_tr_[a-z0-9_]+)\s*\(", header))
    assert declared == set(tnat.EXPORTS), declared ^ set(tnat.EXPORTS)
    lib = tnat.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.poem_tr_abi_version() == 1
    # argument counts of the ctypes table against the C declarations (the stream is the last parameter of every primitive)
    for name, args in tnat.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", header, flags=re.S)
        assert m, name
        assert len([a for a in m.group(1).split(",") if a.strip()]) == len(args), name


def test_workspace_query_and_dim_validation_need_no_gpu():
    import ctypes as C
    lib = nat.load()
    d = nat.make_dims(release_dims("medium"))
    n = lib.poem_workspace_bytes(C.byref(d), 32, 256)
    assert 2 << 30 < n < 16 << 30
    assert lib.poem_transformer_workspace_bytes(C.byref(d), 32) < n
    bad = nat.make_dims(release_dims("medium"))
    bad.embed_dims = 200
    assert lib.poem_workspace_bytes(C.byref(bad), 1, 1) == 0
    assert b"embed_dims" in lib.poem_last_error()


def test_image_half_workspace_queries_need_no_gpu():
    """Workspace planning of the image half is host arithmetic: compact storage (48 / 80 / 160 / 320 channels per
    pixel) sizes it, the heatmap branch adds its concat buffers, bad arguments return 0."""
    import ctypes as C
    lib = nat.load()
    net, fd, uv = nat.PoemHRNet(), nat.PoemFeatDecode(), nat.PoemUVDecode()
    for i, c in enumerate((40, 80, 160, 320)):
        net.channels[i] = c
    fd.out_channels, uv.n_joints = 160, 21
    n256 = lib.poem_hrnet_workspace_bytes(C.byref(net), 256, 256)
    assert 3.0e9 < n256 < 4.5e9
    assert abs(lib.poem_hrnet_workspace_bytes(C.byref(net), 128, 256) * 2 - n256) < 1 << 20      # linear in the image count
    f = lib.poem_image_features_workspace_bytes(C.byref(net), C.byref(fd), None, 256, 256)
    fu = lib.poem_image_features_workspace_bytes(C.byref(net), C.byref(fd), C.byref(uv), 256, 256)
    assert n256 < f < fu < n256 + (2 << 30)
    assert lib.poem_hrnet_workspace_bytes(C.byref(net), 0, 256) == 0
    assert lib.poem_hrnet_workspace_bytes(C.byref(net), 4, 250) == 0                            # not a multiple of 32
    st = nat.PoemHRStage4()
    st.n_modules = 3
    for i, c in enumerate((40, 80, 160, 320)):
        st.channels[i] = c
    s4 = lib.poem_hrnet_stage4_workspace_bytes(C.byref(st), 256, 64)
    assert 0 < s4 < n256


def test_sine_table_matches_oracle():
    for n, f in [(1, 64), (3, 128), (8, 256)]:
        a = pack.sine_pos_3d(n, 16, 16, f)
        b = orc.sine_pos_3d(n, 16, 16, f)
        assert torch.allclose(a, b, atol=1e-6)


def test_packed_folds_are_algebraically_exact():
    """Folded projections (done in fp64) must equal the unfolded reference chain."""
    dims = release_dims("small")
    sd = synth.make_state_dict(dims, 3)
    bps, a_xyz, a_idx = synth.load_assets()
    pw = pack.PackedWeights(sd, dims, "cpu", synth.standin_template(), bps, a_xyz, a_idx, max_views=4)
    by_ptr = {t.data_ptr(): t for t in pw._keep}
    blk = pw.struct.blocks[1]
    W = by_ptr[blk.pt_proj.w].float()
    b = by_ptr[blk.pt_proj.b]
    D = dims.embed_dims
    g = torch.Generator().manual_seed(0)
    x = torch.randn(50, D, generator=g)
    p = "transformer.pt_metro_encoder.1."
    ke = F.linear(x, sd[p + "embedding.weight"], sd[p + "embedding.bias"])
    c = p + "encoder.vec_attn.query_cross_attn."
    xc = F.linear(ke, sd[c + "fc1.weight"], sd[c + "fc1.bias"])
    want = torch.cat([
        F.linear(ke, sd[p + "encoder.attn.self.key.weight"], sd[p + "encoder.attn.self.key.bias"]),
        F.linear(ke, sd[p + "encoder.cross_attn.self.key.weight"], sd[p + "encoder.cross_attn.self.key.bias"]),
        F.linear(F.linear(xc, sd[c + "w_ks.weight"]), sd[c + "fc_gamma.0.weight"]),     # kt = W_g1 k
        F.linear(xc, sd[c + "w_vs.weight"]),
        F.linear(ke, sd[p + "encoder.attn.self.value.weight"], sd[p + "encoder.attn.self.value.bias"]),
        F.linear(ke, sd[p + "encoder.cross_attn.self.value.weight"], sd[p + "encoder.cross_attn.self.value.bias"]),
    ], dim=1)
    got = x @ W.t() + b
    assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item()     # bf16 rounding of the folded matrix
    s = p + "encoder.vec_attn.query_self_attn."
    Wq = by_ptr[blk.self_qkv.w].float()
    xs = F.linear(x, sd[s + "fc1.weight"], sd[s + "fc1.bias"])
    g1w, g1b, d2b = sd[s + "fc_gamma.0.weight"], sd[s + "fc_gamma.0.bias"], sd[s + "fc_delta.2.bias"]
    want = torch.cat([F.linear(F.linear(xs, sd[s + "w_qs.weight"]), g1w) + g1w @ d2b + g1b,   # qt
                      F.linear(F.linear(xs, sd[s + "w_ks.weight"]), g1w),                     # kt
                      F.linear(xs, sd[s + "w_vs.weight"])], dim=1)
    got = x @ Wq.t() + by_ptr[blk.self_qkv.b]
    assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item()
    # positional table rows: N=1 -> row 0, N=2 -> rows 1,2, N=3 -> rows 3..5
    tab = by_ptr[pw.struct.pos_table]
    assert tab.shape == (10, 256, D)
    s3 = orc.sine_pos_3d(3, 16, 16, dims.pos_feats)
    want3 = F.conv2d(s3, sd["adapt_pos3d.weight"], sd["adapt_pos3d.bias"])      # (3, D, 16, 16)
    assert torch.allclose(tab[3:6], want3.flatten(2).transpose(1, 2), atol=1e-4)


def _reference_like_state_dict(dims, seed):
    sd = synth.make_state_dict(dims, seed)
    # dead keys the reference checkpoint also carries: must be accepted by a strict load
    sd["transformer.pt_metro_encoder.0.embeddings.word_embeddings.weight"] = torch.zeros(8, dims.embed_dims)
    sd["transformer.pt_metro_encoder.2.pooler.dense.weight"] = torch.zeros(dims.embed_dims, dims.embed_dims)
    sd["center_shift_layer.0.weight"] = torch.zeros(799, 799)
    sd["reg_branches.1.2.bias"] = torch.zeros(3)
    return sd


def test_module_keeps_reference_state_dict_keys():
    dims = release_dims("small")
    head = POEM_Generalized_Head(dims, template_mesh=synth.standin_template())
    assert head.num_preds == 3
    sd = _reference_like_state_dict(dims, 1)
    head.load_state_dict(sd, strict=True)
    live = synth.live_param_shapes(dims)
    own = head.state_dict()
    assert set(own) == set(live)
    for k in live:
        assert torch.equal(own[k], sd[k])
    # a missing LIVE key is still an error
    bad = dict(sd)
    del bad["input_proj.weight"]
    with pytest.raises(RuntimeError):
        POEM_Generalized_Head(dims, template_mesh=synth.standin_template()).load_state_dict(bad, strict=True)
    # as a sub-module of a model shell the keys carry the 'ptEmb_head.' prefix (reference POEM.py:114)
    shell = torch.nn.Module()
    shell.ptEmb_head = POEM_Generalized_Head(dims, template_mesh=synth.standin_template())
    shell.load_state_dict({"ptEmb_head." + k: v for k, v in sd.items()}, strict=True)
    tr = PtEmbedTRv4(dims)
    tr.load_state_dict({k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}, strict=True)


def test_parametric_head_host_side():
    """medium_MANO: reference key names, MANO parameters from the constructor or from the
    `mano_layer.th_*` buffers a reference checkpoint carries, zero-pose template = the oracle's MANO forward."""
    dims = release_dims("medium_MANO")
    assert dims.parametric and dims.embed_dims == 256
    mano = synth.synthetic_mano(11)
    sd = synth.make_state_dict(dims, 2)
    last = f"transformer.pt_metro_encoder.{dims.n_blocks - 1}."
    assert sd[last + "flat_verts.weight"].shape == (1, 799) and sd[last + "mano_linear.weight"].shape == (106, 256)
    head = POEM_Generalized_Head(dims, mano_params=mano)
    head.load_state_dict(sd, strict=True)
    assert head.parametric_output
    v, j = orc.mano_forward(mano, torch.zeros(1, 48), torch.zeros(1, 10), dims.center_idx)
    want = torch.cat([j, v], dim=1)[0]
    assert (pack.mano_zero_pose_template(mano, dims.center_idx) - want).abs().max().item() <= 1e-7
    # a reference checkpoint: MANO buffers ride along under the last block's mano_layer
    ck = dict(sd)
    shapes = {"v_template": (1, 778, 3), "shapedirs": (778, 3, 10), "posedirs": (778, 3, 135),
              "J_regressor": (16, 778), "weights": (778, 16)}
    for k, shp in shapes.items():
        ck[last + "mano_layer.th_" + k] = mano[k].reshape(shp)
    ck[last + "mano_layer.th_faces"] = torch.zeros(1538, 3, dtype=torch.long)
    h2 = POEM_Generalized_Head(dims)
    h2.load_state_dict(ck, strict=True)
    for k in shapes:
        assert torch.equal(h2._mano[k].reshape(-1), mano[k].reshape(-1))
    # without parameters the parametric head fails loudly when it has to pack them
    h3 = POEM_Generalized_Head(dims, template_mesh=synth.standin_template())
    h3.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError, match="MANO parameters unavailable"):
        h3.packed_mano("cpu")
    with pytest.raises(KeyError):
        POEM_Generalized_Head(dims, mano_params={"v_template": mano["v_template"]})
    # C-ABI: workspace query of the stage-level entry point is host arithmetic
    import ctypes as C
    lib = nat.load()
    d = nat.make_dims(dims)
    assert lib.poem_parametric_tail_workspace_bytes(C.byref(d), 32) == 32 * 256 * 4 + 1024
    assert lib.poem_parametric_tail_workspace_bytes(C.byref(d), 0) == 0
    assert lib.poem_parametric_tail(C.byref(d), None, 1, None, None, None, None, None, None, 0, None) == -2


def test_module_reads_reference_config_and_fails_loudly_without_gpu():
    class Node(dict):
        __getattr__ = dict.__getitem__
    cfg = Node(TYPE="POEM_Generalized_Head", EMBED_DIMS=256, POINTS_FEAT_DIM=256, IN_CHANNELS=160, N_SAMPLE=4096,
               NUM_QUERY=799, NUM_PREDS=3, RADIUS_SAMPLE=0.1, CAM_FEAT_MERGE="attn", QUERY_TYPE="KPT",
               TRANSFORMER=Node(TYPE="PtEmbedTRv4", N_BLOCKS=3, INPUT_FEAT_DIM=256, NUM_ATTENTION_HEADS=4,
                                DROPOUT=0.1, BPS_FEAT_DIM=4096, N_NEIGHBOR=32, N_NEIGHBOR_QUERY=32),
               POSITIONAL_ENCODING=Node(TYPE="SinePositionalEncoding3D", NUM_FEATS=128, NORMALIZE=True))
    from dataclasses import replace
    assert dims_from_cfg(cfg) == replace(release_dims("medium"), dropout=0.1)     # DROPOUT only acts in training mode
    head = POEM_Generalized_Head(cfg, template_mesh=synth.standin_template())
    feat, metas, ref_j = synth.make_inputs(head.dims, 1, [2], 1)
    with pytest.raises(nat.PoemError, match="no CPU implementation"):
        head(mlvl_feat=feat, img_metas=metas, reference_joints=ref_j)
    cfg2 = Node(cfg)
    cfg2["QUERY_TYPE"] = "POEM"
    with pytest.raises(AssertionError):                 # reference asserts query_type == "KPT" (ptEmb_head.py:721)
        POEM_Generalized_Head(cfg2)
    with pytest.raises(RuntimeError, match="MANO template unavailable"):
        POEM_Generalized_Head(cfg).packed("cpu")


@pytest.mark.skipif(not os.path.isdir("/root/reference/lib"), reason="reference tree not mounted")
def test_registers_under_reference_names():
    import subprocess
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import ref_shim; ref_shim.install()\n"
        "import lib.models.layers.ptEmb_transformer, lib.models.heads.ptEmb_head\n"
        "from lib.utils.builder import HEAD, TRANSFORMER, build_from_cfg\n"
        "from poem_v2_b200 import head as h, synth\n"
        "h.register_into(HEAD, TRANSFORMER)\n"
        "import yaml; from lib.utils.config import CN\n"
        "cfg = CN(yaml.safe_load(open('config/release/train_medium.yaml')))\n"
        "m = build_from_cfg(cfg.MODEL.HEAD, HEAD, data_preset=cfg.DATA_PRESET)\n"
        "assert type(m) is h.POEM_Generalized_Head and m.dims.embed_dims == 256 and m.num_preds == 3\n"
        "t = build_from_cfg(cfg.MODEL.HEAD.TRANSFORMER, TRANSFORMER)\n"
        "assert type(t) is h.PtEmbedTRv4\n"
        "print('REGISTERED-OK')\n") % (ROOT, os.path.join(ROOT, "oracle"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "REGISTERED-OK" in r.stdout, r.stdout + r.stderr


def test_backbone_keeps_reference_state_dict_keys():
    """HRNetW40 exposes the live keys of reference `HighResolutionNet` (hrnet.py:239-283): 1/4/3 modules in stages
    2/3/4, Bottleneck layer1 with a downsample in block 0 only, transitions that only add the new branch."""
    from poem_v2_b200.hrnet import HRNetW40, backbone_param_shapes
    shapes = backbone_param_shapes()
    m = HRNetW40()
    assert set(m.state_dict()) == set(shapes)
    convs = [k for k, v in shapes.items() if len(v) == 4]
    assert len(convs) == 2 + 13 + 4 + (16 + 2) + 4 * (24 + 3 + 4) + 3 * (32 + 6 + 10)   # 305 convolutions
    assert shapes["layer1.0.downsample.0.weight"] == (256, 64, 1, 1) and "layer1.1.downsample.0.weight" not in shapes
    assert shapes["transition1.1.0.0.weight"] == (80, 256, 3, 3)
    assert shapes["transition3.3.0.0.weight"] == (320, 160, 3, 3) and "transition3.0.0.weight" not in shapes
    assert shapes["stage3.3.fuse_layers.2.0.1.0.weight"] == (160, 40, 3, 3)
    sd = synth.make_backbone_state_dict(0)
    sd["final_layer.0.weight"] = torch.zeros(2048, 1024, 1, 1)   # dead classification head: dropped
    m.load_state_dict(sd, strict=True)
    assert torch.equal(m.state_dict()["stage2.0.branches.1.3.bn2.running_var"], sd["stage2.0.branches.1.3.bn2.running_var"])
    with pytest.raises(nat.PoemError, match="no CPU implementation"):
        m(torch.zeros(1, 3, 256, 256))


def test_model_keeps_reference_checkpoint_keys():
    """The inference model exposes the reference's flat checkpoint key space (POEM.py:100-194): `img_backbone.*`,
    `feat_delayer.*`, `feat_in.*`, `uv_delayer.*`, `uv_out.*`, `ptEmb_head.*`; dead keys are dropped on load."""
    from poem_v2_b200.model import PtEmbedMultiviewStereoV2
    dims = release_dims("small")
    m = PtEmbedMultiviewStereoV2(dims, template_mesh=synth.standin_template())
    sd = synth.make_model_state_dict(dims, 0)
    extra = dict(sd)
    extra["uv_in.conv.weight"] = torch.zeros(80, 21, 1, 1)               # only feeds the unused uv_feat
    extra["mano_layer.th_betas"] = torch.zeros(1, 10)
    extra["img_backbone.classifier.bias"] = torch.zeros(1000)
    extra["ptEmb_head.center_shift_layer.0.weight"] = torch.zeros(799, 799)
    m.load_state_dict({"module." + k: v for k, v in extra.items()}, strict=True)   # DDP prefix as in net_utils.py
    own = m.state_dict()
    assert set(own) == set(sd)
    assert own["uv_delayer.0.conv.weight"].shape == (160, 480, 3, 3)
    assert own["feat_delayer.2.conv.weight"].shape == (320, 160, 3, 3) and own["feat_in.conv.weight"].shape == (160, 320, 1, 1)
    assert torch.equal(own["ptEmb_head.input_proj.weight"], sd["ptEmb_head.input_proj.weight"])
    with pytest.raises(RuntimeError):
        m.load_state_dict({**sd, "bogus.weight": torch.zeros(1)}, strict=True)
    with pytest.raises(NotImplementedError):
        m(synth.make_batch(1, [2], 1), mode="train")


def test_shard_bounds_cover_and_balance():
    for views, world in [([8] * 32, 8), ([8] * 32, 2), ([1, 8, 2, 2, 7, 3], 2), ([4, 4, 4], 4), ([2] * 5, 4)]:
        b = shard.shard_bounds(views, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == len(views)
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(e >= s for s, e in b)
    assert shard.shard_bounds([8] * 32, 8) == [(4 * r, 4 * r + 4) for r in range(8)]
    # skewed view counts: no rank may be left without a sample when B >= world — it would skip the head
    # while the others wait in all_gather
    for views, world in (([1, 1, 1, 10], 4), ([10, 1, 1, 1], 4), ([1, 10, 1, 1, 5, 5, 2], 3), ([1] * 7 + [10], 8),
                         ([10] + [1] * 9, 8)):
        b = shard.shard_bounds(views, world)
        assert b[0][0] == 0 and b[-1][1] == len(views) and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(e > s_ for s_, e in b), (views, world, b)
    assert [e - s_ for s_, e in shard.shard_bounds([3, 1], 4)].count(1) == 2          # B < world: two ranks idle
    assert shard.image_bounds(8, 8) == [(r, r + 1) for r in range(8)]
    assert shard.image_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)] and shard.image_bounds(2, 4)[2:] == [(2, 2), (2, 2)]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dims = release_dims("small")
    views = [2, 1, 3, 2, 2]
    feat, metas, ref_j = synth.make_inputs(dims, len(views), views, 3)
    f, m, r, (s, e) = shard.shard_inputs(feat, metas, ref_j, rank, world)
    assert f.shape[0] == int(np.sum(m["cam_view_num"])) and r.shape[0] == e - s == len(m["master_id"])
    # stand-in for the per-rank decoder call: something that depends on every input of the slice
    local = torch.stack([r[:, :1, :].expand(-1, 799, -1) + f.sum() * 0 + i for i in range(3)])
    bounds = shard.shard_bounds(views, world)
    full = shard.gather_outputs(local, len(views), bounds)
    want = torch.stack([ref_j[:, :1, :].expand(-1, 799, -1) + i for i in range(3)])
    ok = bool(torch.equal(full, want))
    # image-sharded mode (fewer samples than ranks): every rank "extracts" the features of its images, one all_gather
    n_img = feat.shape[0]
    ib = shard.image_bounds(n_img, world)
    i0, i1 = ib[rank]
    gathered = shard.gather_features(feat[i0:i1] * 2.0, n_img, ib)
    ok = ok and bool(torch.equal(gathered, feat * 2.0))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_bench_line_contract_keys():
    """bench.py's JSON line must carry every key of the measurement contract (a stray comment once swallowed
    `roofline`): checked on the source without a GPU, and the CPU reference arm is run for real."""
    import ast
    import json
    import subprocess
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    dicts = [n.value for n in ast.walk(tree) if isinstance(n, ast.Assign) and isinstance(n.value, ast.Dict)
             and any(isinstance(t, ast.Name) and t.id == "line" for t in n.targets)]
    assert len(dicts) == 2                                   # the reference arm's line and ours
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}
    for d in dicts:
        keys = {k.value for k in d.keys if isinstance(k, ast.Constant)}
        assert need <= keys, need - keys
    ours = max(dicts, key=lambda d: len(d.keys))
    keys = {k.value for k in ours.keys if isinstance(k, ast.Constant)}
    assert {"roofline", "path_roofline", "clocks", "kernel_breakdown_ms_per_step"} <= keys
    # the reference arm is CPU-only: one bounded step of the small workload, one JSON line on stdout
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small_v4_b8",
                          "--steps", "1", "--warmup", "1", "--cpu-sample-batch", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def _ragged_samples(seed=0):
    rng = np.random.default_rng(seed)
    samples = []
    for v in (3, 1, 2):
        samples.append({"image": rng.standard_normal((v, 3, 8, 8)).astype(np.float32),
                        "target_joints_3d": rng.standard_normal((v, 21, 3)).astype(np.float32),
                        "target_cam_intr": rng.standard_normal((v, 3, 3)).astype(np.float32),
                        "target_cam_extr": rng.standard_normal((v, 4, 4)).astype(np.float64),
                        "master_joints_3d": rng.standard_normal((1, 21, 3)).astype(np.float32),
                        "image_path": np.array([f"cam{i}.png" for i in range(v)]),
                        "master_id": 0, "sample_idx": [int(rng.integers(100)) for _ in range(v)]})
    return samples


def test_collation_mirror():
    """`collation_random_n_views` (reference lib/utils/collation.py:7-25): ragged view counts -> flat (sum V, ...)
    float32 tensors + cam_view_num; non-numeric fields are listed per sample; a single sample is accepted."""
    from poem_v2_b200.collation import collation_random_n_views
    samples = _ragged_samples()
    out = collation_random_n_views(samples)
    assert out["cam_view_num"].tolist() == [3, 1, 2]
    assert out["image"].shape == (6, 3, 8, 8) and out["image"].dtype == torch.float32
    assert out["target_cam_extr"].dtype == torch.float32            # torch.Tensor(...) casts like the reference
    assert out["master_joints_3d"].shape == (3, 21, 3)
    assert torch.equal(out["target_joints_3d"][3:4], torch.from_numpy(samples[1]["target_joints_3d"]))
    assert out["master_id"] == [0, 0, 0] and len(out["image_path"]) == 3 and list(out["image_path"][0]) == ["cam0.png", "cam1.png", "cam2.png"]
    one = collation_random_n_views(samples[0])
    assert one["cam_view_num"].tolist() == [3] and one["image"].shape == (3, 3, 8, 8)
    if os.path.isdir("/root/reference/lib"):                         # against the reference function itself
        import importlib.util
        import types
        pkg = types.ModuleType("_refcol"); pkg.__path__ = []
        utils = types.ModuleType("_refcol.utils"); utils.__path__ = []
        tr = types.ModuleType("_refcol.utils.transform")
        tr.batch_cam_extr_transf = tr.batch_cam_intr_projection = None
        sys.modules.update({"_refcol": pkg, "_refcol.utils": utils, "_refcol.utils.transform": tr})
        spec = importlib.util.spec_from_file_location("_refcol.utils.collation", "/root/reference/lib/utils/collation.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        want = mod.collation_random_n_views(_ragged_samples())
        assert set(want) == set(out)
        for k, v in want.items():
            if torch.is_tensor(v):
                assert torch.equal(v, out[k]), k
            elif isinstance(v, np.ndarray):
                assert np.array_equal(v, out[k]), k
            else:
                assert all(np.array_equal(a, b) for a, b in zip(v, out[k])), k


def test_poly_exp2_constants_in_the_kernel_source():
    """The FMA-pipe 2^x of the attention kernel (csrc/common.cuh `poly_exp2`): its constants are read from the source and
    the same arithmetic is replayed in float32 — relative error <= 8e-5 on [-30, 0.5], far below the bf16 rounding of the
    probabilities it produces."""
    src = open(os.path.join(ROOT, "poem-v2_b200", "csrc", "common.cuh")).read()
    body = src[src.index("float poly_exp2(float x)"):]
    body = body[:body.index("return __int_as_float")]
    c3, c2, c1, c0 = [np.float32(v) for v in re.findall(r"(\d\.\d+)f", body)][-4:]
    assert "12582912.f" in body and "-126.f" in body
    x = np.linspace(-30, 0.5, 400001).astype(np.float32)
    t = (x + np.float32(12582912.0)).astype(np.float32)
    f = (x - (t - np.float32(12582912.0))).astype(np.float32)
    assert np.abs(f).max() <= 0.5
    p = ((c3 * f + c2) * f + c1) * f + c0
    got = (p.astype(np.float32).view(np.int32) + (t.view(np.int32) << 23)).view(np.float32)
    ref = np.exp2(x.astype(np.float64))
    assert np.max(np.abs(got / ref - 1.0)) <= 8e-5


# ------------------------------------------------------------------------------------------ training path: host logic
def test_train_parameter_buckets_partition_the_flat_buffer():
    """HeadTrainer keeps parameters / gradients as views of one flat buffer; the all-reduce buckets (head stage, one per
    block) must tile it without gaps or overlap and every tensor must start 16-byte aligned (TMA operand)."""
    from poem_v2_b200 import params, synth
    from poem_v2_b200.config import release_dims
    from poem_v2_b200.train import HeadTrainer
    dims = release_dims("small")
    tr = HeadTrainer(dims, synth.make_state_dict(dims, 0), synth.standin_template(), device="cpu")
    spans = sorted(tr.buckets.values())
    assert spans[0][0] == 0 and spans[-1][1] == tr.p_flat.numel()
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert set(tr.buckets) == {"head"} | {str(i) for i in range(dims.n_blocks)}
    assert list(tr.p) == list(params.live_param_shapes(dims))
    for k, v in tr.p.items():
        assert v.data_ptr() % 16 == 0 and tr.g[k].data_ptr() % 16 == 0, k
        assert v.shape == tr.g[k].shape
    sd = synth.make_state_dict(dims, 0)
    assert all(torch.equal(tr.p[k], sd[k].float()) for k in tr.p)
    tr.g_flat.fill_(1.0)
    tr.zero_grad()
    assert float(tr.g_flat.abs().sum()) == 0.0
    feat, metas, ref_j = synth.make_inputs(dims, 1, [2], 1)
    with pytest.raises(nat.PoemError, match="no CPU implementation"):      # no fallback: the kernels are the only path
        tr.forward(feat, metas, ref_j)


def _bucket_allreduce_worker(rank, world, port, q):
    import torch.distributed as dist
    from poem_v2_b200 import synth
    from poem_v2_b200.config import release_dims
    from poem_v2_b200.train import HeadTrainer, TrainStep
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dims = release_dims("small")
    tr = HeadTrainer(dims, synth.make_state_dict(dims, 0), synth.standin_template(), device="cpu")
    step = TrainStep(tr)
    assert step.world == world
    tr.g_flat.fill_(float(rank + 1))
    for name in ("2", "1", "0", "head"):                 # the order the backward finishes them in
        step._bucket_hook(name)
    for h in step._pending:
        h.wait()
    q.put((rank, float(tr.g_flat.min()), float(tr.g_flat.max()), step.allreduce_bytes, tr.g_flat.numel() * 4))
    dist.destroy_process_group()


def test_train_gradient_buckets_average_over_ranks_gloo():
    """world_size 2 on CPU (gloo): every gradient element ends up as the mean over ranks, each byte reduced once."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_bucket_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=120) for _ in procs]
    for p_ in procs:
        p_.join(60)
    for rank, lo, hi, reduced, total in res:
        assert lo == hi == 1.5, (rank, lo, hi)
        assert reduced == total
