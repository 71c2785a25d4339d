"""ctypes binding of the C-ABI in include/poem_b200.h (libpoem_b200.so, built in-tree by `build()`).

There is no CPU fallback: `load()` raises if the library is missing or cannot be loaded.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("POEM_B200_LIB", os.path.join(CSRC, "libpoem_b200.so"))   # override: kernel experiments
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
SOURCES = ["poem_b200.cu"]
HEADERS = ["common.cuh", "gemm.cuh", "conv3x3.cuh", "mha.cuh", "simt.cuh", "vecattn.cuh", "hrnet.cuh", "mano.cuh",
           "sample_merge.cuh", "qchain.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]

POEM_MAX_BLOCKS = 8


def op16_dtype():
    """torch dtype of the library's 16-bit operand format `poem_op16` (IEEE fp16, csrc/common.cuh)."""
    import torch
    return torch.float16


def to_op16(t):
    """Round a weight / activation tensor to the operand format; raises on overflow (|x| > 65504 would become inf)."""
    import torch
    t = t.detach()
    if t.numel() and float(t.abs().max()) > 65504.0:
        raise ValueError("value outside the fp16 operand range (|x| > 65504); rescale the checkpoint")
    return t.to(torch.float16)


def _nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if p and os.path.exists(p):
            return p
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, "poem_b200.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> csrc/libpoem_b200.so (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


# ------------------------------------------------------------------------------------------ structs
class PoemDims(C.Structure):
    _fields_ = [("embed_dims", C.c_int32), ("in_channels", C.c_int32), ("n_sample", C.c_int32),
                ("n_query", C.c_int32), ("n_blocks", C.c_int32), ("n_heads", C.c_int32), ("n_neighbor", C.c_int32),
                ("feat_h", C.c_int32), ("feat_w", C.c_int32), ("center_idx", C.c_int32), ("radius", C.c_float),
                ("max_views", C.c_int32), ("run_last_ffn", C.c_int32)]


class PoemLinear(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p)]


class PoemVecAttn(C.Structure):
    _fields_ = [("wd1", C.c_void_p), ("bd1", C.c_void_p), ("delta2", PoemLinear), ("gamma1_delta2", PoemLinear),
                ("gamma2", PoemLinear), ("fc2", PoemLinear)]


class PoemBlock(C.Structure):
    _fields_ = [("embedding", PoemLinear), ("pt_proj", PoemLinear), ("q1", PoemLinear), ("o1", PoemLinear),
                ("ln1_g", C.c_void_p), ("ln1_b", C.c_void_p), ("q2", PoemLinear), ("o2", PoemLinear),
                ("ln2_g", C.c_void_p), ("ln2_b", C.c_void_p), ("self_qkv", PoemLinear), ("self_attn", PoemVecAttn),
                ("cross_q", PoemLinear), ("cross_attn", PoemVecAttn), ("reg1", PoemLinear), ("reg2_w", C.c_void_p),
                ("reg2_b", C.c_void_p), ("ffn1", PoemLinear), ("ffn2", PoemLinear), ("ln3_g", C.c_void_p),
                ("ln3_b", C.c_void_p)]


class PoemWeights(C.Structure):
    _fields_ = [("input_proj", PoemLinear), ("pos_table", C.c_void_p), ("merge0a", PoemLinear),
                ("merge0b", PoemLinear), ("merge1a", PoemLinear), ("merge1b", PoemLinear), ("query_embed", C.c_void_p),
                ("bps", C.c_void_p), ("anchor_xyz", C.c_void_p), ("anchor_idx", C.c_void_p),
                ("template_xyz", C.c_void_p), ("bps_perm", C.c_void_p), ("bps_chunk_box", C.c_void_p),
                ("blocks", PoemBlock * POEM_MAX_BLOCKS)]


class PoemManoTail(C.Structure):
    _fields_ = [("flat_w", C.c_void_p), ("flat_b", C.c_void_p), ("lin_w", C.c_void_p), ("lin_b", C.c_void_p),
                ("v_template", C.c_void_p), ("shapedirs", C.c_void_p), ("posedirs", C.c_void_p),
                ("j_regressor", C.c_void_p), ("skin_weights", C.c_void_p)]


POEM_HR_MAX_MODULES = 4


class PoemHRModule(C.Structure):
    _fields_ = [("branch", ((PoemLinear * 2) * 4) * 4), ("fuse", ((PoemLinear * 3) * 4) * 4)]


class PoemHRStage4(C.Structure):
    _fields_ = [("n_modules", C.c_int32), ("channels", C.c_int32 * 4), ("modules", PoemHRModule * POEM_HR_MAX_MODULES)]


class PoemBottleneck(C.Structure):
    _fields_ = [("c1", PoemLinear), ("c2", PoemLinear), ("c3", PoemLinear), ("ds", PoemLinear)]


class PoemHRNet(C.Structure):
    _fields_ = [("stem1_w", C.c_void_p), ("stem1_b", C.c_void_p), ("stem2", PoemLinear),
                ("layer1", PoemBottleneck * 4), ("trans1", PoemLinear * 2), ("trans2", PoemLinear),
                ("trans3", PoemLinear), ("channels", C.c_int32 * 4), ("stage2", PoemHRModule * 1),
                ("stage3", PoemHRModule * 4), ("stage4", PoemHRModule * 3)]


class PoemFeatDecode(C.Structure):
    _fields_ = [("delayer", PoemLinear * 3), ("feat_in", PoemLinear), ("out_channels", C.c_int32)]


class PoemUVDecode(C.Structure):
    _fields_ = [("delayer", PoemLinear * 3), ("out_w", C.c_void_p), ("out_b", C.c_void_p), ("n_joints", C.c_int32)]


class PoemInputs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("n_images", C.c_int32), ("view_counts", C.c_void_p), ("mlvl_feat", C.c_void_p),
                ("cam_intr", C.c_void_p), ("cam_extr", C.c_void_p), ("reference_joints", C.c_void_p),
                ("inp_img_w", C.c_float), ("inp_img_h", C.c_float)]


# every symbol include/poem_b200.h declares
EXPORTS = ["poem_abi_version", "poem_last_error", "poem_kernel_launches", "poem_profile_enable",
           "poem_profile_summary", "poem_debug_force_unfused", "poem_debug_conv_mode", "poem_debug_export_neighbours", "poem_debug_export_pt_feats", "poem_hrnet_stage4_workspace_bytes",
           "poem_hrnet_stage4_forward", "poem_conv_nhwc", "poem_hrnet_workspace_bytes", "poem_hrnet_forward", "poem_image_features_workspace_bytes", "poem_image_features", "poem_triangulate_dlt", "poem_pa_metrics", "poem_workspace_bytes", "poem_head_forward", "poem_staging_bytes",
           "poem_head_forward_host", "poem_transformer_workspace_bytes", "poem_transformer_forward", "poem_linear",
           "poem_mha", "poem_knn32", "poem_knn32_bps", "poem_project_sample", "poem_sample_taps", "poem_vector_attention",
           "poem_vector_attention_workspace_bytes", "poem_layernorm", "poem_parametric_tail_workspace_bytes",
           "poem_parametric_tail", "poem_head_forward_parametric", "poem_head_forward_parametric_host"]

_lib = None


class PoemError(RuntimeError):
    pass


def load():
    """Load libpoem_b200.so; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PoemError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    lib.poem_abi_version.restype = i
    lib.poem_last_error.restype = C.c_char_p
    lib.poem_kernel_launches.restype = C.c_longlong
    lib.poem_profile_enable.argtypes = [i]
    lib.poem_profile_enable.restype = None
    lib.poem_debug_force_unfused.argtypes = [i]
    lib.poem_debug_force_unfused.restype = None
    lib.poem_debug_export_neighbours.argtypes = [vp, sz]
    lib.poem_debug_export_neighbours.restype = None
    lib.poem_debug_export_pt_feats.argtypes = [vp, sz]
    lib.poem_debug_export_pt_feats.restype = None
    lib.poem_profile_summary.restype = sz
    lib.poem_profile_summary.argtypes = [C.c_char_p, sz]
    lib.poem_workspace_bytes.restype = sz
    lib.poem_workspace_bytes.argtypes = [C.POINTER(PoemDims), i, i]
    lib.poem_staging_bytes.restype = sz
    lib.poem_staging_bytes.argtypes = [C.POINTER(PoemDims), i, i]
    lib.poem_head_forward.restype = i
    lib.poem_head_forward.argtypes = [C.POINTER(PoemDims), C.POINTER(PoemWeights), C.POINTER(PoemInputs), vp, vp, vp,
                                      sz, vp]
    lib.poem_head_forward_host.restype = i
    lib.poem_head_forward_host.argtypes = [C.POINTER(PoemDims), C.POINTER(PoemWeights), C.POINTER(PoemInputs), vp, vp,
                                           sz, vp, sz, vp]
    lib.poem_parametric_tail_workspace_bytes.restype = sz
    lib.poem_parametric_tail_workspace_bytes.argtypes = [C.POINTER(PoemDims), i]
    lib.poem_parametric_tail.restype = i
    lib.poem_parametric_tail.argtypes = [C.POINTER(PoemDims), C.POINTER(PoemManoTail), i, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.poem_head_forward_parametric.restype = i
    lib.poem_head_forward_parametric.argtypes = [C.POINTER(PoemDims), C.POINTER(PoemWeights), C.POINTER(PoemManoTail),
                                                 C.POINTER(PoemInputs), vp, vp, vp, vp, sz, vp]
    lib.poem_head_forward_parametric_host.restype = i
    lib.poem_head_forward_parametric_host.argtypes = [C.POINTER(PoemDims), C.POINTER(PoemWeights),
                                                      C.POINTER(PoemManoTail), C.POINTER(PoemInputs), vp, vp, vp, vp, sz,
                                                      vp, sz, vp]
    lib.poem_transformer_workspace_bytes.restype = sz
    lib.poem_transformer_workspace_bytes.argtypes = [C.POINTER(PoemDims), i]
    lib.poem_transformer_forward.restype = i
    lib.poem_transformer_forward.argtypes = [C.POINTER(PoemDims), C.POINTER(PoemWeights), i, vp, vp, vp, vp, vp, vp,
                                             vp, sz, vp]
    lib.poem_linear.restype = i
    lib.poem_linear.argtypes = [vp, i, vp, i, vp, i, i, i, i, vp, i, vp, i, vp, i, vp]
    lib.poem_mha.restype = i
    lib.poem_mha.argtypes = [vp, i, vp, i, vp, i, vp, i, i, i, i, i, i, vp]
    lib.poem_knn32_bps.restype = i
    lib.poem_knn32_bps.argtypes = [vp, vp, vp, vp, vp, i, i, i, vp]
    lib.poem_knn32.restype = i
    lib.poem_knn32.argtypes = [vp, vp, vp, i, i, i, vp]
    lib.poem_project_sample.restype = i
    lib.poem_project_sample.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, f, f, vp, vp, sz, vp]
    lib.poem_sample_taps.restype = i
    lib.poem_sample_taps.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, f, f, vp, vp, vp, sz, vp]
    lib.poem_vector_attention_workspace_bytes.restype = sz
    lib.poem_vector_attention_workspace_bytes.argtypes = [i, i, i]
    lib.poem_vector_attention.restype = i
    lib.poem_vector_attention.argtypes = [C.POINTER(PoemVecAttn), vp, i, vp, i, vp, i, vp, vp, vp, vp, vp, i, i, i, i,
                                          vp, vp, sz, vp]
    lib.poem_hrnet_stage4_workspace_bytes.restype = sz
    lib.poem_hrnet_stage4_workspace_bytes.argtypes = [C.POINTER(PoemHRStage4), i, i]
    lib.poem_hrnet_stage4_forward.restype = i
    lib.poem_hrnet_stage4_forward.argtypes = [C.POINTER(PoemHRStage4), i, i, C.POINTER(vp), C.POINTER(vp), vp, sz, vp]
    lib.poem_hrnet_workspace_bytes.restype = sz
    lib.poem_hrnet_workspace_bytes.argtypes = [C.POINTER(PoemHRNet), i, i]
    lib.poem_hrnet_forward.restype = i
    lib.poem_hrnet_forward.argtypes = [C.POINTER(PoemHRNet), i, i, vp, C.POINTER(vp), vp, sz, vp]
    lib.poem_image_features_workspace_bytes.restype = sz
    lib.poem_image_features_workspace_bytes.argtypes = [C.POINTER(PoemHRNet), C.POINTER(PoemFeatDecode),
                                                        C.POINTER(PoemUVDecode), i, i]
    lib.poem_image_features.restype = i
    lib.poem_image_features.argtypes = [C.POINTER(PoemHRNet), C.POINTER(PoemFeatDecode), C.POINTER(PoemUVDecode), i, i, vp,
                                        vp, vp, vp, C.POINTER(vp), vp, sz, vp]
    lib.poem_pa_metrics.restype = i
    lib.poem_pa_metrics.argtypes = [vp, vp, i, i, vp, vp, vp]
    lib.poem_triangulate_dlt.restype = i
    lib.poem_triangulate_dlt.argtypes = [vp, vp, vp, vp, i, i, vp, vp]
    lib.poem_conv_nhwc.restype = i
    lib.poem_conv_nhwc.argtypes = [vp, i, i, i, i, vp, vp, i, i, i, i, vp, vp, i, i, vp]
    lib.poem_debug_conv_mode.restype = None
    lib.poem_debug_conv_mode.argtypes = [i]
    lib.poem_layernorm.restype = i
    lib.poem_layernorm.argtypes = [vp, vp, vp, vp, vp, i, i, vp]
    _lib = lib
    return lib


def profile_summary():
    """dict {"<kernel>:<stage>": {"ms": total, "n": launches}} of the launches since poem_profile_enable(1)."""
    import json
    buf = C.create_string_buffer(1 << 16)
    n = load().poem_profile_summary(buf, len(buf))
    return json.loads(buf.value.decode()) if n else {}


def check(rc):
    if rc != 0:
        raise PoemError(f"libpoem_b200 error {rc}: {load().poem_last_error().decode()}")


def make_dims(d, max_views=10, run_last_ffn=False):
    return PoemDims(d.embed_dims, d.in_channels, d.n_sample, d.n_query, d.n_blocks, d.n_heads, d.n_neighbor,
                    d.feat_hw, d.feat_hw, d.center_idx, d.radius, max_views, int(run_last_ffn))


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())
