"""ctypes binding of include/poem_train.h (libpoem_train.so, built in-tree by `build()`): the device primitives of the
training path.  There is no CPU fallback: `load()` raises if the library is missing, every call raises on a non-zero
return code with the library's message."""
import ctypes as C
import os
import subprocess

from . import _native as _nat

CSRC = _nat.CSRC
LIB_PATH = os.environ.get("POEM_TRAIN_LIB", os.path.join(CSRC, "libpoem_train.so"))
SOURCES = ["poem_train.cu"]
HEADERS = ["common.cuh", "tgemm.cuh", "train_simt.cuh", "mano.cuh", "mano_bwd.cuh"]

_P, _I, _L, _F, _U = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_ulonglong

SIGNATURES = {
    "poem_tr_gemm": [_P, _I, _L, _L, _L, _P, _I, _L, _L, _L, _P, _L, _L, _L, _I, _I, _I, _I, _I, _F, _P, _I, _I, _I, _P, _L, _I, _I, _P],
    "poem_tr_relu": [_P, _L, _P],
    "poem_tr_relu_bwd": [_P, _P, _L, _P],
    "poem_tr_gelu": [_P, _P, _L, _P],
    "poem_tr_gelu_bwd": [_P, _P, _L, _P],
    "poem_tr_round_tf32": [_P, _P, _L, _P],
    "poem_tr_axpy": [_P, _P, _F, _L, _P],
    "poem_tr_affine_rows": [_P, _P, _F, _P, _L, _I, _I, _I, _P],
    "poem_tr_colsum": [_P, _L, _L, _I, _P, _P],
    "poem_tr_rowsum_groups": [_P, _L, _I, _I, _P, _P],
    "poem_tr_sum_batch": [_P, _I, _L, _P, _P],
    "poem_tr_bcast_batch": [_P, _I, _L, _P, _P],
    "poem_tr_layernorm": [_P, _P, _P, _P, _F, _P, _P, _P, _L, _I, _P],
    "poem_tr_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _L, _I, _P],
    "poem_tr_dropout": [_P, _P, _L, _F, _P, _U, _P],
    "poem_tr_softmax_rows": [_P, _L, _I, _F, _P, _F, _P, _U, _P],
    "poem_tr_softmax_rows_bwd": [_P, _P, _L, _I, _F, _F, _P, _U, _P],
    "poem_tr_va_make_idx": [_P, _P, _I, _I, _I, _P, _P],
    "poem_tr_va_rel": [_P, _P, _P, _P, _L, _P, _P],
    "poem_tr_lin3_relu": [_P, _P, _P, _P, _L, _I, _P],
    "poem_tr_lin3_bwd": [_P, _P, _P, _P, _P, _P, _L, _I, _P],
    "poem_tr_va_gather_t": [_P, _P, _P, _P, _P, _L, _I, _P],
    "poem_tr_va_softmax_agg": [_P, _P, _P, _P, _F, _P, _L, _I, _P],
    "poem_tr_va_softmax_agg_bwd": [_P, _P, _P, _P, _P, _F, _P, _L, _I, _P],
    "poem_tr_va_scatter": [_P, _P, _P, _P, _P, _P, _L, _I, _P],
    "poem_tr_va_drel_scatter": [_P, _P, _P, _P, _L, _P],
    "poem_tr_lin_n3": [_P, _P, _P, _P, _P, _L, _I, _P],
    "poem_tr_lin_n3_bwd": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "poem_tr_project": [_P, _P, _P, _P, _P, _I, _I, _F, _F, _P, _P],
    "poem_tr_sample": [_P, _P, _P, _I, _I, _I, _I, _P],
    "poem_tr_sample_bwd": [_P, _P, _P, _I, _I, _I, _I, _P],
    "poem_tr_merge_agg": [_P, _P, _P, _I, _I, _I, _P, _P],
    "poem_tr_merge_agg_bwd": [_P, _P, _P, _P, _I, _I, _I, _P, _P],
    "poem_tr_merge_out": [_P, _P, _P, _P, _I, _I, _I, _P, _P],
    "poem_tr_merge_out_bwd": [_P, _P, _P, _I, _I, _I, _P, _P, _P],
    "poem_tr_sumsq": [_P, _L, _P, _P],
    "poem_tr_clip_scale": [_P, _L, _P, _F, _P],
    "poem_tr_seg_sumsq": [_P, _P, _P, _I, _P, _P],
    "poem_tr_seg_clip": [_P, _P, _P, _I, _P, _F, _P],
    "poem_tr_adam": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _P],
    "poem_tr_coord_loss": [_P, _P, _P, _I, _I, _I, _I, _F, _F, _P, _P, _P],
    "poem_tr_compute_loss": [_P] * 8 + [_I, _I] + [_F] * 5 + [_P] * 4 + [_F, _F] + [_P] * 4 + [_P],
    "poem_tr_mano_tail": [_P] * 11 + [_I, _I, _I, _I] + [_P] * 4 + [_P],
    "poem_tr_mano_tail_bwd": [_P] * 9 + [_I, _I, _I, _I] + [_P] * 10 + [_P],
}
EXPORTS = ["poem_tr_abi_version", "poem_tr_last_error", "poem_tr_kernel_launches"] + list(SIGNATURES)

_lib = None


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(_nat.INCLUDE, "poem_train.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> csrc/libpoem_train.so (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nat._nvcc()] + _nat.NVCC_FLAGS + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback for the training path)")
    lib = C.CDLL(LIB_PATH)
    lib.poem_tr_abi_version.restype = C.c_int
    lib.poem_tr_last_error.restype = C.c_char_p
    lib.poem_tr_kernel_launches.restype = C.c_longlong
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    return lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_prof = None     # list of (label, start event, end event) while profiling


def profile(on=True):
    """Per-call CUDA-event timing of the primitives (diagnostics: serialises nothing, adds two event records per call)."""
    global _prof
    _prof = [] if on else None


def profile_summary():
    """{label: (calls, total ms)} of the calls since profile(True); synchronises."""
    import torch
    torch.cuda.synchronize()
    out = {}
    for label, a, b in _prof or []:
        n, t = out.get(label, (0, 0.0))
        out[label] = (n + 1, t + a.elapsed_time(b))
    return out


def call(name, *args, label=None):
    """Call a primitive; torch tensors are passed as device pointers, the current torch stream is appended."""
    import torch
    lib = load()
    conv = [(_ptr(a) if (a is None or isinstance(a, torch.Tensor)) else a) for a in args]
    if _prof is not None:
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
    rc = getattr(lib, name)(*conv, _stream())
    if _prof is not None:
        eb.record()
        _prof.append((label or name.replace("poem_tr_", ""), ea, eb))
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.poem_tr_last_error().decode()}")


def gemm(A, B, Cout, M, N, K, *, a_mn=False, b_mn=False, lda=None, ldb=None, ldc=None, batch=(1, 1),
         a_strides=(0, 0), b_strides=(0, 0), c_strides=(0, 0), alpha=1.0, bias=None, bias_on_m=False, accumulate=False,
         relu=False, relu_mask=None, round_ops=3, round_out=False):
    """C (+)= alpha * op(A) op(B)^T (+ bias), see include/poem_train.h.  Default pitches: dense row-major operands."""
    lda = lda if lda is not None else (M if a_mn else K)
    ldb = ldb if ldb is not None else (N if b_mn else K)
    ldc = ldc if ldc is not None else N
    call("poem_tr_gemm", A, int(a_mn), lda, a_strides[0], a_strides[1], B, int(b_mn), ldb, b_strides[0], b_strides[1],
         Cout, ldc, c_strides[0], c_strides[1], M, N, K, batch[0], batch[1], float(alpha), bias, int(bias_on_m),
         int(accumulate), int(relu), relu_mask, (relu_mask.shape[-1] if relu_mask is not None else 0), int(round_ops), int(round_out),
         label=None if _prof is None else f"gemm {'T' if a_mn else 'N'}{'T' if b_mn else 'N'} {M}x{N}x{K} b{batch[0] * batch[1]}")
