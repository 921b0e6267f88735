"""ctypes binding of libclipppo_b200.so (the C ABI in include/clipppo_b200.h).

There is no CPU fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CLIPPPO_LIB=<path>: an experimental build of the same ABI (tools/build_variant.sh) instead of the in-tree library - A/B measurements only
LIB_PATH = os.environ.get("CLIPPPO_LIB") or os.path.join(HERE, "libclipppo_b200.so")

# status codes (include/clipppo_b200.h)
OK = 0
ERR_BAD_SHAPE, ERR_BAD_CHANNELS, ERR_BAD_PAD, ERR_NULL = -1, -2, -3, -4
ERR_WORKSPACE, ERR_UNSUPPORTED, ERR_ALIGN, ERR_CUDA, ERR_DIM_MISMATCH = -5, -6, -7, -8, -9

STAGE_NOISE, STAGE_CONTRAST, STAGE_BLUR, STAGE_CUTOUT, STAGE_ALL = 1, 2, 4, 8, 15
IMG_F32, IMG_U8 = 0, 1
VIT_L2NORM, VIT_PRENORMALIZED, VIT_CLS_LAST_BLOCK = 1, 2, 4
EPI_ROWAFFINE_BF16, EPI_ROWAFFINE_GELU_BF16, EPI_RESID_BF16, EPI_RESID_STATS_BF16 = 6, 7, 8, 9
EPI_BIAS_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_RESID_F32, EPI_PATCH_F32, EPI_F32 = 0, 1, 2, 3, 4

# every symbol the header declares; tests/test_abi.py checks the .so exports all of them
SYMBOLS = (
    "clipppo_abi_version", "clipppo_strerror", "clipppo_last_cuda_error", "clipppo_prof_begin", "clipppo_prof_end", "clipppo_prof_bucket",
    "clipppo_disturb_f32", "clipppo_disturb_u8_f32", "clipppo_disturb_nhwc_u8", "clipppo_disturb_ex",
    "clipppo_cosine_loss_fwd", "clipppo_cosine_loss_bwd", "clipppo_gae_f32", "clipppo_ppo_loss_f32",
    "clipppo_vit_create", "clipppo_vit_destroy", "clipppo_vit_workspace_bytes", "clipppo_vit_encode",
    "clipppo_text_create", "clipppo_text_destroy", "clipppo_text_workspace_bytes", "clipppo_text_encode",
    "clipppo_preprocess_bf16", "clipppo_layernorm_bf16", "clipppo_gemm_bf16", "clipppo_gemm_bf16_fused", "clipppo_gemm_bf16_resid_stats", "clipppo_gemm_bf16_fused_parts", "clipppo_rowstats_bf16", "clipppo_attention_bf16",
    "clipppo_attention_causal_bf16",
    "clipppo_nature_workspace_bytes", "clipppo_nature_forward", "clipppo_nature_backward",
)


class NativeLibraryMissing(RuntimeError):
    pass


DISTURB_PHILOX = 1


class DisturbDesc(C.Structure):       # clipppo_disturb_desc (include/clipppo_b200.h)
    _fields_ = [("x", C.c_void_p), ("x_dtype", C.c_int), ("x_strides_host", C.POINTER(C.c_int64)),
                ("noise", C.c_void_p), ("noise_strides_host", C.POINTER(C.c_int64)), ("out", C.c_void_p),
                ("B", C.c_int), ("C", C.c_int), ("H", C.c_int), ("W", C.c_int), ("stages", C.c_int),
                ("noise_sigma", C.c_float), ("contrast", C.c_float), ("k1d_host", C.POINTER(C.c_float)), ("k", C.c_int),
                ("sh", C.c_int), ("sw", C.c_int), ("ph", C.c_int), ("pw", C.c_int),
                ("out_scale", C.c_float), ("flags", C.c_int),
                ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64), ("first_image", C.c_int64)]


class VitConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("width", "layers", "heads", "patch", "image", "out_dim")]


class VitLayer(C.Structure):          # fp32 device pointers, openai/CLIP state-dict layout (include/clipppo_b200.h)
    _fields_ = [(n, C.c_void_p) for n in ("ln1_g", "ln1_b", "w_qkv", "b_qkv", "w_out", "b_out",
                                          "ln2_g", "ln2_b", "w_fc", "b_fc", "w_proj", "b_proj")]


class VitWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("conv1", "class_embedding", "positional_embedding", "ln_pre_g", "ln_pre_b",
                                          "ln_post_g", "ln_post_b", "proj")] + [("layers_host", C.POINTER(VitLayer))]


class TextConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("width", "layers", "heads", "context", "vocab", "out_dim")]


class TextWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("token_embedding", "positional_embedding", "ln_final_g", "ln_final_b",
                                          "text_projection")] + [("layers_host", C.POINTER(VitLayer))]


_lib = None


def lib() -> C.CDLL:
    """The loaded library.  Raises NativeLibraryMissing if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python clip-ppo_b200/build.py` "
            "(or __graft_entry__.build()); there is no CPU fallback for this path")
    L = C.CDLL(LIB_PATH)
    vp, i, f, d, i64p, fp, sz = C.c_void_p, C.c_int, C.c_float, C.c_double, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.c_size_t
    L.clipppo_abi_version.restype = i
    L.clipppo_strerror.restype = C.c_char_p
    L.clipppo_strerror.argtypes = [i]
    L.clipppo_last_cuda_error.restype = i
    L.clipppo_prof_begin.argtypes = [i]
    L.clipppo_prof_bucket.argtypes = [i, C.POINTER(C.c_longlong), C.POINTER(d), C.POINTER(d), C.POINTER(C.c_longlong)]
    L.clipppo_prof_end.argtypes = [C.POINTER(C.c_longlong), C.POINTER(d), C.POINTER(d), C.POINTER(C.c_longlong)]
    L.clipppo_disturb_f32.argtypes = [vp, i64p, vp, i64p, vp, i, i, i, i, i, f, f, fp, i, i, i, i, i, vp]
    L.clipppo_disturb_u8_f32.argtypes = [vp, vp, vp, i, i, i, i, i, f, f, fp, i, i, i, i, i, vp]
    L.clipppo_disturb_nhwc_u8.argtypes = [vp, i, vp, i64p, vp, i, i, i, i, i, f, f, fp, i, i, i, i, i, vp]
    L.clipppo_disturb_ex.argtypes = [C.POINTER(DisturbDesc), vp]
    L.clipppo_cosine_loss_fwd.argtypes = [vp, vp, i, i, vp, vp, vp]
    L.clipppo_cosine_loss_bwd.argtypes = [vp, vp, vp, vp, i, i, vp, vp, vp]
    L.clipppo_gae_f32.argtypes = [vp, vp, vp, vp, vp, i, i, d, d, vp, vp, vp]
    L.clipppo_ppo_loss_f32.argtypes = [vp] * 8 + [i, f, f, f, f, i, i, vp, vp, vp, vp, vp]
    L.clipppo_vit_create.argtypes = [C.POINTER(vp), C.POINTER(VitConfig), C.POINTER(VitWeights)]
    L.clipppo_vit_destroy.argtypes = [vp]
    L.clipppo_vit_workspace_bytes.argtypes = [vp, i, C.POINTER(sz)]
    L.clipppo_vit_encode.argtypes = [vp, vp, i, i64p, i, i, i, i, f, i, vp, vp, sz, vp]
    L.clipppo_preprocess_bf16.argtypes = [vp, i, i64p, i, i, i, i, f, i, i, i, vp, vp]
    L.clipppo_layernorm_bf16.argtypes = [vp, vp, vp, i, i, C.c_int64, vp, vp]
    L.clipppo_gemm_bf16.argtypes = [vp, vp, i, i, i, i, vp, vp, i, vp, C.c_int64, vp]
    L.clipppo_gemm_bf16_fused.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp, C.c_int64, vp]
    L.clipppo_gemm_bf16_resid_stats.argtypes = [vp, vp, i, i, i, vp, vp, C.c_int64, vp, vp]
    L.clipppo_gemm_bf16_fused_parts.argtypes = [vp, vp, i, i, i, i, vp, vp, i, vp, vp, C.c_int64, vp]
    L.clipppo_rowstats_bf16.argtypes = [vp, i, i, C.c_int64, vp, vp]
    if hasattr(L, "clipppo_gemm_bf16_probe"):            # probe builds only (CLIPPPO_BUILD_PROBES=1)
        L.clipppo_gemm_bf16_probe.argtypes = [vp, vp, i, i, i, i, vp, vp, C.c_int64, i, vp]
    L.clipppo_attention_bf16.argtypes = [vp, i, i, i, i, vp, vp]
    L.clipppo_attention_causal_bf16.argtypes = [vp, i, i, i, i, vp, vp]
    L.clipppo_nature_workspace_bytes.argtypes = [i, i, C.POINTER(sz)]
    L.clipppo_nature_forward.argtypes = [vp, C.POINTER(C.c_int64), C.c_float, i, i] + [vp] * 8 + [vp, vp, sz, vp]
    L.clipppo_nature_backward.argtypes = [vp, vp, vp, C.POINTER(C.c_int64), C.c_float, i, i, vp, vp] + [vp] * 8 + [vp, sz, vp]
    L.clipppo_text_create.argtypes = [C.POINTER(vp), C.POINTER(TextConfig), C.POINTER(TextWeights)]
    L.clipppo_text_destroy.argtypes = [vp]
    L.clipppo_text_workspace_bytes.argtypes = [vp, i, C.POINTER(sz)]
    L.clipppo_text_encode.argtypes = [vp, vp, i, i, vp, vp, sz, vp]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("clipppo_strerror",):
            fn.restype = i
    _lib = L
    return L


_EXC = {
    ERR_BAD_SHAPE: ValueError, ERR_BAD_CHANNELS: TypeError, ERR_BAD_PAD: RuntimeError, ERR_NULL: ValueError,
    ERR_WORKSPACE: RuntimeError, ERR_UNSUPPORTED: NotImplementedError, ERR_ALIGN: ValueError,
    ERR_CUDA: RuntimeError, ERR_DIM_MISMATCH: ValueError,
}


def check(status: int, what: str = "") -> None:
    """Map a C status to the exception type the reference raises in the same situation
    (torchvision: TypeError for bad channel count, RuntimeError for over-large reflect pad;
    reference shared/clip_ppo_utils.py:62-64: ValueError for width mismatch)."""
    if status == OK:
        return
    msg = lib().clipppo_strerror(status).decode()
    if status == ERR_CUDA:
        msg += f" [cudaError {lib().clipppo_last_cuda_error()}]"
    raise _EXC.get(status, RuntimeError)(f"{what}: {msg}" if what else msg)


_NO_SWITCH = contextlib.nullcontext()


def device_ctx(dev):
    """``torch.cuda.device(dev)`` only when ``dev`` is not already current (the guard costs ~5 us of host time
    per call, a fifth of an env-step disturbance call)."""
    import torch
    return _NO_SWITCH if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)


def strides4(t) -> "C.Array":
    return (C.c_int64 * 4)(*t.stride())
