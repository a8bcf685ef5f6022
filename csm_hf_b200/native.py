"""ctypes binding of libcsm_b200.so (include/csm_b200.h).  No fallback: if the library is
missing or fails to load, importing the engine raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcsm_b200.so")

CSM_OK, CSM_EINVAL, CSM_ECUDA, CSM_ECAPACITY, CSM_EUNSUPPORTED = 0, -1, -2, -3, -4
W_PER_LAYER = 9
INFO_SMS, INFO_GRID, INFO_PHASES, INFO_SMEM, INFO_LAUNCHES, INFO_STEPPED = range(6)

EXPORTS = [
    "csm_create", "csm_destroy", "csm_reset", "csm_cache_len", "csm_embed_sum", "csm_generate_frame",
    "csm_generate", "csm_frames_done", "csm_generate_host", "csm_info", "csm_set_stepped",
    "csm_last_decode_ms", "csm_last_error", "csm_debug_copy", "csm_debug_run_phases", "csm_debug_set_cache_len",
    "csm_debug_profile_frame", "csm_debug_progress", "csm_set_sampling", "csm_sample_topk", "csm_linear",
    "csm_generate_more",
    "csm_train_create", "csm_train_destroy", "csm_train_last_error", "csm_train_launches", "csm_train_step",
    "csm_train_debug", "csm_train_step_begin", "csm_train_step_end",
]


class LlamaShape(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("inter", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32),
                ("kv_heads", C.c_int32), ("eps", C.c_float), ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
                ("n_pos", C.c_int32)]


class Shapes(C.Structure):
    _fields_ = [("text_vocab", C.c_int32), ("audio_vocab", C.c_int32), ("n_codebooks", C.c_int32),
                ("backbone", LlamaShape), ("decoder", LlamaShape)]


class Weights(C.Structure):
    _fields_ = [("text_embeddings", C.c_void_p), ("audio_embeddings", C.c_void_p), ("projection", C.c_void_p),
                ("codebook0_head", C.c_void_p), ("audio_head", C.c_void_p), ("backbone_norm", C.c_void_p),
                ("decoder_norm", C.c_void_p), ("backbone_layers", C.POINTER(C.c_void_p)),
                ("decoder_layers", C.POINTER(C.c_void_p))]


_lib = None


def load():
    """dlopen the engine; raises if it has not been built (python -m csm_hf_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CSM_LIB") or LIB_PATH   # experiment variants of the same library (tools/gpu_variants.sh)
    if not os.path.isfile(path):
        raise RuntimeError(f"{path} not found: build it with `python csm_hf_b200/build.py` "
                           "(there is no CPU or PyTorch fallback for the generation path)")
    lib = C.CDLL(path)
    vp, i32, i64p = C.c_void_p, C.c_int, C.c_void_p
    lib.csm_create.argtypes = [C.POINTER(Shapes), C.POINTER(Weights), i32, i32, vp, C.POINTER(vp)]
    lib.csm_create.restype = i32
    lib.csm_destroy.argtypes = [vp]; lib.csm_destroy.restype = i32
    lib.csm_reset.argtypes = [vp]; lib.csm_reset.restype = i32
    lib.csm_cache_len.argtypes = [vp]; lib.csm_cache_len.restype = i32
    lib.csm_embed_sum.argtypes = [vp, i64p, vp, i32, i32, vp, vp]; lib.csm_embed_sum.restype = i32
    lib.csm_generate_frame.argtypes = [vp, i64p, vp, i32, i32, i64p, i64p, vp, vp, vp, vp]
    lib.csm_generate_frame.restype = i32
    lib.csm_generate.argtypes = [vp, i64p, vp, i32, i32, i32, i32, i64p, vp]; lib.csm_generate.restype = i32
    lib.csm_frames_done.argtypes = [vp, vp]; lib.csm_frames_done.restype = i32
    lib.csm_generate_host.argtypes = [vp, i64p, vp, i32, i32, i32, i32, i64p, C.POINTER(C.c_int), vp]
    lib.csm_generate_host.restype = i32
    lib.csm_info.argtypes = [vp, i32]; lib.csm_info.restype = C.c_int64
    lib.csm_set_stepped.argtypes = [vp, i32]; lib.csm_set_stepped.restype = i32
    lib.csm_last_decode_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    lib.csm_last_decode_ms.restype = i32
    lib.csm_last_error.argtypes = [vp]; lib.csm_last_error.restype = C.c_char_p
    lib.csm_debug_copy.argtypes = [vp, i32, vp, C.c_int64, C.POINTER(C.c_int64), vp]; lib.csm_debug_copy.restype = i32
    lib.csm_debug_run_phases.argtypes = [vp, i64p, vp, i32, i32, i32, i32, vp]; lib.csm_debug_run_phases.restype = i32
    lib.csm_debug_set_cache_len.argtypes = [vp, i32]; lib.csm_debug_set_cache_len.restype = i32
    lib.csm_debug_profile_frame.argtypes = [vp, i32, vp, vp, vp]; lib.csm_debug_profile_frame.restype = i32
    lib.csm_debug_progress.argtypes = [vp, vp, vp]; lib.csm_debug_progress.restype = i32
    lib.csm_set_sampling.argtypes = [vp, i32, C.c_float, C.c_uint64, i32]; lib.csm_set_sampling.restype = i32
    lib.csm_sample_topk.argtypes = [vp, i32, i32, i32, C.c_float, C.c_uint64, vp, vp]; lib.csm_sample_topk.restype = i32
    lib.csm_generate_more.argtypes = [vp, i32, i32, i32, i64p, vp]; lib.csm_generate_more.restype = i32
    lib.csm_train_create.argtypes = [C.POINTER(Shapes), i32, i32, C.POINTER(vp)]; lib.csm_train_create.restype = i32
    lib.csm_train_destroy.argtypes = [vp]; lib.csm_train_destroy.restype = i32
    lib.csm_train_last_error.argtypes = [vp]; lib.csm_train_last_error.restype = C.c_char_p
    lib.csm_train_launches.argtypes = [vp]; lib.csm_train_launches.restype = i32
    lib.csm_train_step.argtypes = [vp, C.POINTER(Weights), C.POINTER(Weights), i64p, vp, i64p, i32, i32,
                                   C.POINTER(C.c_float), C.POINTER(C.c_int), vp, vp, vp]
    lib.csm_train_step.restype = i32
    lib.csm_train_step_begin.argtypes = [vp, C.POINTER(Weights), C.POINTER(Weights), i64p, vp, i64p, i32, i32, i32,
                                         C.POINTER(C.c_int), vp, vp, vp]
    lib.csm_train_step_begin.restype = i32
    lib.csm_train_step_end.argtypes = [vp, C.POINTER(C.c_float), vp]; lib.csm_train_step_end.restype = i32
    lib.csm_train_debug.argtypes = [vp, C.c_char_p, vp, C.c_longlong, C.POINTER(C.c_longlong)]; lib.csm_train_debug.restype = i32
    lib.csm_linear.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, i32, vp]; lib.csm_linear.restype = i32
    _lib = lib
    return lib


def check(lib, ctx, rc: int):
    """Map the C error convention onto the exceptions the reference raises."""
    if rc >= 0:
        return rc
    msg = lib.csm_last_error(ctx).decode() if ctx else "csm error"
    if rc in (CSM_EINVAL, CSM_ECAPACITY):
        raise ValueError(msg)
    if rc == CSM_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)
