"""ctypes binding of libgof_b200.so (include/gof_b200.h).

The CUDA library is the product: there is NO fallback.  If it is missing or cannot be loaded
this module raises at import time.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GOF_B200_LIB") or os.path.join(_HERE, "libgof_b200.so")   # override: kernel-variant experiments

FLAG_EXACT_BLEND = 1
FLAG_SAVE_CONTRIB = 2


_environ_get = os.environ.get


def default_flags(*differentiated) -> int:
    """Blend arithmetic mode: bit-exact contributing path iff GOF_EXACT_BLEND=1; FLAG_SAVE_CONTRIB when one of the
    tensors passed requires grad, i.e. a backward may follow this forward (see gof_b200.h; GOF_SAVE_CONTRIB=0/1 forces it)."""
    flags = FLAG_EXACT_BLEND if _environ_get("GOF_EXACT_BLEND", "0") not in ("0", "", "false") else 0
    force = _environ_get("GOF_SAVE_CONTRIB", "")
    if force == "1":
        return flags | FLAG_SAVE_CONTRIB
    if force != "0":
        for t in differentiated:
            if t is not None and t.requires_grad:
                return flags | FLAG_SAVE_CONTRIB
    return flags


GOF_OK = 0
GOF_EINVAL, GOF_ECUDA, GOF_ENOMEM, GOF_EOVERFLOW = -1, -2, -3, -4
SINK_CHW, SINK_HWC = 0, 1
OUTPUT_CHANNELS = 9
SINK_CHANNELS = 5


class GofParams(Structure):
    _fields_ = [
        ("P", c_int32), ("D", c_int32), ("M", c_int32), ("W", c_int32), ("H", c_int32),
        ("tan_fovx", c_float), ("tan_fovy", c_float), ("kernel_size", c_float), ("scale_modifier", c_float),
        ("prefiltered", c_int32), ("debug", c_int32), ("flags", c_int32),
    ]


class GofInputs(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "background", "means3D", "shs", "colors_precomp", "opacities", "scales", "rotations",
        "cov3D_precomp", "view2gaussian_precomp", "viewmatrix", "projmatrix", "campos")]


class GofHeadParams(Structure):
    _fields_ = [("BV", c_int32), ("H", c_int32), ("W", c_int32), ("C", c_int32), ("with_offset", c_int32),
                ("sh_degree", c_int32), ("isotropic", c_int32), ("squre_clip", c_float)]


class GofGrads(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh",
        "dL_dscales", "dL_drotations", "dL_dview2gaussian")]


ALLOC_FN = ctypes.CFUNCTYPE(c_void_p, c_void_p, c_size_t)

# name -> (restype, argtypes); every symbol include/gof_b200.h declares
SIGNATURES = {
    "gof_last_error": (c_char_p, []),
    "gof_version": (c_char_p, []),
    "gof_context_create": (c_int32, [c_int32, POINTER(c_void_p)]),
    "gof_context_destroy": (None, [c_void_p]),
    "gof_profile_enable": (c_int32, [c_void_p, c_int32]),
    "gof_profile_read": (c_int32, [c_void_p, POINTER(ctypes.c_double), POINTER(c_int64), POINTER(ctypes.c_double), POINTER(c_int64)]),
    "gof_state_sizes": (c_int32, [c_int32, c_int32, c_int32, c_int64, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t)]),
    "gof_forward": (c_int32, [c_void_p, POINTER(GofParams), POINTER(GofInputs), c_void_p, c_size_t, c_void_p, c_size_t,
                              c_void_p, c_size_t, ALLOC_FN, c_void_p, c_void_p, c_void_p,
                              POINTER(c_int32), POINTER(c_void_p), c_void_p]),
    "gof_state_sizes_batch": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_int64, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t)]),
    "gof_forward_batch": (c_int32, [c_void_p, POINTER(GofParams), POINTER(GofInputs), c_int32, c_int32, c_void_p, c_size_t,
                                    c_void_p, c_size_t, c_void_p, c_size_t, ALLOC_FN, c_void_p, c_void_p, c_void_p,
                                    POINTER(c_int32), POINTER(c_void_p), c_void_p]),
    "gof_state_get_batch": (c_int64, [c_char_p, c_int32, c_int32, c_int32, c_int32, c_int64, c_void_p, c_void_p, c_size_t,
                                      c_void_p, c_void_p, c_int64, c_void_p]),
    "gof_integrate": (c_int32, [c_void_p, POINTER(GofParams), POINTER(GofInputs), c_int32, c_void_p, c_void_p, c_size_t,
                                c_void_p, c_size_t, ALLOC_FN, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                POINTER(c_int32), c_void_p]),
    "gof_predictor_head": (c_int32, [POINTER(GofHeadParams)] + [c_void_p] * 15),
    "gof_set_frame_sink": (c_int32, [c_void_p, c_void_p, c_size_t, c_int32]),
    "gof_num_rendered": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, POINTER(c_int32)]),
    "gof_num_rendered_async": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "gof_backward": (c_int32, [c_void_p, POINTER(GofParams), POINTER(GofInputs), c_int32, c_void_p,
                               c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, POINTER(GofGrads), c_void_p]),
    "gof_backward_batch": (c_int32, [c_void_p, POINTER(GofParams), POINTER(GofInputs), c_int32, c_int32, c_int64, c_void_p,
                                     c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, POINTER(GofGrads), c_void_p]),
    "gof_preprocess_backward": (c_int32, [c_void_p, POINTER(GofParams), POINTER(GofInputs), c_void_p, c_void_p,
                                          c_void_p, c_void_p, POINTER(GofGrads), c_void_p]),
    "gof_mark_visible": (c_int32, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gof_render_epilogue": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "gof_render_epilogue_batch": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "gof_render_epilogue_backward_batch": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_float, c_void_p,
                                                     c_void_p, c_void_p, c_void_p]),
    "gof_pack_gather": (c_int32, [c_void_p, c_int32, c_int64, c_void_p, c_int32, c_void_p, c_int64, c_void_p]),
    "gof_backward_accumulators": (c_int64, [c_void_p, c_void_p, c_int64, c_void_p]),
    "gof_state_get": (c_int64, [c_char_p, c_int32, c_int32, c_int32, c_int64, c_void_p, c_void_p, c_size_t, c_void_p,
                                c_void_p, c_int64, c_void_p]),
}


def load(path: str = LIB_PATH) -> ctypes.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: the sm_100a CUDA library is required (no fallback exists). "
            "Build it with `python -m f3d_gaus_b200.build` or `__graft_entry__.build()`.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


def raw_stream(device) -> int:
    """cudaStream_t of torch's current stream on `device` (the hot-path form of
    torch.cuda.current_stream(device).cuda_stream: no Stream object is built)."""
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device())
    except AttributeError:                      # private helper not present in this torch build
        return torch.cuda.current_stream(device).cuda_stream


def last_error() -> str:
    msg = lib.gof_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str) -> None:
    if rc != GOF_OK:
        raise RuntimeError(f"{what} failed ({rc}): {last_error()}")


_contexts: dict[int, c_void_p] = {}


def context(device_index: int) -> c_void_p:
    ctx = _contexts.get(device_index)
    if ctx is None:
        out = c_void_p()
        check(lib.gof_context_create(device_index, ctypes.byref(out)), "gof_context_create")
        ctx = _contexts[device_index] = out
    return ctx


FWD_STAGES = ("preprocess", "scan", "num_rendered_handoff", "binning", "blend")
BWD_STAGES = ("clear", "blend_backward", "preprocess_backward")


def profile_enable(device_index: int, on: bool) -> None:
    check(lib.gof_profile_enable(context(device_index), int(on)), "gof_profile_enable")


def profile_read(device_index: int) -> dict:
    """{'fwd_calls', 'bwd_calls', 'fwd_ms': {stage: total ms}, 'bwd_ms': {...}} since the last read."""
    f = (ctypes.c_double * 5)()
    b = (ctypes.c_double * 3)()
    nf, nb = c_int64(0), c_int64(0)
    check(lib.gof_profile_read(context(device_index), f, ctypes.byref(nf), b, ctypes.byref(nb)), "gof_profile_read")
    return {"fwd_calls": nf.value, "bwd_calls": nb.value, "fwd_ms": dict(zip(FWD_STAGES, list(f))),
            "bwd_ms": dict(zip(BWD_STAGES, list(b)))}
