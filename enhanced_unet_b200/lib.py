"""ctypes binding of libeunet_b200.so (the C ABI declared in include/eunet.h).

The product path has no CPU fallback: if the library is missing or a call fails, a RuntimeError is
raised (``eunet_last_error`` text included).  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeunet_b200.so")

F32, BF16, F16 = 0, 1, 2
ABI_VERSION = 2

_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_f = C.c_float

# name -> argtypes (restype is always int unless listed in _RESTYPES)
SIGNATURES = {
    "eunet_abi_version": [],
    "eunet_device_info": [_p, _p, _p, _p],
    "eunet_set_option": [C.c_char_p, _i],
    "eunet_confusion4x4": [_p, _p, _i, _ll, _ll, _p, _p],
    "eunet_pack_mask_bits": [_p, _i, _ll, _p, _p, _p],
    "eunet_pair_intersections": [_p, _i, _p, _i, _ll, _p, _p],
    "eunet_pack_input_nchw": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "eunet_pack_weight3x3": [_p, _p, _i, _i, _i, _i, _i, _i, _p],
    "eunet_pack_weight3x3_multi": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "eunet_unpack_wgrad3x3_multi": [_p, _p, _p, _p, _p, _p, _i, _p, _p],
    "eunet_unpack_wgrad3x3": [_p, _p, _i, _i, _i, _i, _p, _p],
    "eunet_conv3x3_fwd": [_p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _i, _p, _p],
    "eunet_conv3x3_tail_fwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "eunet_conv3x3_dgrad_few": [_p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "eunet_conv3x3_wgrad": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _p],
    "eunet_bn_finalize": [_p, _ll, _p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _i, _p],
    "eunet_bn_fold_eval": [_p, _p, _p, _p, _p, _f, _p, _p, _i, _p],
    "eunet_bn_apply_relu": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "eunet_bn_bwd_reduce": [_p, _i, _p, _i, _i, _ll, _i, _p, _p, _p, _p, _p, _p],
    "eunet_bn_bwd_apply": [_p, _i, _p, _i, _p, _i, _i, _ll, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "eunet_bn_bwd_reduce_pool": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p],
    "eunet_bn_bwd_apply_pool": [_p, _i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "eunet_maxpool2_fwd": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p],
    "eunet_maxpool2_bwd": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "eunet_upsample2_fwd": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p],
    "eunet_upsample2_bwd": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p],
    "eunet_bn_apply_relu_upsample2": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "eunet_tail_dec1_fwd": [_p, _i, _i, _p, _p, _p, _ll, _p],
    "eunet_tail_up_fwd": [_p, _p, _p, _i, _i, _i, _i, _p],
    "eunet_tail_pack3": [_p, _p, _i, _i, _i, _p, _p],
    "eunet_tail_out_fwd": [_p, _p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "eunet_tail_bwd_reduce": [_p, _p, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "eunet_tail_bwd_dmid": [_p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "eunet_tail_bwd_fused": [_p] * 12 + [_i, _i, _i, _i, _p],
    "eunet_tail_up_bwd": [_p, _i, _i, _p, _p, _i, _i, _i, _p, _p],
    "eunet_tail_dec1_bwd": [_p, _p, _i, _p, _i, _i, _p, _p, _ll, _p],
    "eunet_cast_f64_f32": [_p, _p, _ll, _p, _p],
    "eunet_grad_scale": [_p, _ll, _f, _p, _p],
    "eunet_loss_fwd": [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p],
    "eunet_loss_bwd": [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p],
    "eunet_resize_bilinear": [_p, _p, _i, _i, _i, _i, _i, _f, _f, _p],
    "eunet_softmax_probs": [_p, _p, _i, _i, _i, _i, _p],
    "eunet_reflect_pad": [_p, _p, _i, _i, _i, _i, _i, _p],
    "eunet_tta_combine": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "eunet_probs_to_mask": [_p, _p, _p, _i, _i, _i, _p],
    "eunet_fusion_gate_fwd": [_p, _p, _p, _p, _i, _p, _i, _i, _i, _p],
    "eunet_fusion_out_fwd": [_p, _p, _p, _i, _i, _i, _p],
    "eunet_fusion_gate_conv_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "eunet_fusion_gate_mid_fwd": [_p, _p, _p, _p, _p, _p, _ll, _p],
    "eunet_fusion_gate_apply_fwd": [_p, _p, _p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _p],
    "eunet_fusion_gate_bwd": [_p] * 7 + [_i] + [_p] * 8 + [_i, _i, _i, _p],
    "eunet_channel_scale": [_p, _i, _p, _i, _i, _ll, _i, _p],
    "eunet_sumsq": [_p, _ll, _p, _p],
    "eunet_sumsq_multi": [_p, _p, _i, _p, _p],
    "eunet_adamw_multi": [_p, _p, _p, _p, _p, _i, _p, _f, _f, _f, _f, _f, _f, _i, _f, _p, _p],
    "eunet_adamw_prepare": [_p, _p, _f, _f, _p, _p],
    "eunet_adamw_step": [_p, _p, _p, _p, _ll, _p, _f, _f, _f, _f, _f, _f, _i, _f, _p],
    "eunet_probe_cluster": [_p, _i, _i, _p],
    "eunet_probe_umma": [_p, _i, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p, _i, _p],
}

_lib: Optional[C.CDLL] = None


def library_path() -> str:
    return LIB_PATH


def load() -> C.CDLL:
    """Load (once) and type the shared library.  Raises RuntimeError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m enhanced_unet_b200.build` "
            "(there is no CPU / PyTorch fallback for the Enhanced-UNet hot path)")
    lib = C.CDLL(LIB_PATH)
    lib.eunet_last_error.restype = C.c_char_p
    lib.eunet_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = C.c_int
    v = lib.eunet_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"libeunet_b200.so ABI version {v} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def last_error() -> str:
    return load().eunet_last_error().decode("utf-8", "replace")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


# number of C-ABI calls per entry point since the last COUNTERS.clear() (each call launches >= 1 kernel)
COUNTERS: dict = {}
# when a list: every call is bracketed by CUDA events on the launching stream -> (name, tag, e0, e1, flops)
PROFILE: Optional[list] = None


def call(name: str, *args, flops: float = 0.0, tag: str = "") -> None:
    """Invoke ``name`` with the current torch CUDA stream appended; raise on a non-zero status.  ``flops`` (ALGORITHMIC
    FLOPs of the launch) and ``tag`` (roofline group of the launch) only feed the profile."""
    lib = load()
    COUNTERS[name] = COUNTERS.get(name, 0) + 1
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream_ptr())
        e1.record()
        PROFILE.append((name, tag, e0, e1, flops))
    else:
        rc = getattr(lib, name)(*args, stream_ptr())
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {last_error()}")


def collect_profile(by_tag: bool = False) -> dict:
    """name (or (name, tag) with ``by_tag``) -> (total flops, total ms, launches) for the calls recorded since PROFILE was set."""
    torch.cuda.synchronize()
    out: dict = {}
    for name, tag, e0, e1, fl in PROFILE or []:
        key = (name, tag) if by_tag else name
        f, ms, n = out.get(key, (0.0, 0.0, 0))
        out[key] = (f + fl, ms + e0.elapsed_time(e1), n + 1)
    return out


def raw_dtype(act: torch.dtype) -> torch.dtype:
    """Storage dtype of RAW (pre-BatchNorm) conv outputs for a given activation dtype."""
    return torch.float16 if act in (torch.bfloat16, torch.float16) else torch.float32


def set_option(name: str, value: int) -> None:
    rc = load().eunet_set_option(name.encode(), int(value))
    if rc != 0:
        raise RuntimeError(f"eunet_set_option failed: {last_error()}")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.bfloat16:
        return BF16
    if dt == torch.float16:
        return F16
    if dt == torch.float32:
        return F32
    raise ValueError(f"unsupported activation dtype {dt}")
