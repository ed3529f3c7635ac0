"""ctypes binding of libisp_tts_b200.so (the C ABI in include/isp_tts_b200.h).

No fallback: if the library is missing or the device is not a B200-class GPU,
every entry point raises.
"""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# (ISP_TTS_B200_LIB: another build of the same library, for A/B timing of two builds on one box -- tools/ab_build.sh)
SO_PATH = os.environ.get("ISP_TTS_B200_LIB") or os.path.join(_PKG, "libisp_tts_b200.so")

ISP_DTYPE_F32 = 0
ISP_DTYPE_BF16 = 1
ISP_DTYPE_F16 = 2
ISP_ALIGN_WS_CLEAN = 1

EXPORTS = [
    "isp_version", "isp_last_error", "isp_device_check",
    "isp_mas_workspace_bytes", "isp_mas_forward", "isp_mas_forward_path", "isp_bin_loss_sums", "isp_length_regulate", "isp_length_regulate_backward", "isp_path_from_durations", "isp_temporal_average", "isp_ctc_workspace_bytes", "isp_ctc_forward", "isp_ctc_backward", "isp_mas_status",
    "isp_align_workspace_bytes", "isp_align_forward", "isp_loglik_supported", "isp_stage_operands", "isp_unpack_workspace_bytes", "isp_unpack_operands", "isp_loglik_workspace_bytes", "isp_loglik_forward", "isp_split_3xtf32", "isp_loglik_rows", "isp_loglik_backward_ds", "isp_loglik_backward_from_logits", "isp_gemm_batched", "isp_prep_channels_last", "isp_instance_norm_apply", "isp_soft_average_workspace_bytes", "isp_soft_average", "isp_soft_average_backward", "isp_set_option",
]

_lib = None


class GemmDesc(ctypes.Structure):
    """isp_gemm_desc of include/isp_tts_b200.h."""
    _fields_ = [("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("c", ctypes.c_void_p),
                ("m_len", ctypes.c_void_p), ("n_len", ctypes.c_void_p), ("k_len", ctypes.c_void_p),
                ("col_stats", ctypes.c_void_p),
                ("lda", ctypes.c_int64), ("ldb", ctypes.c_int64), ("ldc", ctypes.c_int64),
                ("a_batch", ctypes.c_int64), ("b_batch", ctypes.c_int64), ("c_batch", ctypes.c_int64), ("b_tap_stride", ctypes.c_int64),
                ("batch", ctypes.c_int32), ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
                ("dtype_ab", ctypes.c_int32), ("dtype_c", ctypes.c_int32), ("a_mn_major", ctypes.c_int32), ("b_mn_major", ctypes.c_int32),
                ("taps", ctypes.c_int32), ("tap_shift", ctypes.c_int32), ("act", ctypes.c_int32), ("bn", ctypes.c_int32),
                ("skip_padding", ctypes.c_int32), ("alpha", ctypes.c_float), ("stages", ctypes.c_int32), ("trace", ctypes.c_void_p)]


class IspError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise IspError(
            f"{SO_PATH} is missing: build it with `python -m isp_tts_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU or PyTorch fallback for this path.")
    lib = ctypes.CDLL(SO_PATH)
    c_int, c_i64, c_sz, vp, f32 = ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_float
    lib.isp_version.restype = c_int
    lib.isp_last_error.restype = ctypes.c_char_p
    lib.isp_device_check.restype = c_int
    lib.isp_mas_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.isp_mas_workspace_bytes.restype = c_sz
    lib.isp_mas_forward.argtypes = [vp, c_i64, c_i64, c_i64, vp, vp, c_int, c_int, c_int, vp, vp, vp, c_sz, vp]
    lib.isp_mas_forward.restype = c_int
    lib.isp_mas_forward_path.argtypes = [vp, c_i64, c_i64, c_i64, vp, vp, c_int, c_int, c_int, vp, vp, vp, vp, c_sz, vp]
    lib.isp_mas_forward_path.restype = c_int
    lib.isp_bin_loss_sums.argtypes = [vp, vp, vp, c_int, c_int, c_int, f32, vp, vp]
    lib.isp_bin_loss_sums.restype = c_int
    lib.isp_length_regulate.argtypes = [vp, vp, vp, c_int, c_int, c_int, c_int, c_int, vp]
    lib.isp_length_regulate.restype = c_int
    lib.isp_length_regulate_backward.argtypes = [vp, vp, vp, vp, c_int, c_int, c_int, c_int, vp]
    lib.isp_length_regulate_backward.restype = c_int
    lib.isp_path_from_durations.argtypes = [vp, vp, c_int, c_int, c_int, vp]
    lib.isp_path_from_durations.restype = c_int
    lib.isp_temporal_average.argtypes = [vp, vp, vp, c_int, c_int, c_int, c_int, vp]
    lib.isp_temporal_average.restype = c_int
    lib.isp_ctc_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.isp_ctc_workspace_bytes.restype = c_sz
    lib.isp_ctc_forward.argtypes = [vp, vp, vp, c_int, c_int, c_int, f32, vp, vp, c_sz, vp]
    lib.isp_ctc_forward.restype = c_int
    lib.isp_ctc_backward.argtypes = [vp, vp, vp, c_int, c_int, c_int, f32, vp, vp, vp, vp, c_sz, vp]
    lib.isp_ctc_backward.restype = c_int
    lib.isp_mas_status.argtypes = [vp, vp]
    lib.isp_mas_status.restype = c_int
    lib.isp_stage_operands.argtypes = [vp, vp, c_int, vp, vp, c_int, c_int, c_int, c_int, vp, vp, vp]
    lib.isp_stage_operands.restype = c_int
    lib.isp_unpack_workspace_bytes.argtypes = [c_int]
    lib.isp_unpack_workspace_bytes.restype = c_sz
    lib.isp_unpack_operands.argtypes = [vp, vp, c_int, vp, vp, c_int, c_int, c_int, c_int, vp, vp, vp, c_sz, vp]
    lib.isp_unpack_operands.restype = c_int
    if os.environ.get("ISP_TTS_B200_LIB") and not hasattr(lib, "isp_align_forward"):      # an older build under A/B timing
        _lib = lib
        return lib
    lib.isp_align_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int]
    lib.isp_align_workspace_bytes.restype = c_sz
    lib.isp_align_forward.argtypes = [vp, vp, c_int, vp, vp, c_int, c_int, c_int, c_int, f32, c_int, vp, vp, vp, vp, vp, vp, vp, c_sz, c_int, vp]
    lib.isp_align_forward.restype = c_int
    lib.isp_loglik_supported.argtypes = [c_int, c_int, c_int]
    lib.isp_loglik_supported.restype = c_int
    lib.isp_loglik_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int]
    lib.isp_loglik_workspace_bytes.restype = c_sz
    lib.isp_loglik_forward.argtypes = [vp, vp, c_int, vp, vp, c_int, c_int, c_int, c_int, f32, c_int, vp, vp, vp, c_sz, vp]
    lib.isp_loglik_forward.restype = c_int
    lib.isp_split_3xtf32.argtypes = [vp, c_i64, c_int, c_int, vp, vp]
    lib.isp_split_3xtf32.restype = c_int
    lib.isp_loglik_rows.argtypes = [vp, c_i64, vp, vp, c_int, c_int, c_int, f32, c_int, vp, vp, vp]
    lib.isp_loglik_rows.restype = c_int
    lib.isp_loglik_backward_ds.argtypes = [vp, vp, vp, vp, c_int, c_int, c_int, f32, c_int, vp, c_int, vp]
    lib.isp_loglik_backward_ds.restype = c_int
    lib.isp_loglik_backward_from_logits.argtypes = [vp, vp, vp, vp, vp, vp, c_int, c_int, c_int, f32, c_int, vp, c_int, vp]
    lib.isp_loglik_backward_from_logits.restype = c_int
    lib.isp_soft_average_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    lib.isp_soft_average_workspace_bytes.restype = c_sz
    lib.isp_soft_average.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, c_int, c_int, vp, c_sz, vp]
    lib.isp_soft_average.restype = c_int
    lib.isp_soft_average_backward.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, c_int, c_int, vp]
    lib.isp_soft_average_backward.restype = c_int
    lib.isp_prep_channels_last.argtypes = [vp, c_int, c_int, vp, vp, c_int, c_int, c_int, c_int, c_int, vp]
    lib.isp_prep_channels_last.restype = c_int
    lib.isp_instance_norm_apply.argtypes = [vp, c_int, vp, c_int, vp, vp, vp, vp, c_int, c_int, c_int, c_i64, c_i64, f32, vp, vp]
    lib.isp_instance_norm_apply.restype = c_int
    lib.isp_gemm_batched.argtypes = [ctypes.POINTER(GemmDesc), vp]
    lib.isp_gemm_batched.restype = c_int
    lib.isp_set_option.argtypes = [ctypes.c_char_p, c_int]
    lib.isp_set_option.restype = c_int
    _lib = lib
    return lib


def dtype_code(dtype) -> int:
    """ISP_DTYPE_* of a torch dtype."""
    import torch
    try:
        return {torch.float32: ISP_DTYPE_F32, torch.bfloat16: ISP_DTYPE_BF16, torch.float16: ISP_DTYPE_F16}[dtype]
    except KeyError:
        raise ValueError(f"dtype must be float32, bfloat16 or float16, got {dtype}") from None


def check(rc: int, what: str):
    if rc != 0:
        msg = load().isp_last_error().decode("utf-8", "replace")
        raise IspError(f"{what} failed (code {rc}): {msg}")


def set_option(key: str, value: int) -> int:
    rc = load().isp_set_option(key.encode(), int(value))
    if rc == -1 and key not in ("mas.cols_per_lane", "mas.ring_rows", "mas.slots", "mas.dbg", "mas.bits_global", "mas.no_tma", "mas.impl",
                               "mas2.min_pair_stages", "mas2.single", "mas2.fill_us", "mas2.together", "mas2.pace", "loglik.debug_scores", "stage.ctas"):
        raise IspError(f"unknown option {key!r}")
    return rc


_checked_devices = set()


def require_device(device) -> None:
    """Raise unless `device` is a CUDA device the kernels were built for."""
    import torch
    if device.type != "cuda":
        raise IspError(f"tensor is on {device}; this path runs on a B200 (CUDA, sm_100) only -- there is no CPU fallback")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _checked_devices:
        return
    with torch.cuda.device(idx):
        check(load().isp_device_check(), "isp_device_check")
    _checked_devices.add(idx)
