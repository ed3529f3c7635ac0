"""The ConvAttention projection stacks on the sm_100a kernels.

Reference: tts/models/acoustic/modules/alignment.py:40-83 (ConvBlock1D: mask -> Conv1d -> activation -> masked norm ->
dropout), :118-154 (the key and query stacks), :176-187 (how forward calls them); tts/modules/normalization.py:160-208.

Per block:  y = act(conv(x * mask))        isp_gemm_batched as an implicit GEMM over the kernel taps of a channels-last
                                           activation (tcgen05, GELU / ReLU and the norm's column sums in the epilogue)
            x' = norm(y) * mask            isp_instance_norm_apply (one read, one write)
and the last, pointwise, un-normalised block writes the (B, T, attention_dim) K-major operand of isp_loglik_forward directly.
The input goes through isp_prep_channels_last once (transpose to channels-last, mask, cast).  Inference / no-grad only:
with gradients enabled the torch restatement in alignment.py runs instead (the backward of the stacks is not built).
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from .gemm import ACT, conv1d_channels_last

__all__ = ["fused_supported", "project_stack"]

_ACT_NAME = {nn.Identity: "linear", nn.ReLU: "relu", nn.GELU: "gelu"}


def _block_ok(blk, vec: int) -> bool:
    from .alignment import MaskedInstanceNorm1d
    conv = blk.conv
    if type(blk.act) not in _ACT_NAME or (isinstance(blk.act, nn.GELU) and getattr(blk.act, "approximate", "none") != "none"):
        return False
    if conv.bias is not None or conv.stride != (1,) or conv.dilation != (1,) or conv.groups != 1:
        return False
    k = conv.kernel_size[0]
    if k % 2 != 1 or conv.padding != ((k - 1) // 2,):
        return False
    if blk.norm is not None and type(blk.norm) is not MaskedInstanceNorm1d:
        return False
    return conv.out_channels % vec == 0


def fused_supported(blocks, dtype: torch.dtype) -> bool:
    """True when every block of the stack is something the kernels cover: odd 'same' kernel, no bias, linear / ReLU / GELU,
    instance norm or none, channel counts that are whole 16 B vectors."""
    vec = 4 if dtype == torch.float32 else 8
    return all(_block_ok(b, vec) for b in blocks) and blocks[-1].norm is None


_WCACHE: dict = {}


def _taps(conv: nn.Conv1d, dtype: torch.dtype, cin_p: int) -> torch.Tensor:
    """conv.weight (Cout, Cin, k) -> (k, Cout, cin_p) in the operand dtype, input channels zero-padded; cached per version."""
    w = conv.weight
    key = (id(conv), dtype, cin_p)
    hit = _WCACHE.get(key)
    if hit is not None and hit[0] == w._version and hit[1] == w.data_ptr():
        return hit[2]
    t = w.detach().permute(2, 0, 1)
    if cin_p != t.shape[2]:
        t = torch.nn.functional.pad(t, (0, cin_p - t.shape[2]))
    t = t.to(dtype).contiguous()
    _WCACHE[key] = (w._version, w.data_ptr(), t)
    return t


def prep_channels_last(x: torch.Tensor, lengths: torch.Tensor, channels: int, dtype: torch.dtype) -> torch.Tensor:
    """(B, C, T) or (B, T, C) -> (B, T, Cp) `dtype`, masked; Cp = C rounded up to a whole 16 B vector."""
    dev = x.device
    lib = _lib.load()
    channels_first = x.shape[1] == channels           # the rule of ConvAttention.forward (alignment.py:176-177 transposes otherwise)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    x = x.contiguous()
    B = x.shape[0]
    T = x.shape[2] if channels_first else x.shape[1]
    vec = 4 if dtype == torch.float32 else 8
    cp = (channels + vec - 1) // vec * vec
    out = torch.empty((B, T, cp), dtype=dtype, device=dev)
    dt = _lib.dtype_code
    with torch.cuda.device(dev):
        rc = lib.isp_prep_channels_last(x.data_ptr(), dt(x.dtype), 1 if channels_first else 0, lengths.data_ptr(), out.data_ptr(),
                                        dt(dtype), B, channels, T, cp, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "isp_prep_channels_last")
    return out


def instance_norm_apply(y: torch.Tensor, stats: torch.Tensor, norm, lengths: torch.Tensor) -> torch.Tensor:
    """In-place masked instance norm of the channels-last y (B, T, C) from the GEMM's column statistics."""
    dev = y.device
    lib = _lib.load()
    B, T, C = y.shape
    w = norm.weight.detach().float().contiguous() if norm.weight is not None else None
    b = norm.bias.detach().float().contiguous() if norm.bias is not None else None
    dt = _lib.dtype_code(y.dtype)
    ws = torch.empty((B, C, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.isp_instance_norm_apply(y.data_ptr(), dt, stats.data_ptr(), stats.shape[1], w.data_ptr() if w is not None else None,
                                         b.data_ptr() if b is not None else None, lengths.data_ptr(), y.data_ptr(), B, T, C,
                                         y.stride(1), y.stride(1), float(norm.eps), ws.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "isp_instance_norm_apply")
    return y


@torch.no_grad()
def project_stack(blocks, x: torch.Tensor, lengths: torch.Tensor, in_channels: int, dtype: torch.dtype,
                  out_dtype: torch.dtype | None = None) -> torch.Tensor:
    """One projection stack: x (B, C, T) or (B, T, C) -> (B, T, attention_dim) in `out_dtype` (default `dtype`), rows past
    `lengths` zero.  `dtype` is the type of the activations and weights inside the stack: float16 (the reference trains under
    fp16 autocast, recipes/default.yaml:56), bfloat16, or float32 (TF32 products)."""
    dev = x.device
    _lib.require_device(dev)
    lengths = lengths.to(device=dev, dtype=torch.int64).contiguous()
    h = prep_channels_last(x, lengths, in_channels, dtype)
    for blk in blocks:
        w = _taps(blk.conv, dtype, h.shape[2])
        act = _ACT_NAME[type(blk.act)]
        if blk.norm is not None:
            h, stats = conv1d_channels_last(h, w, lengths, act=act, out_dtype=dtype, col_stats=True)
            h = instance_norm_apply(h, stats, blk.norm, lengths)
        else:
            h = conv1d_channels_last(h, w, lengths, act=act, out_dtype=(out_dtype or dtype) if blk is blocks[-1] else dtype)
    return h
