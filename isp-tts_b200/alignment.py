"""Aligner / ConvAttention with the reference's constructor, forward signature,
outputs and state_dict keys; the hot path runs in the sm_100a kernels.

Mirrors tts/models/acoustic/modules/alignment.py of the reference:
  batch_diagonal_prior :18-37   ConvBlock1D :40-83   ConvAttention :98-208
  AlignerOutput :216-220        Aligner :223-331
What is different underneath:
  * from the matmul on (:189-208) everything is ONE fused kernel (isp_loglik_forward):
    tcgen05 GEMM + scale + log_softmax + closed-form diagonal prior + masked softmax;
  * MAS, the dense hard path and the durations (:272-275, :291-331) are ONE kernel
    (isp_mas_forward); there is no numba, no CPU route, no device->host check per step;
  * the last 1x1 projection of each stack is evaluated as a matmul that emits the
    (B, T, D) layout the GEMM's TMA loads want (same weights, same state_dict keys).

"""
from __future__ import annotations

import inspect
import warnings
from collections.abc import Sequence
from dataclasses import dataclass
from typing import NamedTuple

import torch
import torch.nn as nn
from torch import Tensor
from torch.nn import functional as F

from . import _lib
from .gemm import bgemm
from .mas import mas_forward

__all__ = ["stage_operands", "unpack_operands", "pack_rows", "batch_diagonal_prior", "ConvBlock1D", "ConvAttention", "ConvAttentionConfig",
           "Aligner", "AlignerConfig", "AlignerOutput", "loglik_forward", "align_forward"]


MISSING = "???"   # same sentinel string omegaconf uses; the reference's configs compare against it


# ----------------------------------------------------------------------------------------------
# helpers the reference takes from tts/utils/functions.py and tts/modules/{layers,normalization}.py
# ----------------------------------------------------------------------------------------------
def _length_mask(lengths: Tensor, max_len: int) -> Tensor:
    """(B, max_len) bool, True on valid positions.  No .item() sync (functions.py:61-66 has one)."""
    return torch.arange(max_len, device=lengths.device)[None, :] < lengths[:, None]


_ACTIVATIONS = {
    "linear": nn.Identity, "relu": nn.ReLU, "leaky_relu": nn.LeakyReLU, "selu": nn.SELU, "tanh": nn.Tanh,
    "mish": nn.Mish, "swish": nn.SiLU, "gelu": nn.GELU, "sigmoid": nn.Sigmoid,
}   # tts/modules/layers.py:9-31


def _masked_normalize(x: Tensor, mask: Tensor, dims, eps: float):
    cnt = mask.sum(dims, keepdim=True)
    mean = (x * mask).sum(dims, keepdim=True) / cnt
    var = (((x * mask - mean) * mask) ** 2).sum(dims, keepdim=True) / cnt
    return mean, var


class MaskedInstanceNorm1d(nn.InstanceNorm1d):
    """Instance norm whose statistics only see valid positions (normalization.py:104-124, 160-208)."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1,
                 affine: bool = True, track_running_stats: bool = False):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)

    def forward(self, x: Tensor, mask: Tensor | None = None) -> Tensor:
        if mask is None:
            return super().forward(x)
        mean, var = _masked_normalize(x, mask.to(x.dtype), [2], self.eps)
        y = (x - mean) / (var + self.eps).sqrt()
        if self.weight is not None and self.bias is not None:
            y = y * self.weight.view(1, -1, 1) + self.bias.view(1, -1, 1)
        return y


class MaskedBatchNorm1d(nn.BatchNorm1d):
    """Batch norm over valid positions only (normalization.py:15-66, 160-208)."""

    def forward(self, x: Tensor, mask: Tensor | None = None) -> Tensor:
        if mask is None:
            return super().forward(x)
        use_batch_stats = self.training or (self.running_mean is None and self.running_var is None)
        if use_batch_stats:
            mean, var = _masked_normalize(x, mask.to(x.dtype), [0, 2], self.eps)
            if self.training and self.track_running_stats:
                mom = 0.0 if self.momentum is None else self.momentum
                if self.num_batches_tracked is not None:
                    self.num_batches_tracked.add_(1)
                    if self.momentum is None:
                        mom = 1.0 / float(self.num_batches_tracked)
                with torch.no_grad():
                    self.running_mean.mul_(1 - mom).add_(mom * mean.detach().view(-1))
                    self.running_var.mul_(1 - mom).add_(mom * var.detach().view(-1))
        else:
            mean, var = self.running_mean.view(1, -1, 1), self.running_var.view(1, -1, 1)
        y = (x - mean) / (var + self.eps).sqrt()
        if self.weight is not None and self.bias is not None:
            y = y * self.weight.view(1, -1, 1) + self.bias.view(1, -1, 1)
        return y


_NORMS = {"instance": MaskedInstanceNorm1d, "batch": MaskedBatchNorm1d}


class _ConfigInit:
    """`Cls.init(config=..., **overrides)` as the reference builds its modules (constructor.py:68-84):
    merge a mapping / dataclass config with keyword overrides, drop keys the constructor does not
    take (with a warning), refuse unset mandatory values."""

    @classmethod
    def init(cls, config=None, **parameters):
        merged = {}
        if config is not None:
            if hasattr(config, "to_dict"):
                merged.update(config.to_dict())
            elif hasattr(config, "items"):
                merged.update(dict(config.items()))
            else:
                merged.update({k: v for k, v in vars(config).items()})
        merged.update(parameters)
        merged = {k: v for k, v in merged.items() if not str(k).startswith("_")}
        accepted = inspect.signature(cls.__init__).parameters
        if "kwargs" not in accepted:
            unknown = [k for k in merged if k not in accepted]
            if unknown:
                warnings.warn(f"The following params are incompatible with the {cls.__name__} constructor, "
                              f"so they will be ignored: {unknown}.")
                for k in unknown:
                    merged.pop(k)
        unset = [k for k, v in merged.items() if isinstance(v, str) and v == MISSING]
        if unset:
            raise RuntimeError(f"The following params are mandatory to set: {unset}")
        return cls(**merged)


# ----------------------------------------------------------------------------------------------
# the fused log-likelihood as an autograd function
# ----------------------------------------------------------------------------------------------
def batch_diagonal_prior(text_lengths: Tensor, mel_lengths: Tensor, gamma: float = 0.1, threshold: float = 1e-4,
                         max_text: int | None = None, max_mel: int | None = None) -> Tensor:
    """Dense (B, T1max, T2max) prior, torch ops, for callers that want the tensor itself
    (alignment.py:18-37).  The Aligner does NOT call this: the fused kernel evaluates the
    same expression in registers."""
    dev = text_lengths.device
    t2 = int(text_lengths.max()) if max_text is None else max_text
    t1 = int(mel_lengths.max()) if max_mel is None else max_mel
    gt = torch.arange(t2, dtype=torch.float32, device=dev)[None, :] / text_lengths[:, None]
    gm = torch.arange(t1, dtype=torch.float32, device=dev)[None, :] / mel_lengths[:, None]
    g = gt[:, None, :] - gm[:, :, None]
    prior = torch.exp(-g ** 2 / (2 * gamma ** 2))
    valid = _length_mask(mel_lengths, t1)[:, :, None] & _length_mask(text_lengths, t2)[:, None, :]
    prior = prior * valid
    prior = prior / (prior.sum(dim=-1, keepdim=True) + 1e-5)
    return prior.masked_fill(prior < threshold, 0.0)


def _split_3xtf32(x: Tensor, role: int) -> Tensor:
    """(B, T, D) fp32 -> (B, T, 3 D): [hi | hi | lo] for frames (role 0), [hi | lo | hi] for tokens (role 1) (isp_split_3xtf32)."""
    lib = _lib.load()
    B, T, D = x.shape
    out = torch.empty((B, T, 3 * D), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.isp_split_3xtf32(x.data_ptr(), B * T, D, role, out.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "isp_split_3xtf32")
    return out


def _scores_fp32(q: Tensor, k: Tensor, text_len: Tensor | None, mel_len: Tensor | None) -> Tensor:
    """S = Q.K^T with fp32-faithful products on the tensor cores: the three TF32 partial products hi.hi' + hi.lo' + lo.hi' as one
    contraction over 3 D (what the reference's torch.matmul computes in true fp32, alignment.py:189)."""
    return bgemm(_split_3xtf32(q.contiguous(), 0), _split_3xtf32(k.contiguous(), 1).transpose(1, 2), m_len=mel_len, n_len=text_len)


def _loglik_cuda(q: Tensor, k: Tensor, text_len: Tensor, mel_len: Tensor, scale: float, prior: bool, want_rowsum: bool = False,
                 precision: str = "tf32"):
    """(attn_soft, attn_logits) -- and, with want_rowsum, the prior's row sums (B, T1) the fused kernel leaves for the backward
    pass (None when the shape went through the stand-alone row epilogue).  precision (fp32 operands only): "tf32" = one
    tensor-core product per term inside the fused kernel (10-bit mantissa), "fp32" = the 3xTF32 split (_scores_fp32) followed by
    the stand-alone row epilogue: the reference's own precision, ~3x the GEMM work and a round trip of the scores through HBM."""
    dev = q.device
    _lib.require_device(dev)
    lib = _lib.load()
    if q.dtype != k.dtype or q.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("Q and K must both be float32 or both bfloat16")
    B, T1, D = q.shape
    T2 = k.shape[1]
    if k.shape[0] != B or k.shape[2] != D:
        raise ValueError(f"shape mismatch: Q {tuple(q.shape)} vs K {tuple(k.shape)}")
    q = q.contiguous()
    k = k.contiguous()
    tl = text_len.to(device=dev, dtype=torch.int64).contiguous()
    ml = mel_len.to(device=dev, dtype=torch.int64).contiguous()
    logits = torch.empty((B, T1, T2), dtype=torch.float32, device=dev)
    soft = torch.empty((B, T1, T2), dtype=torch.float32, device=dev)
    dt = _lib.ISP_DTYPE_BF16 if q.dtype == torch.bfloat16 else _lib.ISP_DTYPE_F32
    faithful = precision == "fp32" and q.dtype == torch.float32
    if faithful or not lib.isp_loglik_supported(T2, D, dt):
        # outside the fused kernel's range (long-form text, odd attention_dim, fp32 operands too wide for shared memory), or fp32-faithful
        # products asked for: scores from the batched GEMM, then the stand-alone row epilogue (isp_loglik_rows) -- slower by the
        # scores' round trip through HBM
        s = _scores_fp32(q, k, tl, ml) if faithful and D % 4 == 0 else bgemm(q, k.transpose(1, 2), m_len=ml, n_len=tl)
        with torch.cuda.device(dev):
            rc = lib.isp_loglik_rows(s.data_ptr(), s.stride(1), tl.data_ptr(), ml.data_ptr(), B, T1, T2, float(scale),
                                     1 if prior else 0, logits.data_ptr(), soft.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "isp_loglik_rows")
        return (soft, logits, None) if want_rowsum else (soft, logits)
    rowsum = torch.empty((B, T1), dtype=torch.float32, device=dev) if want_rowsum else None
    with torch.cuda.device(dev):
        rc = lib.isp_loglik_forward(q.data_ptr(), k.data_ptr(), dt, tl.data_ptr(), ml.data_ptr(), B, T1, T2, D,
                                    float(scale), 1 if prior else 0, logits.data_ptr(), soft.data_ptr(),
                                    rowsum.data_ptr() if rowsum is not None else None, 4 * B * T1 if rowsum is not None else 0,
                                    torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "isp_loglik_forward")
    return (soft, logits, rowsum) if want_rowsum else (soft, logits)


def _align_cuda(q: Tensor, k: Tensor, text_len: Tensor, mel_len: Tensor, scale: float, prior: bool, return_path: bool = False,
                dense: bool = True, want_rowsum: bool = False, precision: str = "tf32"):
    """isp_align_forward: the log-likelihood kernel and the MAS kernel linked through per-utterance ready counts (the second
    starts under the first's last wave).  Returns (soft, logits, hard, durations, path) -- and the prior's row sums for the
    backward pass with want_rowsum (None when the fused log-likelihood kernel does not cover the shape)."""
    dev = q.device
    _lib.require_device(dev)
    lib = _lib.load()
    if q.dtype != k.dtype or q.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("Q and K must both be float32 or both bfloat16")
    B, T1, D = q.shape
    T2 = k.shape[1]
    if k.shape[0] != B or k.shape[2] != D:
        raise ValueError(f"shape mismatch: Q {tuple(q.shape)} vs K {tuple(k.shape)}")
    if not dense and not return_path:
        raise ValueError("dense=False needs return_path=True")
    dt = _lib.ISP_DTYPE_BF16 if q.dtype == torch.bfloat16 else _lib.ISP_DTYPE_F32
    if (precision == "fp32" and q.dtype == torch.float32) or not lib.isp_loglik_supported(T2, D, dt):
        soft, logits = _loglik_cuda(q, k, text_len, mel_len, scale, prior, precision=precision)
        out = mas_forward(logits, text_len, mel_len, durations=True, return_path=return_path, dense=dense)
        res = (soft, logits, out[0], out[1], (out[2] if return_path else None))
        return res + (None,) if want_rowsum else res
    q = q.contiguous()
    k = k.contiguous()
    tl = text_len.to(device=dev, dtype=torch.int64).contiguous()
    ml = mel_len.to(device=dev, dtype=torch.int64).contiguous()
    if tl.numel() != B or ml.numel() != B:
        raise ValueError("text_len / mel_len must have one entry per utterance")
    logits = torch.empty((B, T1, T2), dtype=torch.float32, device=dev)
    soft = torch.empty((B, T1, T2), dtype=torch.float32, device=dev)
    hard = torch.empty((B, T1, T2), dtype=torch.int16, device=dev) if dense else None
    dur = torch.empty((B, T2), dtype=torch.int64, device=dev)
    path = torch.empty((B, T1), dtype=torch.int16, device=dev) if return_path else None
    rowsum = torch.empty((B, T1), dtype=torch.float32, device=dev) if want_rowsum else None
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws, clean = _align_workspace(lib, dev, stream, B, T1, T2, D, dt)
        rc = lib.isp_align_forward(q.data_ptr(), k.data_ptr(), dt, tl.data_ptr(), ml.data_ptr(), B, T1, T2, D, float(scale),
                                   1 if prior else 0, logits.data_ptr(), soft.data_ptr(), hard.data_ptr() if hard is not None else None,
                                   dur.data_ptr(), path.data_ptr() if path is not None else None,
                                   rowsum.data_ptr() if rowsum is not None else None, ws.data_ptr(), ws.numel(),
                                   _lib.ISP_ALIGN_WS_CLEAN if clean else 0, stream)
    if rc != 0:
        _ALIGN_WS.clear()
    _lib.check(rc, "isp_align_forward")
    return (soft, logits, hard, dur, path, rowsum) if want_rowsum else (soft, logits, hard, dur, path)


# Workspaces of the linked call, one per (device, stream, shape): a workspace that only ever saw successful calls of one shape in
# one stream is left clean by each of them (ISP_ALIGN_WS_CLEAN: no memset in front of the kernels).  9 MB at batch 256 x 1000 x 200
# (isp_mas_workspace_bytes covers every MAS kernel); at most four are kept, none above 32 MB -- the linked kernels stop at 512
# utterances anyway.
_ALIGN_WS: dict = {}


def _align_workspace(lib, dev, stream, B, T1, T2, D, dt):
    key = (dev.index, int(stream), B, T1, T2)
    ws = _ALIGN_WS.get(key)
    if ws is not None:
        return ws, True
    nb = lib.isp_align_workspace_bytes(B, T1, T2, D, dt)
    ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
    if nb <= (32 << 20) and not torch.cuda.is_current_stream_capturing():
        if len(_ALIGN_WS) >= 4:
            _ALIGN_WS.pop(next(iter(_ALIGN_WS)))
        _ALIGN_WS[key] = ws
    return ws, False


def stage_operands(q_host: Tensor, k_host: Tensor, text_len: Tensor, mel_len: Tensor,
                   out_q: Tensor | None = None, out_k: Tensor | None = None):
    """Host -> device copy of the encoded operands, ragged (isp_stage_operands): only the rows below each utterance's
    length cross PCIe, the padding rows of the device tensors are written as zeros.  q_host (B, T1, D), k_host (B, T2, D):
    pinned host tensors (float32 or bfloat16); text_len, mel_len: int64 tensors already on the device.  Enqueued on the
    current stream.  Returns (q_dev, k_dev), ready for loglik_forward."""
    dev = text_len.device
    _lib.require_device(dev)
    lib = _lib.load()
    if q_host.is_cuda or k_host.is_cuda or not q_host.is_pinned() or not k_host.is_pinned():
        raise ValueError("q_host and k_host must be pinned host tensors")
    if q_host.dtype != k_host.dtype or q_host.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("Q and K must both be float32 or both bfloat16")
    if not q_host.is_contiguous() or not k_host.is_contiguous():
        raise ValueError("q_host and k_host must be contiguous")
    B, T1, D = q_host.shape
    T2 = k_host.shape[1]
    if k_host.shape[0] != B or k_host.shape[2] != D:
        raise ValueError(f"shape mismatch: q_host {tuple(q_host.shape)} vs k_host {tuple(k_host.shape)}")
    # the kernel reads the lengths as int64 on the device: coerce like _loglik_cuda does (an int32 tensor would be misread)
    text_len = text_len.to(device=dev, dtype=torch.int64).contiguous()
    mel_len = mel_len.to(device=dev, dtype=torch.int64).contiguous()
    if text_len.numel() != B or mel_len.numel() != B:
        raise ValueError("text_len and mel_len must hold one length per utterance")
    if out_q is None:
        out_q = torch.empty((B, T1, D), dtype=q_host.dtype, device=dev)
    if out_k is None:
        out_k = torch.empty((B, T2, D), dtype=k_host.dtype, device=dev)
    for name, t, shape in (("out_q", out_q, (B, T1, D)), ("out_k", out_k, (B, T2, D))):
        if tuple(t.shape) != shape or t.dtype != q_host.dtype or t.device != dev or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous {q_host.dtype} tensor of shape {shape} on {dev}")
    dt = _lib.ISP_DTYPE_BF16 if q_host.dtype == torch.bfloat16 else _lib.ISP_DTYPE_F32
    with torch.cuda.device(dev):
        rc = lib.isp_stage_operands(q_host.data_ptr(), k_host.data_ptr(), dt, text_len.data_ptr(), mel_len.data_ptr(),
                                    B, T1, T2, D, out_q.data_ptr(), out_k.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "isp_stage_operands")
    return out_q, out_k


def pack_rows(x: Tensor, lengths) -> Tensor:
    """Host-side helper (tests, benchmarks, data loaders): the valid rows of a padded (B, T, D) tensor back to back,
    (sum lengths, D) -- the packed form unpack_operands takes."""
    return torch.cat([x[b, :int(n)] for b, n in enumerate(lengths)], dim=0).contiguous()


def unpack_operands(q_packed: Tensor, k_packed: Tensor, text_len: Tensor, mel_len: Tensor, t1max: int, t2max: int,
                    out_q: Tensor | None = None, out_k: Tensor | None = None):
    """Packed operands ON THE DEVICE (valid rows back to back: (sum mel_len, D) and (sum text_len, D), e.g. the result of one
    `packed_host.to(device, non_blocking=True)` per tensor -- a plain DMA) -> the padded q (B, t1max, D), k (B, t2max, D) that
    loglik_forward loads, padding rows zero (isp_unpack_operands).  Enqueued on the current stream."""
    dev = q_packed.device
    _lib.require_device(dev)
    lib = _lib.load()
    if q_packed.dtype != k_packed.dtype or q_packed.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("q_packed and k_packed must both be float32 or both bfloat16")
    if q_packed.dim() != 2 or k_packed.dim() != 2 or q_packed.shape[1] != k_packed.shape[1] or not q_packed.is_contiguous() or not k_packed.is_contiguous():
        raise ValueError("q_packed (rows, D) and k_packed (rows, D) must be contiguous with the same D")
    text_len = text_len.to(device=dev, dtype=torch.int64).contiguous()
    mel_len = mel_len.to(device=dev, dtype=torch.int64).contiguous()
    B, D = text_len.numel(), q_packed.shape[1]
    if mel_len.numel() != B:
        raise ValueError("text_len and mel_len must hold one length per utterance")
    if out_q is None:
        out_q = torch.empty((B, t1max, D), dtype=q_packed.dtype, device=dev)
    if out_k is None:
        out_k = torch.empty((B, t2max, D), dtype=k_packed.dtype, device=dev)
    for name, t, shape in (("out_q", out_q, (B, t1max, D)), ("out_k", out_k, (B, t2max, D))):
        if tuple(t.shape) != shape or t.dtype != q_packed.dtype or t.device != dev or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous {q_packed.dtype} tensor of shape {shape} on {dev}")
    nb = lib.isp_unpack_workspace_bytes(B)
    ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
    dt = _lib.ISP_DTYPE_BF16 if q_packed.dtype == torch.bfloat16 else _lib.ISP_DTYPE_F32
    with torch.cuda.device(dev):
        rc = lib.isp_unpack_operands(q_packed.data_ptr(), k_packed.data_ptr(), dt, text_len.data_ptr(), mel_len.data_ptr(), B, int(t1max),
                                     int(t2max), D, out_q.data_ptr(), out_k.data_ptr(), ws.data_ptr(), nb, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "isp_unpack_operands")
    return out_q, out_k


def _scores(q: Tensor, k: Tensor, text_len: Tensor | None = None, mel_len: Tensor | None = None) -> Tensor:
    """Unscaled S = Q.K^T in fp32 for the backward pass: the tcgen05 batched GEMM (isp_gemm_batched), both operands K-major.
    Tiles of padded frames / tokens are zero-filled without arithmetic (S is exactly 0 there: the operands' padding is 0)."""
    return bgemm(q, k.transpose(1, 2), m_len=mel_len, n_len=text_len)


def loglik_backward_ds(scores: Tensor, soft: Tensor, g_logits: Tensor | None, g_soft: Tensor | None, scale: float,
                       prior: bool, out_dtype: torch.dtype = torch.float32) -> Tensor:
    """dL/dS from the incoming gradients: the sm_100a kernel behind isp_loglik_backward_ds
    (alignment.py:190-206 differentiated; one pass over the four (B, T1, T2) inputs)."""
    dev = scores.device
    _lib.require_device(dev)
    lib = _lib.load()
    B, T1, T2 = scores.shape
    pad = (-T2) % 4
    def prep(t):
        if t is None:
            return None
        t = t.float()
        if pad:
            t = torch.nn.functional.pad(t, (0, pad))    # the kernel wants 16 B rows; padded scores must not count
        return t.contiguous()
    s_, a_, gl_, gs_ = prep(scores), prep(soft), prep(g_logits), prep(g_soft)
    if pad:
        s_[:, :, T2:] = float("-inf")
    ds = torch.empty((B, T1, T2 + pad), dtype=out_dtype, device=dev)
    dt = _lib.ISP_DTYPE_BF16 if out_dtype == torch.bfloat16 else _lib.ISP_DTYPE_F32
    with torch.cuda.device(dev):
        rc = lib.isp_loglik_backward_ds(s_.data_ptr(), a_.data_ptr(), gl_.data_ptr() if gl_ is not None else None,
                                        gs_.data_ptr() if gs_ is not None else None, B, T1, T2 + pad, float(scale),
                                        1 if prior else 0, ds.data_ptr(), dt, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "isp_loglik_backward_ds")
    return ds[:, :, :T2] if pad else ds


def loglik_backward_from_logits(logits: Tensor, g_logits: Tensor | None, g_soft: Tensor | None, rowsum: Tensor | None,
                                text_len: Tensor, mel_len: Tensor, scale: float, prior: bool,
                                out_dtype: torch.dtype = torch.float32) -> Tensor:
    """dL/dS from the incoming gradients and the forward's own attn_logits (isp_loglik_backward_from_logits): no score GEMM, no
    attn_soft read.  rowsum: the prior's row sums the fused forward kernel saved.  T2max % 4 == 0."""
    dev = logits.device
    _lib.require_device(dev)
    lib = _lib.load()
    B, T1, T2 = logits.shape
    lg = logits.detach().float().contiguous()
    gl = g_logits.float().contiguous() if g_logits is not None else None
    gs = g_soft.float().contiguous() if g_soft is not None else None
    tl = text_len.to(device=dev, dtype=torch.int64).contiguous()
    ml = mel_len.to(device=dev, dtype=torch.int64).contiguous()
    if prior and (rowsum is None or tuple(rowsum.shape) != (B, T1) or rowsum.dtype != torch.float32 or not rowsum.is_contiguous()):
        raise ValueError("rowsum must be the contiguous float32 (B, T1max) tensor the fused forward kernel wrote")
    ds = torch.empty((B, T1, T2), dtype=out_dtype, device=dev)
    dt = _lib.ISP_DTYPE_BF16 if out_dtype == torch.bfloat16 else _lib.ISP_DTYPE_F32
    with torch.cuda.device(dev):
        rc = lib.isp_loglik_backward_from_logits(lg.data_ptr(), gl.data_ptr() if gl is not None else None,
                                                 gs.data_ptr() if gs is not None else None, rowsum.data_ptr() if rowsum is not None else None,
                                                 tl.data_ptr(), ml.data_ptr(), B, T1, T2, float(scale), 1 if prior else 0, ds.data_ptr(), dt,
                                                 torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "isp_loglik_backward_from_logits")
    return ds


class _LogLikelihood(torch.autograd.Function):
    """forward: the fused sm_100a kernel.  backward: three launches of the tcgen05 batched GEMM
    (isp_gemm_batched) around the one-pass Jacobian kernel (isp_loglik_backward_ds) -- scores S = Q.K^T recomputed,
    dS from the incoming gradients, dQ = dS.K (K read MN-major), dK = dS^T.Q (both operands MN-major: no transposed copy).
    No library GEMM anywhere.  The contractions stop at the lengths (K rows >= text_len and Q rows >= mel_len are zero, so
    nothing is lost); every row of dQ / dK is computed, padded ones included, exactly as autograd would."""

    @staticmethod
    def forward(ctx, q, k, text_len, mel_len, scale, prior, with_mas=False, precision="tf32"):
        ctx.scale, ctx.prior, ctx.with_mas, ctx.precision = scale, prior, with_mas, precision
        if with_mas:
            # the linked call (isp_align_forward): the hard path and the durations come with it, outside autograd
            soft, logits, hard, dur, _, rowsum = _align_cuda(q.detach(), k.detach(), text_len, mel_len, scale, prior, want_rowsum=True,
                                                             precision=precision)
        else:
            soft, logits, rowsum = _loglik_cuda(q.detach(), k.detach(), text_len, mel_len, scale, prior, want_rowsum=True, precision=precision)
        # The backward pass works from attn_logits and the prior's row sums when the fused kernel produced them (no score GEMM, no
        # attn_soft read); otherwise (stand-alone row epilogue, token axis not a multiple of 4) from recomputed scores and attn_soft.
        ctx.from_logits = rowsum is not None and logits.shape[2] % 4 == 0
        if ctx.from_logits:
            ctx.save_for_backward(q, k, logits, rowsum, text_len, mel_len)
        else:
            ctx.save_for_backward(q, k, soft, None, text_len, mel_len)
        if with_mas:
            ctx.mark_non_differentiable(hard, dur)
            return soft, logits, hard, dur
        return soft, logits

    @staticmethod
    def backward(ctx, g_soft, g_logits, *_):
        q, k, saved, rowsum, text_len, mel_len = ctx.saved_tensors
        if g_soft is None and g_logits is None:
            return None, None, None, None, None, None, None, None
        qd, kd = q.detach().contiguous(), k.detach().contiguous()
        if ctx.from_logits:
            d_s = loglik_backward_from_logits(saved, g_logits, g_soft, rowsum, text_len, mel_len, ctx.scale, ctx.prior, out_dtype=q.dtype)
        else:
            faithful = ctx.precision == "fp32" and qd.dtype == torch.float32 and qd.shape[2] % 4 == 0
            sc = _scores_fp32(qd, kd, text_len, mel_len) if faithful else _scores(qd, kd, text_len, mel_len)
            d_s = loglik_backward_ds(sc, saved, g_logits, g_soft, ctx.scale, ctx.prior, out_dtype=q.dtype)
        gq = bgemm(d_s, kd, out_dtype=q.dtype, k_len=text_len) if ctx.needs_input_grad[0] else None
        gk = bgemm(d_s.transpose(1, 2), qd, out_dtype=k.dtype, k_len=mel_len) if ctx.needs_input_grad[1] else None
        return gq, gk, None, None, None, None, None, None


def align_forward(q: Tensor, k: Tensor, text_len: Tensor, mel_len: Tensor, scale: float | None = None,
                  attention_prior: bool = True, precision: str = "tf32"):
    """The whole hot path in one linked call (isp_align_forward): (attn_soft, attn_logits, attn_hard int16, durations int64)
    from encoded frames q (B, T1, D) and tokens k (B, T2, D).  attn_soft and attn_logits carry gradients to q and k exactly as
    loglik_forward's do; the hard path and the durations are outside autograd (alignment.py:291 torch.no_grad)."""
    scale = q.shape[-1] ** -0.5 if scale is None else scale
    return _LogLikelihood.apply(q, k, text_len, mel_len, scale, attention_prior, True, precision)


def loglik_forward(q: Tensor, k: Tensor, text_len: Tensor, mel_len: Tensor, scale: float | None = None,
                   attention_prior: bool = True, precision: str = "tf32"):
    """(attn_soft, attn_logits) from encoded frames q (B, T1, D) and tokens k (B, T2, D);
    rows of q / k past each utterance's length must be zero (the projections guarantee it).
    precision, for float32 operands: "tf32" (the fused kernel, one tensor-core product per term) or "fp32" (3xTF32 split: the
    reference's own precision; bfloat16 operands ignore it)."""
    scale = q.shape[-1] ** -0.5 if scale is None else scale
    return _LogLikelihood.apply(q, k, text_len, mel_len, scale, attention_prior, False, precision)


# ----------------------------------------------------------------------------------------------
# modules
# ----------------------------------------------------------------------------------------------
class ConvBlock1D(nn.Module):
    """conv -> activation -> (masked) norm -> dropout, input masked first (alignment.py:40-83)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 1, stride: int = 1,
                 padding: int | None = None, dilation: int = 1, bias: bool = True, activation: str = "relu",
                 normalization: str | None = "batch", dropout_p: float | None = None):
        super().__init__()
        pad = int(dilation * (kernel_size - 1) / 2) if padding is None else padding
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=pad,
                              dilation=dilation, bias=bias and normalization is None)
        self.act = _ACTIVATIONS[str(getattr(activation, "value", activation))]()
        self.norm = _NORMS[str(getattr(normalization, "value", normalization))](num_features=out_channels) \
            if normalization is not None else None
        # the reference crashes on dropout_p=None; None / 0 mean identity here
        self.dropout = nn.Dropout(p=float(dropout_p)) if dropout_p else nn.Identity()

    def forward(self, x: Tensor, input_mask: Tensor | None = None, output_mask: Tensor | None = None) -> Tensor:
        if input_mask is not None:
            x = x * input_mask
        x = self.act(self.conv(x))
        if self.norm is not None:
            x = self.norm(x, mask=output_mask)
        return self.dropout(x)

    def forward_tokens_major(self, x: Tensor, input_mask: Tensor) -> Tensor:
        """Same block for the pointwise, un-normalised last layer, emitting (B, T, C_out):
        a 1x1 convolution is a matmul over channels, and K-major rows are what the GEMM loads."""
        w = self.conv.weight[:, :, 0]
        y = torch.matmul((x * input_mask).transpose(1, 2), w.t())
        if self.conv.bias is not None:
            y = y + self.conv.bias
        return self.dropout(self.act(y))


@dataclass
class ConvAttentionConfig:
    mel_dim: int = MISSING
    text_dim: int = 512
    attention_dim: int = 80
    key_kernel_size: int = 3
    query_kernel_size: int | Sequence[int] = (3, 3)
    dropout: float = 0.0
    normalization: str | None = "instance"
    activation: str = "relu"

    def to_dict(self):
        return dict(self.__dict__)


@dataclass
class AlignerConfig(ConvAttentionConfig):
    ...


class ConvAttention(nn.Module, _ConfigInit):
    def __init__(self, mel_dim: int, text_dim: int = 512, attention_dim: int = 80, key_kernel_size: int = 3,
                 query_kernel_size: int | Sequence[int] = (3, 3), dropout: float = 0.0,
                 normalization: str | None = "instance", activation: str = "relu", attention_prior: bool = True):
        super().__init__()
        self.mel_dim, self.text_dim = mel_dim, text_dim
        self.scale = attention_dim ** -0.5
        self.attention_prior = attention_prior
        #: "auto": bf16 operands under autocast, else "fp32" = what the reference computes outside autocast (torch.matmul in true
        #: fp32, alignment.py:189): fp32 operands, products as the 3xTF32 split on the tensor cores (within ~1e-6 of fp32), the
        #: epilogue as a stand-alone kernel.  "tf32": fp32 operands in memory, ONE tensor-core product per term inside the fused
        #: kernel (10-bit mantissa, fp32 accumulate: within 1e-3 relative of the fp32 reference, ~2x faster than "fp32").  "bf16".
        self.gemm_dtype = "auto"
        #: projection stacks on the sm_100a kernels (stacks.py) when no gradient is needed: "auto" = in bf16 mode (autocast,
        #: the recipe's mixed-precision setting) -- in fp32 mode the torch ops run, because TF32 products through two
        #: convolution layers leave the logits 1e-2 away from the fp32 reference; True = also in fp32 mode; False = never
        self.fused_stacks = "auto"
        #: type of the activations and weights inside the fused stacks in bf16 mode: float16 (10-bit mantissa, what the
        #: reference's fp16 autocast computes its convolutions in, recipes/default.yaml:56) or bfloat16
        self.stack_dtype = torch.float16
        drop = dropout if dropout and dropout > 0.0 else None
        if isinstance(query_kernel_size, int):
            query_kernel_size = [query_kernel_size] * 2

        def stack(spec):
            last = len(spec) - 1
            return nn.ModuleList([
                ConvBlock1D(cin, cout, kernel_size=ks, bias=False, activation=act,
                            normalization=normalization if n < last else None, dropout_p=drop)
                for n, (cin, cout, ks, act) in enumerate(spec)])

        self.key_proj = stack([(text_dim, text_dim * 2, key_kernel_size, activation),
                               (text_dim * 2, attention_dim, 1, "linear")])
        self.query_proj = stack([(mel_dim, mel_dim * 2, query_kernel_size[0], activation),
                                 (mel_dim * 2, mel_dim, query_kernel_size[1], activation),
                                 (mel_dim, attention_dim, 1, "linear")])

    @staticmethod
    def _project(blocks, x: Tensor, mask: Tensor) -> Tensor:
        for blk in blocks[:-1]:
            x = blk(x, input_mask=mask, output_mask=mask)
        last = blocks[-1]
        if last.norm is None and last.conv.kernel_size == (1,):
            return last.forward_tokens_major(x, mask)          # (B, T, D), zero on padded rows
        return last(x, input_mask=mask, output_mask=mask).transpose(1, 2)

    def _mode(self) -> str:
        mode = self.gemm_dtype
        if mode == "auto":
            mode = "bf16" if torch.is_autocast_enabled() else "fp32"
        return mode

    def _needs_autograd(self, *inputs: Tensor) -> bool:
        return torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or any(t.requires_grad for t in inputs))

    def encode(self, queries: Tensor, keys: Tensor, query_len: Tensor, key_len: Tensor):
        """The two projection stacks (alignment.py:176-187) -> q (B, T1, D), k (B, T2, D), zero past the lengths.

        Without gradients (evaluation, alignment extraction) both stacks run on the sm_100a kernels (stacks.py: implicit-GEMM
        convolutions on the tensor cores with the activation and the norm statistics fused, in the log-likelihood's operand
        type).  With gradients enabled, or for block configurations the kernels do not cover (batch norm, other activations,
        channel counts that are not whole 16 B vectors), the torch restatement below runs so that autograd can record it."""
        from . import stacks
        dtype = torch.bfloat16 if self._mode() == "bf16" else torch.float32
        want = self.fused_stacks is True or (self.fused_stacks == "auto" and dtype == torch.bfloat16)
        if (want and queries.is_cuda and not self._needs_autograd(queries, keys)
                and stacks.fused_supported(self.key_proj, dtype) and stacks.fused_supported(self.query_proj, dtype)):   # 2-byte types share a rule
            inner = self.stack_dtype if dtype == torch.bfloat16 else torch.float32
            k = stacks.project_stack(self.key_proj, keys, key_len, self.text_dim, inner, out_dtype=dtype)
            q = stacks.project_stack(self.query_proj, queries, query_len, self.mel_dim, inner, out_dtype=dtype)
            return q, k
        keys = keys.transpose(1, 2) if keys.shape[1] != self.text_dim else keys
        queries = queries.transpose(1, 2) if queries.shape[1] != self.mel_dim else queries
        key_mask = _length_mask(key_len, keys.shape[2]).unsqueeze(1)
        query_mask = _length_mask(query_len, queries.shape[2]).unsqueeze(1)
        k = self._project(self.key_proj, keys, key_mask)
        q = self._project(self.query_proj, queries, query_mask)
        # keep the zero-padding contract of the kernel even when dropout / bias made pads non-zero
        return q * query_mask.transpose(1, 2), k * key_mask.transpose(1, 2)

    def forward(self, queries: Tensor, keys: Tensor, query_len: Tensor, key_len: Tensor):
        """queries (B, mel_dim, T1) mel, keys (B, text_dim, T2) encoded text, lengths (B,)
        -> (attn_soft, attn_logits), both (B, T1, T2) fp32 (alignment.py:159-208)."""
        q, k = self._operands(queries, keys, query_len, key_len)
        return loglik_forward(q, k, key_len, query_len, self.scale, self.attention_prior, precision=self._precision())

    def _precision(self) -> str:
        return "tf32" if self._mode() == "tf32" else "fp32"

    def _operands(self, queries: Tensor, keys: Tensor, query_len: Tensor, key_len: Tensor):
        q, k = self.encode(queries, keys, query_len, key_len)
        if self._mode() == "bf16":
            return q.to(torch.bfloat16), k.to(torch.bfloat16)
        return q.float(), k.float()

    def forward_aligned(self, queries: Tensor, keys: Tensor, query_len: Tensor, key_len: Tensor):
        """forward() plus the MAS hard path and the durations from the same linked call (align_forward):
        (attn_soft, attn_logits, attn_hard, durations)."""
        q, k = self._operands(queries, keys, query_len, key_len)
        return align_forward(q, k, key_len, query_len, self.scale, self.attention_prior, precision=self._precision())


class AlignerOutput(NamedTuple):
    attn_soft: Tensor
    attn_logits: Tensor
    attn_hard: Tensor
    attn_hard_duration: Tensor


class Aligner(nn.Module, _ConfigInit):
    def __init__(self, mel_dim: int, text_dim: int = 512, attention_dim: int = 80, key_kernel_size: int = 3,
                 query_kernel_size: int | Sequence[int] = (3, 3), dropout: float = 0.0,
                 normalization: str | None = "instance", activation: str = "relu", attention_prior: bool = True):
        super().__init__()
        self.attention = ConvAttention(mel_dim=mel_dim, text_dim=text_dim, attention_dim=attention_dim,
                                       key_kernel_size=key_kernel_size, query_kernel_size=query_kernel_size,
                                       dropout=dropout, normalization=normalization, activation=activation,
                                       attention_prior=attention_prior)

    def forward(self, mel: Tensor, enc_text: Tensor, mel_len: Tensor, text_len: Tensor) -> AlignerOutput:
        # alignment.py:253 + :267-275 as one linked call: the MAS kernel starts under the log-likelihood kernel's last wave
        attn_soft, attn_logits, attn_hard, duration = self.attention.forward_aligned(queries=mel, keys=enc_text, query_len=mel_len,
                                                                                    key_len=text_len)
        # every valid frame gets exactly one 1, so durations sum to mel_len by construction and the
        # reference's print-and-patch check (alignment.py:278-282, a device->host sync) never fires
        return AlignerOutput(attn_soft=attn_soft, attn_logits=attn_logits, attn_hard=attn_hard,
                             attn_hard_duration=duration)

    @staticmethod
    @torch.no_grad()
    def _align(attn_logits: Tensor, text_len: Tensor, mel_len: Tensor):
        return mas_forward(attn_logits, text_len, mel_len, durations=True)

    @torch.no_grad()
    def binarize_attention_parallel(self, attn_logits: Tensor, text_len: Tensor, mel_len: Tensor) -> Tensor:
        """MAS hard path, int16 (B, T1max, T2max); no gradient (alignment.py:291-301)."""
        return self._align(attn_logits, text_len, mel_len)[0]

    # the reference's two static routes, kept so external callers keep working; both run the kernel
    @staticmethod
    @torch.no_grad()
    def cuda_binarize_attention_parallel(attn_logits: Tensor, text_len: Tensor, mel_len: Tensor) -> Tensor:
        return mas_forward(attn_logits, text_len, mel_len, durations=False)[0]

    @staticmethod
    @torch.no_grad()
    def cpu_binarize_attention_parallel(attn_logits: Tensor, text_len: Tensor, mel_len: Tensor) -> Tensor:
        if not attn_logits.is_cuda:
            raise _lib.IspError("there is no CPU MAS in this package: move attn_logits to a B200")
        return mas_forward(attn_logits, text_len, mel_len, durations=False)[0]
