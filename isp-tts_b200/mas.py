"""Monotonic Alignment Search entry points, same call shapes as the reference.

Reference (paths relative to the reference root):
  b_mas(b_attn_map, in_lens, out_lens)              tts/modules/aligner/mas.py:30-35
  cuda_b_mas[grid, block](log_p, prev_log_p, prev_ind, attn_out, in_lens, out_lens)
                                                    tts/modules/aligner/cuda_mas.py:11-46
  both called from Aligner.*_binarize_attention_parallel,
                                                    tts/models/acoustic/modules/alignment.py:303-331

Everything here runs the sm_100a kernel behind isp_mas_forward (include/isp_tts_b200.h).
There is no CPU fallback: host (numpy) inputs are staged to the GPU and back.
Unlike the reference `b_mas`, the input is never modified.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

__all__ = ["mas_forward", "b_mas", "cuda_b_mas", "mas_durations", "binarization_loss"]


def _as_len(t, device, name):
    if isinstance(t, torch.Tensor):
        out = t.to(device=device, dtype=torch.int64, non_blocking=True)
    else:
        out = torch.as_tensor(np.asarray(t), dtype=torch.int64).to(device, non_blocking=True)
    if out.dim() != 1:
        raise ValueError(f"{name} must be a 1-D tensor of lengths")
    return out.contiguous()


def mas_forward(attn_logits: torch.Tensor, text_len, mel_len, *, attn_out: torch.Tensor | None = None,
                durations: bool = True, check_lengths: bool = False, return_path: bool = False, dense: bool = True):
    """MAS + hard path + durations in one launch.

    attn_logits (B, T1max, T2max) fp32 CUDA tensor (rows = mel frames); not modified.
    text_len = in_lens, mel_len = out_lens (int64).  Returns (attn_hard int16
    (B, T1max, T2max), durations int64 (B, T2max) or None).  Enqueued on the
    current stream; no host synchronisation unless check_lengths=True.  With return_path=True a third
    value follows: the path as one token index per frame, int16 (B, T1max), -1 past mel_len.  dense=False (with
    return_path=True) skips the dense attn_hard altogether -- the first returned value is then None.
    """
    if attn_logits.dim() != 3:
        raise ValueError("attn_logits must be (B, T1max, T2max)")
    dev = attn_logits.device
    _lib.require_device(dev)
    lib = _lib.load()
    x = attn_logits.detach()
    if x.dtype != torch.float32:
        x = x.float()          # the reference's MAS is fp32 (cuda_mas.py:11)
    if x.stride(2) != 1 or x.stride(1) < x.shape[2] or (x.shape[0] > 1 and x.stride(0) < x.stride(1) * x.shape[1]):
        x = x.contiguous()
    B, T1, T2 = x.shape
    tl = _as_len(text_len, dev, "text_len")
    ml = _as_len(mel_len, dev, "mel_len")
    if tl.numel() != B or ml.numel() != B:
        raise ValueError("text_len / mel_len must have one entry per utterance")
    if not dense:
        if not return_path or attn_out is not None:
            raise ValueError("dense=False needs return_path=True and no attn_out")
        hard = None
    elif attn_out is None:
        hard = torch.empty((B, T1, T2), dtype=torch.int16, device=dev)
    else:
        hard = attn_out
        if hard.shape != (B, T1, T2) or hard.dtype != torch.int16 or not hard.is_contiguous() or hard.device != dev:
            raise ValueError("attn_out must be a contiguous int16 CUDA tensor shaped like attn_logits")
    dur = torch.empty((B, T2), dtype=torch.int64, device=dev) if durations else None
    with torch.cuda.device(dev):
        ws_bytes = lib.isp_mas_workspace_bytes(B, T1, T2)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        path = torch.empty((B, T1), dtype=torch.int16, device=dev) if return_path else None
        if return_path:
            rc = lib.isp_mas_forward_path(x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), tl.data_ptr(), ml.data_ptr(),
                                          B, T1, T2, hard.data_ptr() if hard is not None else None,
                                          dur.data_ptr() if dur is not None else None,
                                          path.data_ptr(), ws.data_ptr(), ws_bytes, stream)
        else:
            rc = lib.isp_mas_forward(x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), tl.data_ptr(), ml.data_ptr(),
                                     B, T1, T2, hard.data_ptr(), dur.data_ptr() if dur is not None else None,
                                     ws.data_ptr(), ws_bytes, stream)
        _lib.check(rc, "isp_mas_forward")
        if check_lengths:
            bad = lib.isp_mas_status(ws.data_ptr(), stream)
            if bad != 0:
                raise _lib.IspError(f"{bad} utterance(s) have a length outside [1, Tmax] (out of contract, SURVEY.md A.6)")
    return (hard, dur, path) if return_path else (hard, dur)


class _BinLoss(torch.autograd.Function):
    """tts/models/acoustic/loss.py:97-105 from the path: forward = one gather of sum(mel_len) floats
    (isp_bin_loss_sums); backward = -1 / (count * soft) on the path's cells, 0 elsewhere."""

    @staticmethod
    def forward(ctx, attn_soft, path, mel_len, eps):
        dev = attn_soft.device
        _lib.require_device(dev)
        lib = _lib.load()
        soft = attn_soft.detach().float().contiguous()
        B, T1, T2 = soft.shape
        ml = _as_len(mel_len, dev, "mel_len")
        sums = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.isp_bin_loss_sums(soft.data_ptr(), path.data_ptr(), ml.data_ptr(), B, T1, T2, float(eps),
                                       sums.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "isp_bin_loss_sums")
        ctx.save_for_backward(soft, path, sums)
        ctx.eps = float(eps)
        return -sums[0] / sums[1]

    @staticmethod
    def backward(ctx, g):
        soft, path, sums = ctx.saved_tensors
        idx = path.clamp(min=0).long().unsqueeze(-1)
        at = soft.gather(2, idx)
        val = torch.where((path.unsqueeze(-1) >= 0) & (at > ctx.eps), -g / (sums[1] * at), torch.zeros_like(at))
        return torch.zeros_like(soft).scatter_(2, idx, val), None, None, None


def binarization_loss(attn_soft: torch.Tensor, path: torch.Tensor, mel_len, eps: float = 1e-6) -> torch.Tensor:
    """Drop-in value for AttentionBinarizationLoss.forward(soft_attention, hard_attention) (loss.py:97-105), taking the
    path (mas_forward(..., return_path=True)) instead of the dense hard attention."""
    if path.dtype != torch.int16 or path.shape != attn_soft.shape[:2] or not path.is_contiguous():
        raise ValueError("path must be the contiguous int16 (B, T1max) tensor returned by mas_forward(..., return_path=True)")
    return _BinLoss.apply(attn_soft, path, mel_len, eps)


def mas_durations(attn_logits, text_len, mel_len):
    """Durations only (int64 (B, T2max)); the dense path is still produced internally."""
    return mas_forward(attn_logits, text_len, mel_len)[1]


def b_mas(b_attn_map, in_lens, out_lens, *, device=None):
    """Drop-in for the reference `b_mas` (mas.py:30-35).

    numpy in -> numpy int16 out, as the reference; torch CUDA tensors in -> torch out.
    Host arrays are copied to `device` (default: current CUDA device), aligned there,
    and the int16 result is copied back -- the reference's CPU route does the mirror
    image of this (alignment.py:307-312).
    """
    if isinstance(b_attn_map, torch.Tensor) and b_attn_map.is_cuda:
        return mas_forward(b_attn_map, in_lens, out_lens, durations=False)[0]
    if not torch.cuda.is_available():
        raise _lib.IspError("b_mas needs a B200: no CUDA device is visible and there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    host = torch.as_tensor(np.ascontiguousarray(b_attn_map, dtype=np.float32)) if not isinstance(b_attn_map, torch.Tensor) \
        else b_attn_map.to(torch.float32).contiguous()
    x = host.to(dev, non_blocking=True)
    hard, _ = mas_forward(x, in_lens, out_lens, durations=False)
    out = hard.cpu()
    return out if isinstance(b_attn_map, torch.Tensor) else out.numpy()


class _Launcher:
    """What `cuda_b_mas[grid, block]` evaluates to: call it with the reference's six arguments."""

    def __init__(self, grid=None, block=None):
        self.grid, self.block = grid, block      # accepted and ignored: the kernel sizes itself

    def __call__(self, log_p, prev_log_p, prev_ind, attn_out, in_lens, out_lens):
        # prev_log_p / prev_ind are the reference kernel's global scratch (alignment.py:325-326);
        # this kernel keeps that state in registers / shared memory, so they are untouched.
        if not (isinstance(log_p, torch.Tensor) and log_p.is_cuda):
            raise _lib.IspError("cuda_b_mas expects torch CUDA tensors (alignment.py:321-330)")
        mas_forward(log_p, in_lens, out_lens, attn_out=attn_out, durations=False)
        return None


class _CudaBMas:
    """Keeps numba's launch syntax working: cuda_b_mas[(grid), (block)](...) (alignment.py:328-330)."""

    def __getitem__(self, cfg):
        if isinstance(cfg, tuple) and len(cfg) >= 2:
            return _Launcher(cfg[0], cfg[1])
        return _Launcher(cfg, None)

    def __call__(self, *args):
        return _Launcher()(*args)


cuda_b_mas = _CudaBMas()
