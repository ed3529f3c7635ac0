"""Consumers of the alignment that take the path from the MAS kernel.

Reference (paths relative to the reference root):
  LengthRegulator.forward(x, durations)     tts/models/acoustic/modules/temporal_adaptor.py:411-436
  TemporalAverager.forward(x, durations)    tts/models/acoustic/modules/temporal_adaptor.py:439-465
  AttentionBinarizationLoss.forward         tts/models/acoustic/loss.py:97-105   (isp_tts_b200.mas.binarization_loss)
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["length_regulate", "path_from_durations", "LengthRegulator", "temporal_average", "TemporalAverager"]


class _LengthRegulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, path, durations):
        dev = x.device
        _lib.require_device(dev)
        lib = _lib.load()
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("x must be float32 or bfloat16")
        x = x.contiguous()
        B, T2, C = x.shape
        T1 = path.shape[1]
        out = torch.empty((B, T1, C), dtype=x.dtype, device=dev)
        dt = _lib.ISP_DTYPE_BF16 if x.dtype == torch.bfloat16 else _lib.ISP_DTYPE_F32
        with torch.cuda.device(dev):
            rc = lib.isp_length_regulate(x.data_ptr(), path.data_ptr(), out.data_ptr(), dt, B, T1, T2, C,
                                         torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "isp_length_regulate")
        ctx.save_for_backward(durations)
        ctx.shape = (B, T1, T2, C)
        ctx.dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        (durations,) = ctx.saved_tensors
        B, T1, T2, C = ctx.shape
        lib = _lib.load()
        dev = g.device
        g32 = g.float().contiguous()
        dur = _round_durations(durations)              # the same rounding as the forward (temporal_adaptor.py:423)
        starts = (torch.cumsum(dur, dim=1) - dur).contiguous()
        gx = torch.empty((B, T2, C), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.isp_length_regulate_backward(g32.data_ptr(), dur.data_ptr(), starts.data_ptr(), gx.data_ptr(), B, T1, T2, C,
                                                  torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "isp_length_regulate_backward")
        return gx.to(ctx.dtype), None, None


def _round_durations(durations: torch.Tensor) -> torch.Tensor:
    """reps = (durations.float() + 0.5).long(), temporal_adaptor.py:423 (integer durations pass through unchanged)."""
    if durations.dtype in (torch.int64, torch.int32, torch.int16):
        return durations.to(torch.int64).contiguous()
    return (durations.float() + 0.5).long().contiguous()


def path_from_durations(durations: torch.Tensor, t1max: int) -> torch.Tensor:
    """(B, t1max) int16 token index per frame from durations (rounded like the reference does), -1 past each utterance's
    total: what mas_forward(..., return_path=True) returns, for callers that only hold durations (inference)."""
    dev = durations.device
    _lib.require_device(dev)
    lib = _lib.load()
    reps = _round_durations(durations)
    B, T2 = reps.shape
    path = torch.empty((B, int(t1max)), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        rc = lib.isp_path_from_durations(reps.data_ptr(), path.data_ptr(), B, int(t1max), T2, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "isp_path_from_durations")
    return path


def length_regulate(x: torch.Tensor, path: torch.Tensor, durations: torch.Tensor) -> torch.Tensor:
    """x (B, T2, C) expanded to frames: out[b, t] = x[b, path[b, t]] (0 past the utterance).  `path` and `durations`
    are what mas_forward(..., return_path=True) returned for the same batch."""
    if path.dtype != torch.int16 or not path.is_contiguous():
        raise ValueError("path must be the contiguous int16 (B, T1max) tensor returned by mas_forward(..., return_path=True)")
    return _LengthRegulate.apply(x, path, durations)


class LengthRegulator(torch.nn.Module):
    """The reference module's call, argument for argument (temporal_adaptor.py:411-436):
    forward(x, durations, max_len=None, alignment=None) -> (out, dec_lens), plus the optional `path`.

    * `alignment` given (the recipe's soft_duration route, temporal_adaptor.py:300,325): out = alignment @ x, the
      (B, T1, T2) x (B, T2, C) contraction of :418-419, in the sm_100a batched GEMM (soft_expand).
    * else hard durations: a gather along the path.  `path` is what mas_forward(..., return_path=True) returned; without it
      (inference with predicted durations) it is rebuilt on the device from the rounded durations."""

    def forward(self, x, durations, max_len=None, alignment=None, path=None):
        if alignment is not None:
            from .soft import soft_expand
            dec_lens = (durations.sum(dim=1) + 0.5).long()
            out = soft_expand(alignment, x)
        else:
            reps = _round_durations(durations)
            dec_lens = reps.sum(dim=1)
            if path is None:
                t1 = int(dec_lens.max())                       # a host sync, as in the reference (:429)
                path = path_from_durations(reps, t1)
            out = length_regulate(x, path, durations)
        if max_len is not None:
            out = out[:, :max_len]
            dec_lens = torch.clamp_max(dec_lens, max_len)
        return out, dec_lens


def temporal_average(x: torch.Tensor, durations: torch.Tensor) -> torch.Tensor:
    """x (B, C, T1) frame-level features, durations (B, T2) -> (B, C, T2): each token's mean over its frames, counting
    only non-zero values (unvoiced pitch frames are 0), 0 where there is none.  No gradient (the reference feeds targets)."""
    dev = x.device
    _lib.require_device(dev)
    lib = _lib.load()
    if x.dim() != 3 or durations.dim() != 2 or durations.shape[0] != x.shape[0]:
        raise ValueError("x must be (B, C, T1) and durations (B, T2)")
    xf = x.detach().float().contiguous()
    dur = durations.to(device=dev, dtype=torch.int64).contiguous()
    B, C, T1 = xf.shape
    T2 = dur.shape[1]
    out = torch.empty((B, C, T2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.isp_temporal_average(xf.data_ptr(), dur.data_ptr(), out.data_ptr(), B, C, T1, T2,
                                      torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "isp_temporal_average")
    return out


class TemporalAverager(torch.nn.Module):
    """Same call as the reference module (temporal_adaptor.py:439-465).  With hard durations the average is one kernel
    (temporal_average); with a soft `alignment` (the recipe's soft_duration route) it is soft_average: x @ alignment divided
    by the alignment's column sums + 1e-5 (:446-449) in one pass over the alignment."""

    def forward(self, x, durations, alignment=None):
        if alignment is not None:
            from .soft import soft_average
            return soft_average(x, alignment)
        return temporal_average(x, durations)
