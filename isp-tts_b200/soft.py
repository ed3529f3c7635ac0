"""The soft-alignment route of the duration consumers -- the recipe's default (`soft_duration: true`,
recipes/acoustic/core.yaml:148): the TemporalAdaptor passes the Aligner's attn_soft as `alignment`
(tts/models/acoustic/model.py:154 -> temporal_adaptor.py:250-251,300).

  soft_expand(alignment, x)    LengthRegulator, temporal_adaptor.py:417-419:  out = alignment @ x
                               (B, T1, T2) x (B, T2, C): the tcgen05 batched GEMM (isp_gemm_batched), TF32 products on
                               the fp32 alignment as it left the log-likelihood kernel; both gradients are the same
                               kernel on transposed views (no copies).
  soft_average(x, alignment)   TemporalAverager, temporal_adaptor.py:446-449:  x @ alignment / (colsum + 1e-5)
                               one to four feature rows: a stream over the alignment (isp_soft_average), fp32.
"""
from __future__ import annotations

import torch

from . import _lib
from .gemm import bgemm

__all__ = ["soft_expand", "soft_average"]


class _SoftExpand(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alignment, x, frame_len, token_len):
        a = alignment.detach().float()
        xf = x.detach().float()
        out = bgemm(a, xf, m_len=frame_len, k_len=token_len)
        ctx.save_for_backward(a, xf)
        ctx.lens = (frame_len, token_len)
        ctx.x_dtype = x.dtype
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        a, xf = ctx.saved_tensors
        frame_len, token_len = ctx.lens
        g = g.float()
        ga = gx = None
        if ctx.needs_input_grad[0]:      # d alignment = g @ x^T: (B, T1, C) x (B, C, T2)
            ga = bgemm(g, xf.transpose(1, 2), m_len=frame_len, n_len=token_len)
        if ctx.needs_input_grad[1]:      # d x = alignment^T @ g: (B, T2, T1) x (B, T1, C)
            gx = bgemm(a.transpose(1, 2), g, m_len=token_len, k_len=frame_len).to(ctx.x_dtype)
        return ga, gx, None, None


def soft_expand(alignment: torch.Tensor, x: torch.Tensor, frame_len=None, token_len=None) -> torch.Tensor:
    """alignment (B, T1, T2) @ x (B, T2, C) -> (B, T1, C), the reference's `(x.T @ alignment.T).T`.
    frame_len / token_len (optional, (B,)): mel and text lengths; rows / columns of the alignment past them must be zero
    (the Aligner's attn_soft is) -- tiles past them are then skipped instead of multiplied."""
    if alignment.dim() != 3 or x.dim() != 3 or alignment.shape[0] != x.shape[0] or alignment.shape[2] != x.shape[1]:
        raise ValueError(f"alignment (B, T1, T2) and x (B, T2, C) expected, got {tuple(alignment.shape)} and {tuple(x.shape)}")
    _lib.require_device(alignment.device)
    return _SoftExpand.apply(alignment, x, frame_len, token_len)


class _SoftAverage(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alignment, row_len):
        dev = alignment.device
        lib = _lib.load()
        xf = x.detach().float().contiguous()
        a = alignment.detach().float()
        B, C, T1 = xf.shape
        T2 = a.shape[2]
        pad = (-T2) % 4
        if pad or not a.is_contiguous() or a.data_ptr() % 16:
            a = torch.nn.functional.pad(a, (0, pad)).contiguous()
        T2p = T2 + pad
        rl = row_len.to(device=dev, dtype=torch.int64).contiguous() if row_len is not None else None
        out = torch.empty((B, C, T2p), dtype=torch.float32, device=dev)
        colsum = torch.empty((B, T2p), dtype=torch.float32, device=dev)
        ws_bytes = lib.isp_soft_average_workspace_bytes(B, C, T1, T2p)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.isp_soft_average(xf.data_ptr(), a.data_ptr(), rl.data_ptr() if rl is not None else None, out.data_ptr(),
                                      colsum.data_ptr(), B, C, T1, T2p, ws.data_ptr(), ws_bytes, torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "isp_soft_average")
        ctx.save_for_backward(xf, out, colsum)
        ctx.dims = (B, C, T1, T2, T2p)
        return out[:, :, :T2] if pad else out

    @staticmethod
    def backward(ctx, g):
        xf, out, colsum = ctx.saved_tensors
        B, C, T1, T2, T2p = ctx.dims
        if not ctx.needs_input_grad[1]:
            return None, None, None
        dev = g.device
        lib = _lib.load()
        g = g.float()
        if T2p != T2:
            g = torch.nn.functional.pad(g, (0, T2p - T2))
        g = g.contiguous()
        ga = torch.empty((B, T1, T2p), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.isp_soft_average_backward(g.data_ptr(), xf.data_ptr(), out.data_ptr(), colsum.data_ptr(), ga.data_ptr(),
                                               B, C, T1, T2p, torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "isp_soft_average_backward")
        return None, (ga[:, :, :T2] if T2p != T2 else ga), None


def soft_average(x: torch.Tensor, alignment: torch.Tensor, row_len=None) -> torch.Tensor:
    """x (B, C, T1) frame-level features (C <= 4), alignment (B, T1, T2) -> (B, C, T2): x @ alignment / (column sums of the
    alignment + 1e-5).  Differentiable with respect to the alignment (x is a target).  row_len (optional, (B,)): rows of the
    alignment from row_len[b] on are known to be zero and are not read."""
    if x.dim() != 3 or alignment.dim() != 3 or x.shape[0] != alignment.shape[0] or x.shape[2] != alignment.shape[1]:
        raise ValueError(f"x (B, C, T1) and alignment (B, T1, T2) expected, got {tuple(x.shape)} and {tuple(alignment.shape)}")
    _lib.require_device(alignment.device)
    return _SoftAverage.apply(x, alignment, row_len)
