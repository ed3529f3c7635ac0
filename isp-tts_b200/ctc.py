"""Forward-sum (CTC) alignment loss on the attention log-likelihoods.

Reference: AttentionCTCLoss, tts/models/acoustic/loss.py:41-79.  The blank column, the log_softmax, the CTC recursion and
its gradient are three kernels behind isp_ctc_forward / isp_ctc_backward (include/isp_tts_b200.h); the padded
(B, T1, T2 + 1) tensor, its log_softmax and the transposed copy of the reference are never formed.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["ctc_nll", "attention_ctc_loss", "AttentionCTCLoss"]


class _CtcNll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attn_logits, text_len, mel_len, blank_logprob):
        dev = attn_logits.device
        _lib.require_device(dev)
        lib = _lib.load()
        x = attn_logits.detach()
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        B, T1, T2 = x.shape
        tl = text_len.to(device=dev, dtype=torch.int64).contiguous()
        ml = mel_len.to(device=dev, dtype=torch.int64).contiguous()
        nll = torch.empty((B,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws_bytes = lib.isp_ctc_workspace_bytes(B, T1, T2)
            if ws_bytes == 0:
                raise _lib.IspError(f"isp_ctc_forward: shape (B={B}, T1={T1}, T2={T2}) is not covered")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            rc = lib.isp_ctc_forward(x.data_ptr(), tl.data_ptr(), ml.data_ptr(), B, T1, T2, float(blank_logprob),
                                     nll.data_ptr(), ws.data_ptr(), ws_bytes, torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "isp_ctc_forward")
        ctx.save_for_backward(x, tl, ml, nll, ws)
        ctx.blank = float(blank_logprob)
        ctx.in_dtype = attn_logits.dtype
        return nll

    @staticmethod
    def backward(ctx, g):
        x, tl, ml, nll, ws = ctx.saved_tensors
        lib = _lib.load()
        dev = x.device
        B, T1, T2 = x.shape
        gs = g.detach().float().contiguous()
        grad = torch.empty_like(x)
        with torch.cuda.device(dev):
            rc = lib.isp_ctc_backward(x.data_ptr(), tl.data_ptr(), ml.data_ptr(), B, T1, T2, ctx.blank, nll.data_ptr(),
                                      gs.data_ptr(), grad.data_ptr(), ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "isp_ctc_backward")
        return grad.to(ctx.in_dtype), None, None, None


def ctc_nll(attn_logits: torch.Tensor, text_len: torch.Tensor, mel_len: torch.Tensor, blank_logprob: float = -1.0) -> torch.Tensor:
    """Per-utterance negative log-likelihood of the monotonic alignments, (B,) fp32; +inf where mel_len < text_len."""
    if attn_logits.dim() != 3:
        raise ValueError("attn_logits must be (B, T1max, T2max)")
    return _CtcNll.apply(attn_logits, text_len, mel_len, blank_logprob)


def attention_ctc_loss(attn_logits, text_len, mel_len, blank_logprob: float = -1.0) -> torch.Tensor:
    """nn.CTCLoss(zero_infinity=True) of loss.py:73-78: mean over the batch of nll_b / text_len[b], 0 for impossible ones."""
    nll = ctc_nll(attn_logits, text_len, mel_len, blank_logprob)
    nll = torch.where(torch.isinf(nll), torch.zeros_like(nll), nll)
    tl = text_len.to(device=nll.device, dtype=torch.float32).clamp_min(1.0)
    return (nll / tl).mean()


class AttentionCTCLoss(torch.nn.Module):
    """Same call as the reference module (loss.py:41-79) without its WeightedLoss wrapper: forward(attn_logits,
    text_lengths, mel_lengths) -> scalar loss."""

    def __init__(self, blank_logprob: float = -1):
        super().__init__()
        self.blank_logprob = blank_logprob

    def forward(self, attn_logits, text_lengths, mel_lengths, step=None):
        return attention_ctc_loss(attn_logits, text_lengths, mel_lengths, self.blank_logprob)
