"""CPU restatement of the reference's forward-sum (CTC) alignment loss -- TEST INFRASTRUCTURE, never on the product path.

Follows tts/models/acoustic/loss.py:41-79 (AttentionCTCLoss) operation by operation, with torch on the CPU:
  :53-57  get_target_seqs      targets 1 .. text_len[b], 0 past it
  :67     F.pad(attn_logits, (1, 0), value=blank_logprob)
  :69-70  log_softmax(dim=2), transpose(0, 1)
  :73-78  nn.CTCLoss(zero_infinity=True)  (reduction 'mean': nll_b / text_len[b], averaged over the batch)
The arithmetic itself lives in torch (torch==2.5.1 in the reference's requirements.txt:1; 2.11 here); float64 by default so
that the oracle is exact to the tolerance the tests state.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def get_target_seqs(lengths: torch.Tensor) -> torch.Tensor:          # loss.py:53-57
    ids = torch.arange(1, int(lengths.max()) + 1, device=lengths.device)
    ids = ids[None].expand(lengths.numel(), -1).clone()
    ids[ids > lengths.unsqueeze(1)] = 0
    return ids


def attention_ctc_loss(attn_logits: torch.Tensor, text_lengths: torch.Tensor, mel_lengths: torch.Tensor,
                       blank_logprob: float = -1.0, dtype=torch.float64, reduction: str = "mean") -> torch.Tensor:
    """loss.py:59-79 on CPU.  attn_logits (B, T1, T2); returns the scalar loss (or per-utterance nll with reduction='none')."""
    x = attn_logits.to(dtype)
    padded = F.pad(input=x, pad=(1, 0), value=blank_logprob)          # :67
    logprob = F.log_softmax(padded, dim=2).transpose(0, 1)            # :69-70
    targets = get_target_seqs(text_lengths)
    return F.ctc_loss(logprob, targets, mel_lengths, text_lengths, blank=0, reduction=reduction, zero_infinity=True)   # :73-78
