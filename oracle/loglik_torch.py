"""CPU port of the reference's log-likelihood ops in torch (multi-threaded) -- TEST /
BASELINE INFRASTRUCTURE ONLY.  Same op sequence the reference itself executes when it runs
on CPU (tts/models/acoustic/modules/alignment.py:18-37, 189-208); used by bench.py as the
timed CPU baseline together with oracle/mas_oracle.c.  Never imported by the product."""
from __future__ import annotations

import torch


def _mask(lengths, max_len):
    return torch.arange(max_len)[None, :] < lengths[:, None]            # functions.py:61-66


def batch_diagonal_prior(text_lengths, mel_lengths, t2max, t1max, gamma=0.1, threshold=1e-4):
    gt = torch.arange(t2max, dtype=torch.float32).view(1, -1) / text_lengths.view(-1, 1)      # :21-22
    gm = torch.arange(t1max, dtype=torch.float32).view(1, -1) / mel_lengths.view(-1, 1)       # :24-25
    grid = gt.unsqueeze(1) - gm.unsqueeze(2)                                                  # :27
    prior = torch.exp(-grid ** 2 / (2 * gamma ** 2))                                          # :29
    prior.transpose(2, 1)[~_mask(text_lengths, t2max)] = 0.0                                   # :31
    prior[~_mask(mel_lengths, t1max)] = 0.0                                                    # :32
    prior = prior / (prior.sum(dim=-1, keepdim=True) + 1e-5)                                   # :34
    return prior.masked_fill(prior < threshold, 0.0)                                           # :35


@torch.no_grad()
def loglik(q, k, text_len, mel_len, scale):
    """q (B,T1,D), k (B,T2,D) fp32 CPU tensors -> (attn_soft, attn_logits)."""
    T1, T2 = q.shape[1], k.shape[1]
    key_mask = _mask(text_len, T2).unsqueeze(1)
    query_mask = _mask(mel_len, T1).unsqueeze(1)
    mask = query_mask.transpose(1, 2) & key_mask                                               # :176-178
    attn = torch.matmul(q, k.transpose(1, 2))                                                  # :189
    attn = scale * attn                                                                        # :190
    attn = torch.clamp(attn, max=3.4028234663852886e38)                                        # :192
    prior = batch_diagonal_prior(text_len, mel_len, T2, T1)                                    # :195
    attn = torch.log_softmax(attn, dim=2, dtype=torch.float32) + torch.log(prior + 1e-6)       # :196
    attn_logits = attn.clone()                                                                 # :198
    attn.masked_fill_(~mask[:, :1], -3.4028234663852886e38)                                    # :201
    attn = torch.softmax(attn, dim=2, dtype=torch.float32)                                     # :203
    return attn * mask, attn_logits                                                            # :206-208
