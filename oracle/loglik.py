"""numpy restatement of the reference's log-likelihood path -- TEST INFRASTRUCTURE ONLY.

Follows, op for op and in fp32 like the reference:
  batch_diagonal_prior   <- tts/models/acoustic/modules/alignment.py:18-37
  loglik                 <- ConvAttention.forward, alignment.py:176-178 (masks),
                            :189-192 (matmul, scale, clamp), :194-196
                            (log_softmax over ALL T2max columns + log prior),
                            :198-208 (attn_logits clone, masked softmax, mask)

Parity status: PINNED against tests/golden/loglik_*.npz, which
oracle/gen_golden.py produced by running the reference's own
`ConvAttention.forward` (loaded by file path) on its own captured Q/K.
The product package never imports this module.
"""
from __future__ import annotations

import numpy as np

F32_MAX = np.float32(3.4028234663852886e38)   # tts/utils/functions.py:36-41
F32_MIN = np.float32(-3.4028234663852886e38)  # tts/utils/functions.py:28-33
PRIOR_EPS = np.float32(1e-6)                  # alignment.py:196
LOG_PRIOR_FLOOR = float(np.log(np.float32(1e-6)))


def mask_from_lengths(lengths, max_len):
    """get_mask_from_lengths, tts/utils/functions.py:61-66."""
    ids = np.arange(int(max_len))
    return ids[None, :] < np.asarray(lengths)[:, None]


def batch_diagonal_prior(text_lengths, mel_lengths, gamma=0.1, threshold=1e-4,
                         t2max=None, t1max=None, return_unthresholded=False):
    """(B, T1max, T2max) fp32.  alignment.py:18-37."""
    tl = np.asarray(text_lengths)
    ml = np.asarray(mel_lengths)
    t2max = int(tl.max()) if t2max is None else int(t2max)
    t1max = int(ml.max()) if t1max is None else int(t1max)
    f32 = np.float32
    grid_text = np.arange(t2max, dtype=f32)[None, :] / tl.astype(f32)[:, None]   # :21-22
    grid_mel = np.arange(t1max, dtype=f32)[None, :] / ml.astype(f32)[:, None]    # :24-25
    grid = grid_text[:, None, :] - grid_mel[:, :, None]                           # :27
    prior = np.exp(-(grid * grid) / f32(2 * gamma ** 2)).astype(f32)              # :29
    prior[~np.broadcast_to(mask_from_lengths(tl, t2max)[:, None, :], prior.shape)] = 0.0  # :31
    prior[~mask_from_lengths(ml, t1max)] = 0.0                                    # :32
    prior = prior / (prior.sum(axis=-1, keepdims=True, dtype=f32) + f32(1e-5))    # :34
    raw = prior
    prior = np.where(prior < f32(threshold), f32(0.0), prior).astype(f32)         # :35
    if return_unthresholded:
        return prior, raw
    return prior


def log_softmax(x, axis):
    m = x.max(axis=axis, keepdims=True)
    s = x - m
    return (s - np.log(np.exp(s).sum(axis=axis, keepdims=True, dtype=np.float32))).astype(np.float32)


def loglik(Q, K, text_len, mel_len, scale=None, attention_prior=True, return_parts=False):
    """Q: (B, T1max, D), K: (B, T2max, D) -- i.e. queries_enc^T and keys_enc^T
    of alignment.py:182,187 -- fp32 (or anything castable).  Returns
    (attn_soft, attn_logits), both (B, T1max, T2max) fp32, as alignment.py:208.
    """
    f32 = np.float32
    Q = np.asarray(Q, dtype=np.float64)
    K = np.asarray(K, dtype=np.float64)
    B, T1, D = Q.shape
    T2 = K.shape[1]
    tl = np.asarray(text_len)
    ml = np.asarray(mel_len)
    scale = D ** -0.5 if scale is None else scale                  # :116
    # :189 (fp32 GEMM in the reference; accumulated in fp64 here and rounded once)
    S = np.matmul(Q, K.transpose(0, 2, 1)).astype(f32)
    S = (f32(scale) * S).astype(f32)                               # :190
    S = np.minimum(S, F32_MAX)                                     # :192
    key_mask = mask_from_lengths(tl, T2)[:, None, :]               # :176
    query_mask = mask_from_lengths(ml, T1)[:, :, None]             # :177
    mask = query_mask & key_mask                                   # :178
    prior = prior_raw = None
    if attention_prior:
        prior, prior_raw = batch_diagonal_prior(tl, ml, t2max=T2, t1max=T1, return_unthresholded=True)   # :195
        attn = log_softmax(S, axis=2) + np.log(prior + PRIOR_EPS).astype(f32)  # :196
    else:
        attn = S
    attn = attn.astype(f32)
    attn_logits = attn.copy()                                      # :198
    attn = np.where(np.broadcast_to(key_mask, attn.shape), attn, F32_MIN)  # :201
    m = attn.max(axis=2, keepdims=True)
    e = np.exp(attn - m).astype(f32)
    soft = (e / e.sum(axis=2, keepdims=True, dtype=f32)).astype(f32)       # :203
    soft = (soft * mask).astype(f32)                               # :206
    if return_parts:
        return soft, attn_logits, dict(S=S, prior=prior, prior_raw=prior_raw, mask=mask)
    return soft, attn_logits


def threshold_ambiguous(prior_raw, threshold=1e-4, rel=2e-5):
    """Cells whose normalised prior lies within `rel` of the hard threshold of alignment.py:35.
    The reference zeroes a prior below 1e-4, so an implementation that differs from it by one
    rounding in exp / sum can land on the other side there (its own CPU and CUDA builds do);
    parity tests exclude these cells and bound how many there are."""
    t = np.float32(threshold)
    return np.abs(prior_raw - t) <= np.float32(rel) * t


def durations_from_hard(attn_hard):
    """attn_hard.sum(dim=1) -> int64 (B, T2).  alignment.py:275."""
    return np.asarray(attn_hard).sum(axis=1, dtype=np.int64)
