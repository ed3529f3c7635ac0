"""CPU restatement (numpy, float64) of the reference's duration / alignment consumers.  TEST INFRASTRUCTURE ONLY: imported
by tests/ (and nothing in the product).  Pinned by tests/golden/consumers.npz, which holds outputs of the reference's own
modules (oracle/gen_golden.py::gen_consumers).

Reference: tts/models/acoustic/modules/temporal_adaptor.py
  LengthRegulator.forward   :411-436     TemporalAverager.forward   :439-465
"""
from __future__ import annotations

import numpy as np


def round_durations(durations):
    """reps = (durations.float() + 0.5).long()   (:423)"""
    return np.floor(np.asarray(durations, dtype=np.float32) + np.float32(0.5)).astype(np.int64)


def length_regulate_hard(x, durations, max_len=None):
    """:422-431 -- out[b, t] = x[b, j] for the token j whose cumulated-duration interval holds frame t."""
    x = np.asarray(x, dtype=np.float64)
    reps = round_durations(durations)
    dec = reps.sum(1)
    B, T2, C = x.shape
    out = np.zeros((B, int(dec.max()), C))
    for b in range(B):
        idx = np.repeat(np.arange(T2), reps[b])
        out[b, :len(idx)] = x[b, idx]
    if max_len is not None:                                    # :433-435
        out = out[:, :max_len]
        dec = np.minimum(dec, max_len)
    return out, dec


def length_regulate_soft(x, durations, alignment, max_len=None):
    """:417-419 -- dec_lens = (sum durations + 0.5).long(); out = (x^T @ alignment^T)^T = alignment @ x."""
    dec = np.floor(np.asarray(durations, dtype=np.float64).sum(1) + 0.5).astype(np.int64)
    out = np.einsum("btj,bjc->btc", np.asarray(alignment, dtype=np.float64), np.asarray(x, dtype=np.float64))
    if max_len is not None:
        out = out[:, :max_len]
        dec = np.minimum(dec, max_len)
    return out, dec


def temporal_average_hard(x, durations):
    """:451-465 -- per-token sum of x over the token's frames / number of non-zero x among them (0 when none)."""
    x = np.asarray(x, dtype=np.float64)
    d = np.asarray(durations, dtype=np.int64)
    B, C, T1 = x.shape
    out = np.zeros((B, C, d.shape[1]))
    for b in range(B):
        ends = np.cumsum(d[b])
        starts = ends - d[b]
        for j in range(d.shape[1]):
            seg = x[b, :, starts[j]:ends[j]]
            n = (seg != 0.0).sum(1)
            out[b, :, j] = np.where(n == 0, 0.0, seg.sum(1) / np.maximum(n, 1))
    return out


def temporal_average_soft(x, alignment):
    """:446-449 -- x @ alignment / (alignment.sum(dim=1) + 1e-5)."""
    a = np.asarray(alignment, dtype=np.float64)
    return np.einsum("bct,btj->bcj", np.asarray(x, dtype=np.float64), a) / (a.sum(1, keepdims=True) + 1e-5)
