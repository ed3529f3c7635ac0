"""Seeded synthetic LJSpeech-shaped batches (numpy, host side).

The workloads are the ones BASELINE.json lists.  There
is no dataset in the image, so lengths follow the LJSpeech-like law below and
values are random.  numpy's legacy RandomState is used because its streams are
stable across numpy versions, so a seed names the same batch on every box.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Workload:
    name: str
    batch: int
    t2max: int          # text tokens
    t1max: int          # mel frames
    ragged: bool
    seed: int
    dim: int = 128      # attention_dim, recipes/acoustic/core.yaml:151 in the reference


# BASELINE.json `configs`, in order.
WORKLOADS = {
    "cfg1": Workload("cfg1: single utterance 80 tok x 400 fr", 1, 80, 400, False, 1235),
    "cfg2": Workload("cfg2: LJSpeech-shaped batch 32, <=200 tok x <=1000 fr", 32, 200, 1000, True, 1236),
    "cfg3": Workload("cfg3: batch 256 ragged LJSpeech-shaped, <=200 tok x <=1000 fr", 256, 200, 1000, True, 1237),
    "cfg3d": Workload("cfg3d: batch 256 dense 200 tok x 1000 fr", 256, 200, 1000, False, 1237),
    "cfg4": Workload("cfg4: long-form batch 16, 512 tok x 4096 fr", 16, 512, 4096, False, 1238),
    # BASELINE.json configs[4]: ONE global batch (64..4096 utterances, --batch) sharded by utterance across the ranks
    "cfg5": Workload("cfg5: one global batch of 1024 ragged utterances sharded across the ranks, <=200 tok x <=1000 fr", 1024, 200, 1000, True, 1239),
}


def lengths(batch: int, t2max: int, t1max: int, ragged: bool, seed: int):
    """-> (text_len, mel_len) int64 arrays.  One utterance is forced to
    (t2max, t1max) so the padded shape is always the nominal one."""
    if not ragged:
        return (np.full(batch, t2max, dtype=np.int64), np.full(batch, t1max, dtype=np.int64))
    rs = np.random.RandomState(seed)
    lo = min(20, t2max)
    text = np.clip(np.rint(rs.normal(0.55 * t2max, 0.2 * t2max, size=batch)), lo, t2max).astype(np.int64)
    ratio = rs.uniform(4.5, 6.5, size=batch)
    mel = np.clip(np.rint(text * ratio), text, t1max).astype(np.int64)
    k = int(rs.randint(batch))
    text[k], mel[k] = t2max, t1max
    return text, mel


def workload_lengths(w: Workload, batch: int | None = None):
    return lengths(w.batch if batch is None else batch, w.t2max, w.t1max, w.ragged, w.seed)


def noise_logits(batch: int, t1max: int, t2max: int, seed: int, quantize: float = 0.0):
    """N(0,1) fp32 (B, T1, T2); `quantize` > 0 rounds to that step so that
    finite ties are frequent (exercises the tie rule, mas.py:17)."""
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((batch, t1max, t2max)).astype(np.float32)
    if quantize > 0:
        x = (np.rint(x / quantize) * quantize).astype(np.float32)
    return x


def encoded_pair(batch: int, t1max: int, t2max: int, dim: int, text_len, mel_len, seed: int,
                 std: float = 0.57):
    """Stand-ins for the ConvAttention projections: Q (B, T1, D), K (B, T2, D)
    fp32 with std ~0.57 and exact zeros at padded positions
    (alignment.py:75-76 guarantees that in the reference)."""
    rs = np.random.RandomState(seed)
    q = (rs.standard_normal((batch, t1max, dim)) * std).astype(np.float32)
    k = (rs.standard_normal((batch, t2max, dim)) * std).astype(np.float32)
    q[np.arange(t1max)[None, :] >= np.asarray(mel_len)[:, None]] = 0.0
    k[np.arange(t2max)[None, :] >= np.asarray(text_len)[:, None]] = 0.0
    return q, k


# state_dict of the reference Aligner at the recipe's hyper-parameters (recipes/acoustic/core.yaml:150-156 with mel 80,
# text 384: SURVEY.md A.8): 11 tensors, 1,713,120 parameters
RECIPE_HP = dict(mel_dim=80, text_dim=384, attention_dim=128, key_kernel_size=5, query_kernel_size=[5, 5],
                 dropout=0.1, normalization="instance", activation="gelu")
RECIPE_SHAPES = {
    "attention.key_proj.0.conv.weight": (768, 384, 5), "attention.key_proj.0.norm.weight": (768,), "attention.key_proj.0.norm.bias": (768,),
    "attention.key_proj.1.conv.weight": (128, 768, 1),
    "attention.query_proj.0.conv.weight": (160, 80, 5), "attention.query_proj.0.norm.weight": (160,), "attention.query_proj.0.norm.bias": (160,),
    "attention.query_proj.1.conv.weight": (80, 160, 5), "attention.query_proj.1.norm.weight": (80,), "attention.query_proj.1.norm.bias": (80,),
    "attention.query_proj.2.conv.weight": (128, 80, 1),
}


def recipe_state(seed: int):
    """Seeded weights of the recipe-shape Aligner (numpy): conv weights ~ U(-b, b) with b = 1/sqrt(fan_in) like torch's
    default init, norm gains around 1, norm biases around 0.  The same seed gives the same weights on every box, so a
    golden file only has to hold the reference's OUTPUTS for them."""
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in RECIPE_SHAPES.items():
        if name.endswith("conv.weight"):
            bound = 1.0 / np.sqrt(shape[1] * shape[2])
            out[name] = rs.uniform(-bound, bound, size=shape).astype(np.float32)
        elif name.endswith("norm.weight"):
            out[name] = (1.0 + 0.1 * rs.standard_normal(shape)).astype(np.float32)
        else:
            out[name] = (0.1 * rs.standard_normal(shape)).astype(np.float32)
    return out


def recipe_inputs(seed: int, B: int, T1: int, T2: int, text_len, mel_len):
    """mel (B, 80, T1) log-mel-like and enc_text (B, 384, T2), zero at padded positions (collator.py:47, transformer.py:206)."""
    rs = np.random.RandomState(seed)
    mel = np.clip(rs.standard_normal((B, 80, T1)) * 2 - 5, -11.5, 2.0).astype(np.float32)
    txt = rs.standard_normal((B, 384, T2)).astype(np.float32)
    mel *= (np.arange(T1)[None, None] < np.asarray(mel_len)[:, None, None])
    txt *= (np.arange(T2)[None, None] < np.asarray(text_len)[:, None, None])
    return mel, txt
