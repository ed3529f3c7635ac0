"""ctypes front-end of oracle/mas_oracle.c -- TEST INFRASTRUCTURE ONLY.

Same call shape as the reference's `b_mas`
(/root/reference/tts/modules/aligner/mas.py:30-35): numpy fp32 (B, T1, T2),
text lengths, mel lengths -> int16 (B, T1, T2).  Unlike the reference it does
not mutate its input.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmas_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile mas_oracle.c with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "mas_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        lib.oracle_b_mas.argtypes = [p, i64, i64, i64, p, p, p, p, ctypes.c_int]
        lib.oracle_b_mas.restype = ctypes.c_int
        lib.oracle_mas_accumulate.argtypes = [p, i64, i64, i64, p]
        lib.oracle_mas_accumulate.restype = None
        lib.oracle_num_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().oracle_num_threads())


def b_mas_with_durations(b_attn_map, in_lens, out_lens, nthreads: int = 0):
    """-> (attn_hard int16 (B,T1,T2), durations int64 (B,T2))."""
    lib = _load()
    x = np.ascontiguousarray(b_attn_map, dtype=np.float32)
    if x.ndim != 3:
        raise ValueError("b_attn_map must be (B, T1, T2)")
    B, T1, T2 = x.shape
    il = np.ascontiguousarray(in_lens, dtype=np.int64)
    ol = np.ascontiguousarray(out_lens, dtype=np.int64)
    out = np.empty((B, T1, T2), dtype=np.int16)
    dur = np.empty((B, T2), dtype=np.int64)
    rc = lib.oracle_b_mas(x.ctypes.data, B, T1, T2, il.ctypes.data, ol.ctypes.data,
                          out.ctypes.data, dur.ctypes.data, int(nthreads))
    if rc != 0:
        raise ValueError(f"oracle_b_mas failed (rc={rc}): lengths must be in [1, T]")
    return out, dur


def b_mas(b_attn_map, in_lens, out_lens, nthreads: int = 0):
    return b_mas_with_durations(b_attn_map, in_lens, out_lens, nthreads)[0]


def accumulate(x2d):
    """Accumulated Q (n, m) of one utterance (reference: mas.py:11-14)."""
    lib = _load()
    x = np.ascontiguousarray(x2d, dtype=np.float32)
    n, m = x.shape
    q = np.empty((n, m), dtype=np.float32)
    lib.oracle_mas_accumulate(x.ctypes.data, m, n, m, q.ctypes.data)
    return q


def path_from_hard(attn_hard, out_lens):
    """Token index per frame, -1 on padded frames: (B, T1) int16."""
    a = np.asarray(attn_hard)
    p = a.argmax(axis=2).astype(np.int16)
    for b, n in enumerate(np.asarray(out_lens)):
        p[b, int(n):] = -1
    return p
