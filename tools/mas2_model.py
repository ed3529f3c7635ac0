"""Executable model (numpy, host) of the data layouts behind isp_mas2.cu -- the index arithmetic of the
backtrack, stage by stage, so that it can be checked against the oracle without a GPU:

  chain     lane L owns columns 2L, 2L+1; at local step u it is on row u - (L & 31); per 16-step chunk it
            emits one word W[ch][L], bit 2k+c = backpointer of (row 16ch+k-l, column 2L+c)
  transpose 32 rows x 64 columns at a time: W (bit index = step) -> rm[row][word] (bit index = column)
  map       per group of 32 rows, bit-sliced over columns: plane k of "which column does the path through
            (bottom row, column j) come from at the top of the group"
  hop       the path's column at every group boundary, from (n-1, m-1) upwards
  expand    every group walks its own 32 rows from its entry column (lanes in parallel on the GPU)

tests/test_mas2_model.py runs this against the C oracle.  Test infrastructure only: nothing in the
product imports it.
"""
from __future__ import annotations

import numpy as np

U32 = 0xFFFFFFFF


def forward_bits(x: np.ndarray) -> np.ndarray:
    """mas.py:11-17 on one utterance window x (n, m) fp32 -> b (n, m) uint8, b[i, j] = 1 iff the
    predecessor of (i, j) is (i-1, j-1).  Row 0 is left 0."""
    n, m = x.shape
    q = np.full(m, -np.inf, dtype=np.float32)
    q[0] = x[0, 0]
    b = np.zeros((n, m), dtype=np.uint8)
    for i in range(1, n):
        left = np.concatenate(([np.float32(-np.inf)], q[:-1]))
        with np.errstate(invalid="ignore"):
            b[i] = left >= q                      # ties and -inf >= -inf -> diagonal
        b[i, 0] = 0
        q = (x[i] + np.maximum(left, q)).astype(np.float32)
    return b


def chain_words(b: np.ndarray, rng=None) -> np.ndarray:
    """The words the strip warps store: W[ch, L]; cells outside the window hold garbage (random)."""
    n, m = b.shape
    nl = (m + 1) // 2
    ns = (nl + 31) // 32
    nch = (n + 31 + 15) // 16
    rng = rng or np.random.RandomState(0)
    w = np.zeros((nch + 2, ns * 32), dtype=np.uint64)
    garbage = rng.randint(0, 2, size=(nch + 2, ns * 32, 32)).astype(np.uint64)
    for L in range(ns * 32):
        l = L & 31
        for ch in range(nch + 2):
            word = 0
            for k in range(16):
                r = 16 * ch + k - l
                for c in range(2):
                    j = 2 * L + c
                    bit = int(b[r, j]) if (0 <= r < n and j < m) else int(garbage[ch, L, 2 * k + c])
                    word |= bit << (2 * k + c)
            w[ch, L] = word
    return w.astype(np.uint32)


def transpose_block(w: np.ndarray, g: int, s: int, n: int):
    """Rows 32g..32g+31 of strip s -> 32 pairs of row-major words (columns 64s..64s+63)."""
    xs = []
    for i in range(32):
        L = 32 * s + i
        p0 = 64 * g + 2 * i
        wi, sh = p0 >> 5, p0 & 31
        stream = int(w[wi, L]) | (int(w[wi + 1, L]) << 32) | (int(w[wi + 2, L]) << 64)
        xs.append((stream >> sh) & 0xFFFFFFFFFFFFFFFF)
    out = []
    for t in range(32):
        row = 32 * g + t
        v = 0
        if 1 <= row < n:
            for i in range(32):
                v |= ((xs[i] >> (2 * t)) & 3) << (2 * i)
        out.append((v & U32, v >> 32))
    return out


def row_major(w: np.ndarray, n: int, m: int) -> np.ndarray:
    nl = (m + 1) // 2
    ns = (nl + 31) // 32
    nblk = (n + 31) // 32
    rm = np.zeros((nblk * 32, 2 * ns), dtype=np.uint32)
    for g in range(nblk):
        for s in range(ns):
            for t, (lo, hi) in enumerate(transpose_block(w, g, s, n)):
                rm[32 * g + t, 2 * s], rm[32 * g + t, 2 * s + 1] = lo, hi
    return rm


def group_map(rm: np.ndarray, g: int, n: int, planes: int) -> np.ndarray:
    nw = rm.shape[1]
    P = np.zeros((planes, nw), dtype=np.uint64)
    for k in range(planes):
        for wd in range(nw):
            v = 0
            for bb in range(32):
                v |= (((32 * wd + bb) >> k) & 1) << bb
            P[k, wd] = v
    for i in range(max(32 * g, 1), min(32 * g + 31, n - 1) + 1):
        for k in range(planes):
            for wd in range(nw - 1, -1, -1):
                a = int(rm[i, wd])
                below = int(P[k, wd - 1]) >> 31 if wd > 0 else 0
                sft = ((int(P[k, wd]) << 1) | below) & U32
                P[k, wd] = (a & sft) | (~a & U32 & int(P[k, wd]))
    return P.astype(np.uint32)


def bit_from_words(w: np.ndarray, r: int, j: int) -> int:
    """Backpointer of (r, j) read from the chain's own layout (what the expansion does)."""
    L = j >> 1
    u = r + (L & 31)
    return (int(w[u >> 4, L]) >> (2 * (u & 15) + (j & 1))) & 1


def backtrack(w: np.ndarray, n: int, m: int) -> np.ndarray:
    planes = max(1, int(m - 1).bit_length())
    rm = row_major(w, n, m)
    G = (n - 1) >> 5
    maps = [group_map(rm, g, n, planes) for g in range(G + 1)]
    entry = [0] * (G + 1)
    j = m - 1
    for g in range(G, -1, -1):
        entry[g] = j
        j = sum(((int(maps[g][k, j >> 5]) >> (j & 31)) & 1) << k for k in range(planes))
    path0 = j
    path = np.full(n, -1, dtype=np.int64)
    for g in range(G + 1):
        j = entry[g]
        for i in range(min(32 * g + 31, n - 1), max(32 * g, 1) - 1, -1):
            path[i] = j
            j -= bit_from_words(w, i, j)
        if g == 0:
            path[0] = j
            assert j == path0
        else:
            assert j == entry[g - 1], (g, j, entry[g - 1])
    return path


def mas_path(x: np.ndarray) -> np.ndarray:
    n, m = x.shape
    return backtrack(chain_words(forward_bits(x)), n, m)
