"""Utterance sharding across the GPUs of one box.

Nothing on the hot path crosses utterances (the reference processes each `b` alone,
tts/modules/aligner/mas.py:32-34, and the GEMM is batched per `b`), so a batch is split by
utterance, every rank runs the log-likelihood and MAS kernels on its own slice, and there is
NO collective on the data path.  The only exchange offered is an all-gather of the small
int64 duration tensors for callers that want one tensor for the whole batch.

Host-side only: no CUDA code here, so it is importable (and tested, gloo backend) without a GPU.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

__all__ = ["shard_bounds", "shard_slice", "balanced_assignment", "gather_durations", "max_over_ranks"]


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous [start, stop) per rank; the first n % world ranks get one more utterance."""
    if world < 1 or n < 0:
        raise ValueError("world must be >= 1 and n >= 0")
    base, extra = divmod(n, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


def shard_slice(n: int, rank: int, world: int) -> slice:
    s, e = shard_bounds(n, world)[rank]
    return slice(s, e)


def balanced_assignment(text_len: Sequence[int], mel_len: Sequence[int], world: int) -> list[np.ndarray]:
    """Length-balanced alternative to contiguous slices: utterances sorted by cell count
    (T1_b * T2_b, the DP's work) and dealt to the currently lightest rank (LPT).  Returns, per
    rank, the sorted utterance indices it owns; concatenated they are a permutation of range(B)."""
    cells = np.asarray(text_len, dtype=np.int64) * np.asarray(mel_len, dtype=np.int64)
    order = np.argsort(-cells, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    count = np.zeros(world, dtype=np.int64)
    cap = -(-len(cells) // world)            # keep shard sizes within one of each other
    owner = [[] for _ in range(world)]
    for b in order:
        open_ranks = np.flatnonzero(count < cap)
        r = int(open_ranks[np.argmin(load[open_ranks])])
        owner[r].append(int(b))
        load[r] += cells[b]
        count[r] += 1
    return [np.sort(np.asarray(o, dtype=np.int64)) for o in owner]


def max_over_ranks(value: float, device, group=None) -> float:
    """MAX all-reduce of one host number (bench timing, global T2max)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_durations(durations: torch.Tensor, group=None, *, t2max: int | None = None,
                     counts: Sequence[int] | None = None) -> torch.Tensor:
    """All-gather per-rank durations (B_local, T2max_local) int64 -> (sum B_local, T2max) on every rank.

    Ranks may hold different numbers of utterances and different padded widths.  `t2max` (global
    padded width) and `counts` (utterances per rank) can be passed when the caller knows them;
    otherwise they are agreed with one small all-reduce(MAX) / all-gather first.  Uses NCCL for
    CUDA tensors and gloo for CPU tensors -- whatever backend the group was created with."""
    import torch.distributed as dist
    if durations.dim() != 2 or durations.dtype != torch.int64:
        raise ValueError("durations must be an int64 (B_local, T2max) tensor")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if t2max is not None and t2max > durations.shape[1]:
            return torch.nn.functional.pad(durations, (0, t2max - durations.shape[1]))
        return durations
    world = dist.get_world_size(group)
    dev = durations.device
    if t2max is None or counts is None:
        meta = torch.tensor([durations.shape[0], durations.shape[1]], dtype=torch.int64, device=dev)
        metas = torch.empty(world * 2, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(metas, meta, group=group)
        metas = metas.view(world, 2).cpu()
        counts = [int(c) for c in metas[:, 0]] if counts is None else list(counts)
        t2max = int(metas[:, 1].max()) if t2max is None else t2max
    if len(counts) != world:
        raise ValueError("counts must have one entry per rank")
    bmax = max(counts)
    send = torch.zeros((bmax, t2max), dtype=torch.int64, device=dev)
    send[: durations.shape[0], : durations.shape[1]] = durations
    recv = torch.empty((world, bmax, t2max), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(recv.view(world * bmax, t2max), send, group=group)
    return torch.cat([recv[r, : counts[r]] for r in range(world)], dim=0)
