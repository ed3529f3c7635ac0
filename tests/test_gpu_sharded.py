"""Sharding a batch by utterance across two GPUs gives, bit for bit, what one GPU computes for the whole batch
(no arithmetic crosses utterances): attn_logits, attn_soft, attn_hard and the durations, the last also
through the path's one collective (NCCL all-gather, isp_tts_b200.sharding.gather_durations).  Needs two GPUs; skipped on
a one-GPU box (the host-side logic is covered with gloo in tests/test_host_boundary.py)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from isp_tts_b200 import sharding, synth
    from isp_tts_b200.alignment import loglik_forward
    from isp_tts_b200.mas import mas_forward
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B, T1, T2, D = 37, 420, 96, 64
    tl, ml = synth.lengths(B, T2, T1, True, 4242)
    q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, 4243)
    parts = sharding.balanced_assignment(tl, ml, world)
    mine = parts[rank]

    def run(idx):
        qd = torch.from_numpy(q[idx]).to(dev).to(torch.bfloat16)
        kd = torch.from_numpy(k[idx]).to(dev).to(torch.bfloat16)
        tld, mld = torch.from_numpy(tl[idx]).to(dev), torch.from_numpy(ml[idx]).to(dev)
        soft, logits = loglik_forward(qd, kd, tld, mld)
        hard, dur = mas_forward(logits, tld, mld)
        return soft, logits, hard, dur

    soft, logits, hard, dur = run(mine)
    full = sharding.gather_durations(dur, counts=[len(p) for p in parts], t2max=T2)          # the path's one collective
    order = np.concatenate(parts)
    if rank == 0:
        ref_soft, ref_logits, ref_hard, ref_dur = run(np.arange(B))
        assert torch.equal(full, ref_dur[torch.from_numpy(order).to(dev)]), "gathered durations differ from the single-GPU run"
    # every rank checks its own shard against a single-GPU run of the whole batch on ITS device
    ref_soft, ref_logits, ref_hard, ref_dur = run(np.arange(B))
    sel = torch.from_numpy(mine).to(dev)
    ok = (torch.equal(soft, ref_soft[sel]) and torch.equal(logits, ref_logits[sel]) and torch.equal(hard, ref_hard[sel])
          and torch.equal(dur, ref_dur[sel]))
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write("ok" if ok else "MISMATCH")
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shards_equal_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"
