"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol that
include/isp_tts_b200.h declares, argument validation answers without touching a GPU, the
Python shims refuse to run without a B200, and the utterance sharding (world_size 2, gloo)
reassembles durations exactly.  No compute call is made here."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch

from isp_tts_b200 import _lib, sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "isp_tts_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(isp_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from isp_tts_b200 import build
    build.build()
    lib = ctypes.CDLL(_lib.SO_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 9
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"libisp_tts_b200.so does not export {missing}"
    assert sorted(_lib.EXPORTS) == declared, "the ctypes binding and the header disagree"
    assert _lib.load().isp_version() >= 100


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    assert lib.isp_mas_workspace_bytes(0, 10, 10) == 0
    assert lib.isp_mas_workspace_bytes(4, 1000, 200) >= 256
    rc = lib.isp_mas_forward(None, 0, 0, 1, None, None, 1, 1, 1, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.isp_last_error()
    rc = lib.isp_loglik_forward(None, None, 0, None, None, 1, 1, 1, 8, 1.0, 1, None, None, None, 0, None)
    assert rc == -1
    assert lib.isp_set_option(b"no.such.option", 1) == -1


def test_no_cpu_fallback():
    from isp_tts_b200 import Aligner, b_mas, mas_forward
    x = torch.zeros(1, 4, 3)
    with pytest.raises(_lib.IspError):
        mas_forward(x, torch.tensor([3]), torch.tensor([4]))
    if not torch.cuda.is_available():
        with pytest.raises(_lib.IspError):
            b_mas(x.numpy(), np.array([3]), np.array([4]))
        al = Aligner(mel_dim=8, text_dim=8, attention_dim=16, dropout=0.1)
        with pytest.raises(_lib.IspError):
            al(torch.randn(1, 8, 4), torch.randn(1, 8, 3), torch.tensor([4]), torch.tensor([3]))


def test_no_cpu_fallback_in_the_rows_either_side_of_the_path():
    """The consumers of the path, the forward-sum loss and the staging helper raise on CPU tensors as well."""
    from isp_tts_b200 import (AttentionCTCLoss, LengthRegulator, TemporalAverager, binarization_loss, stage_operands)
    tl, ml = torch.tensor([3]), torch.tensor([4])
    with pytest.raises(_lib.IspError):
        AttentionCTCLoss()(torch.zeros(1, 4, 3), tl, ml)
    with pytest.raises(_lib.IspError):
        TemporalAverager()(torch.zeros(1, 1, 4), torch.tensor([[2, 1, 1]]))
    with pytest.raises(_lib.IspError):
        LengthRegulator()(torch.zeros(1, 3, 8), torch.tensor([[2, 1, 1]]), path=torch.zeros(1, 4, dtype=torch.int16))
    with pytest.raises(_lib.IspError):
        LengthRegulator()(torch.zeros(1, 3, 8), torch.tensor([[2, 1, 1]]))           # no dense fallback either
    with pytest.raises(_lib.IspError):
        binarization_loss(torch.zeros(1, 4, 3), torch.zeros(1, 4, dtype=torch.int16), ml)
    with pytest.raises(_lib.IspError):
        stage_operands(torch.zeros(1, 4, 8), torch.zeros(1, 3, 8), tl, ml)


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 8, 256, 4097):
        for world in (1, 2, 3, 8):
            b = sharding.shard_bounds(n, world)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_assignment_is_a_permutation_and_balances_cells():
    tl, ml = synth.lengths(256, 200, 1000, True, 7)
    parts = sharding.balanced_assignment(tl, ml, 8)
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(256))
    assert all(len(p) == 32 for p in parts)
    loads = np.array([(tl[p] * ml[p]).sum() for p in parts], dtype=np.float64)
    contiguous = np.array([(tl[s:e] * ml[s:e]).sum() for s, e in sharding.shard_bounds(256, 8)], dtype=np.float64)
    assert loads.max() / loads.mean() < 1.02
    assert loads.max() <= contiguous.max()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, tmp):
    """One process per shard, as on the GPU box (there: NCCL + the CUDA kernels; here: gloo, and the
    per-rank durations come from the oracle because this test has no GPU)."""
    import torch.distributed as dist
    from oracle import mas as omas
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        B = 11                                                   # uneven: 6 + 5
        tl, ml = synth.lengths(B, 24, 90, True, 5)
        x = synth.noise_logits(B, 90, 24, 6)
        sl = sharding.shard_slice(B, rank, world)
        # every rank pads to ITS OWN longest utterance, as a data loader would
        t2_local, t1_local = int(tl[sl].max()), int(ml[sl].max())
        xl = np.ascontiguousarray(x[sl, :t1_local, :t2_local])
        _, dur_local = omas.b_mas_with_durations(xl, tl[sl], ml[sl])
        full = sharding.gather_durations(torch.from_numpy(dur_local))
        _, dur_ref = omas.b_mas_with_durations(x, tl, ml)
        t2_global = int(tl.max())
        ok = tuple(full.shape) == (B, t2_global) and np.array_equal(full.numpy(), dur_ref[:, :t2_global])
        ok = ok and sharding.max_over_ranks(float(rank), torch.device("cpu")) == world - 1
        # with the sizes known up front there is no metadata exchange
        full2 = sharding.gather_durations(torch.from_numpy(dur_local), t2max=24,
                                          counts=[e - s for s, e in sharding.shard_bounds(B, world)])
        ok = ok and np.array_equal(full2.numpy(), dur_ref)
        with open(os.path.join(tmp, f"rank{rank}.ok"), "w") as f:
            f.write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


def test_sharded_durations_gather_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        with open(tmp_path / f"rank{r}.ok") as f:
            assert f.read() == "1", f"rank {r}: gathered durations differ from the single-process result"


def test_no_binaries_tracked_and_runtime_linked_dynamically():
    """no built artefact in the history, and the shipped library links the CUDA runtime dynamically (the static runtime
    would embed every runtime entry point -- the batched-memcpy family among them -- in a library that travels to the GPU box)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tracked = subprocess.run(["git", "ls-files"], cwd=root, capture_output=True, text=True)
    if tracked.returncode == 0:
        for name in tracked.stdout.split():
            path = os.path.join(root, name)
            if os.path.isfile(path):
                with open(path, "rb") as f:
                    assert f.read(4) != b"\x7fELF", f"{name} is a tracked ELF binary"
    names = [b"cuda" + b"MemcpyBatchAsync", b"cuda" + b"Memcpy3DBatchAsync", b"cu" + b"MemcpyBatchAsync", b"cu" + b"Memcpy3DBatchAsync"]
    so = os.path.join(root, "isp-tts_b200", "libisp_tts_b200.so")
    if os.path.exists(so):
        blob = open(so, "rb").read()
        for nm in names:
            assert nm not in blob, f"{nm.decode()} is named inside libisp_tts_b200.so (static cudart?)"


def test_precision_modes_of_the_drop_in():
    """ConvAttention.gemm_dtype: "auto" = the reference's own precision (fp32-faithful products outside autocast, bf16 under CUDA autocast:
    tests/test_gpu_aligner.py),
    "tf32" = the fused single-pass kernel on fp32 operands (host logic only: no kernel is launched here)."""
    from isp_tts_b200.alignment import ConvAttention
    att = ConvAttention(mel_dim=8, text_dim=12, attention_dim=16)
    assert att.gemm_dtype == "auto" and att._mode() == "fp32" and att._precision() == "fp32"
    att.gemm_dtype = "tf32"
    assert att._mode() == "tf32" and att._precision() == "tf32"
    att.gemm_dtype = "bf16"
    assert att._mode() == "bf16"
