"""GPU parity of the MAS kernel: CUDA path (through the C ABI) vs the pinned CPU oracle and
the committed golden vectors.  Everything here is bit-exact: int16 path, int64 durations."""
import numpy as np
import pytest
import torch

from conftest import golden
from isp_tts_b200 import _lib, synth
from isp_tts_b200.mas import b_mas, cuda_b_mas, mas_forward
from oracle import mas as omas

pytestmark = pytest.mark.gpu


def run_cuda(x, tl, ml, dev, **kw):
    xt = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    hard, dur = mas_forward(xt, torch.from_numpy(np.asarray(tl)), torch.from_numpy(np.asarray(ml)), **kw)
    torch.cuda.synchronize()
    return hard.cpu().numpy(), dur.cpu().numpy(), xt


def assert_same(hard, dur, ref_hard, ref_dur, what=""):
    if not np.array_equal(hard, ref_hard):
        bad = np.argwhere(hard != ref_hard)
        b = bad[0][0]
        rows = np.unique(bad[bad[:, 0] == b][:, 1])
        raise AssertionError(f"{what}: path differs in {len(np.unique(bad[:, 0]))} utterance(s); first b={b}, "
                             f"{len(rows)} rows, first rows {rows[:8].tolist()}; cuda cols "
                             f"{hard[b, rows[:8]].argmax(1).tolist()} vs oracle {ref_hard[b, rows[:8]].argmax(1).tolist()}; "
                             f"ones per utt cuda={hard[b].sum()} oracle={ref_hard[b].sum()}")
    assert np.array_equal(dur, ref_dur), f"{what}: durations differ"


@pytest.fixture(autouse=True)
def _reset_options():
    yield
    for key in ("mas.ring_rows", "mas.slots", "mas.bits_global", "mas.no_tma", "mas.impl", "mas.dbg", "mas2.min_pair_stages", "mas2.single", "mas2.together", "mas2.fill_us"):
        _lib.set_option(key, 0)
    _lib.set_option("mas2.pace", -1)


# kernel variants.  isp_mas2.cu (the default up to 256 tokens): tiled TMA or 4 B async copies, one utterance per CTA,
# pairs side by side, pairs one after the other.  isp_mas.cu (wider utterances, or mas.impl = 1): backpointer bits in shared
# memory or in the workspace, one to three utterances per CTA.
MODES = {"auto": {}, "unpaced_pairs": {"mas2.pace": 0, "mas.slots": 2}, "no_tma": {"mas.no_tma": 1}, "one_slot": {"mas.slots": 1}, "two_slots": {"mas.slots": 2},
         "pairs_in_turn": {"mas.slots": 2, "mas2.min_pair_stages": 64}, "two_slots_no_tma": {"mas.slots": 2, "mas.no_tma": 1},
         "v1": {"mas.impl": 1}, "v1_three_slots": {"mas.impl": 1, "mas.slots": 3}, "v1_bits_global": {"mas.impl": 1, "mas.bits_global": 1},
         "v1_two_slots_global_no_tma": {"mas.impl": 1, "mas.slots": 2, "mas.bits_global": 1, "mas.no_tma": 1},
         "wide": {"mas.impl": 3},          # isp_mas_wide.cu, the general kernel behind T2max > 1024, forced onto every shape
         "cluster": {"mas.impl": 4}}       # isp_mas_cluster.cu, one thread-block cluster per utterance (641 .. 1024 tokens), forced onto every shape


def set_mode(mode):
    for key, val in MODES[mode].items():
        _lib.set_option(key, val)


def test_known_answers(cuda_device):
    g = golden("mas_kat.npz")
    hard, dur, _ = run_cuda(g["a12_x"][None], [3], [5], cuda_device)
    assert np.array_equal(hard[0], g["a12_hard"])
    assert dur[0].tolist() == [1, 2, 2]
    for name in ["zeros_10x4", "zeros_3x6", "zeros_1x3", "zeros_5x1", "zeros_1x1", "zeros_7x7"]:
        ref = g[name + "_hard"]
        n, m = ref.shape
        hard, dur, _ = run_cuda(np.zeros((1, n, m), np.float32), [m], [n], cuda_device)
        assert np.array_equal(hard[0], ref), name
        assert np.array_equal(dur[0], ref.sum(0)), name


@pytest.mark.parametrize("mode", list(MODES))
def test_cfg1_single_utterance(cuda_device, mode):
    """BASELINE.json configs[0]: 80 tokens x 400 frames, reference maximum path, bit-exact."""
    set_mode(mode)
    g = golden("mas_cfg1.npz")
    hard, dur, xt = run_cuda(g["x"], g["text_len"], g["mel_len"], cuda_device)
    assert np.array_equal(omas.path_from_hard(hard, g["mel_len"]), g["path"])
    assert np.array_equal(dur, g["durations"])
    assert np.array_equal(xt.cpu().numpy(), g["x"])          # input untouched


@pytest.mark.parametrize("mode", list(MODES))
def test_ragged_batch_with_ties(cuda_device, mode):
    set_mode(mode)
    g = golden("mas_ragged_ties.npz")
    hard, dur, _ = run_cuda(g["x"], g["text_len"], g["mel_len"], cuda_device)
    assert_same(hard, dur, g["hard"], g["hard"].sum(axis=1, dtype=np.int64), "ragged_ties")
    assert np.array_equal(dur.sum(1), g["mel_len"])


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("tag", ["cfg2_noise", "cfg2_ties", "odd_shapes", "wide", "long"])
def test_seeded_golden(cuda_device, tag, mode):
    set_mode(mode)
    g = golden(f"mas_seeded_{tag}.npz")
    B, T1, T2 = int(g["B"]), int(g["T1"]), int(g["T2"])
    x = synth.noise_logits(B, T1, T2, int(g["seed"]), quantize=float(g["quantize"]))
    hard, dur, _ = run_cuda(x, g["text_len"], g["mel_len"], cuda_device)
    assert np.array_equal(omas.path_from_hard(hard, g["mel_len"]), g["path"]), tag
    assert np.array_equal(dur, g["durations"]), tag
    # nothing outside each window
    for b in range(B):
        assert hard[b, int(g["mel_len"][b]):].sum() == 0 and hard[b, :, int(g["text_len"][b]):].sum() == 0


@pytest.mark.parametrize("ring", [80, 96, 128])
def test_small_ring_wraparound(cuda_device, ring):
    """Few rows in flight: every ring stage is reused many times."""
    x = synth.noise_logits(9, 257, 132, 5, quantize=0.25)
    tl, ml = synth.lengths(9, 132, 257, True, 5)
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    for mode in ("auto", "unpaced_pairs", "no_tma", "two_slots", "v1", "pairs_in_turn"):
        set_mode(mode)
        _lib.set_option("mas.ring_rows", ring)
        hard, dur, _ = run_cuda(x, tl, ml, cuda_device)
        assert_same(hard, dur, rh, rd, f"ring={ring} mode={mode}")


def test_full_size_cfg3_against_oracle(cuda_device):
    """BASELINE.json configs[2]: batch 256 ragged, <=200 x <=1000 (the bench workload)."""
    w = synth.WORKLOADS["cfg3"]
    tl, ml = synth.workload_lengths(w)
    x = synth.noise_logits(w.batch, w.t1max, w.t2max, w.seed, quantize=0.125)
    hard, dur, _ = run_cuda(x, tl, ml, cuda_device)
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    assert_same(hard, dur, rh, rd, "cfg3")
    # size-independent properties
    assert np.array_equal(hard.sum(axis=2)[np.arange(w.t1max)[None] < ml[:, None]], np.ones(int(ml.sum()), np.int64))
    assert np.array_equal(dur.sum(1), ml)
    p = omas.path_from_hard(hard, ml)
    for b in range(w.batch):
        pb = p[b, :ml[b]].astype(np.int64)
        d = np.diff(pb)
        assert pb[-1] == tl[b] - 1 and np.all((d == 0) | (d == 1))


def test_full_size_cfg4_long_form(cuda_device):
    """BASELINE.json configs[3]: 512 tokens x 4096 frames, batch 16: the cluster kernel by default (isp_mas_cluster.cu, 4 CTAs per
    utterance), the single-CTA strip kernel with its bits spilled to the workspace when forced (v1, with and without TMA)."""
    w = synth.WORKLOADS["cfg4"]
    tl, ml = synth.workload_lengths(w)
    x = synth.noise_logits(w.batch, w.t1max, w.t2max, w.seed)
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    for mode in ({}, {"mas.impl": 1}, {"mas.impl": 1, "mas.no_tma": 1}):
        for key, val in mode.items():
            _lib.set_option(key, val)
        hard, dur, _ = run_cuda(x, tl, ml, cuda_device)
        assert_same(hard, dur, rh, rd, f"cfg4 mode={mode}")
        for key in mode:
            _lib.set_option(key, 0)


def test_maximum_width(cuda_device):
    T2 = 640                                              # ISP_MAS_MAX_T2
    x = synth.noise_logits(2, 300, T2, 11, quantize=0.5)
    tl, ml = np.array([640, 450]), np.array([300, 211])
    hard, dur, _ = run_cuda(x, tl, ml, cuda_device)       # text longer than mel: pure diagonal tail
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    assert_same(hard, dur, rh, rd, "T2=640")
    # one token more: the cluster kernel (isp_mas_cluster.cu) takes over, same answers
    x = synth.noise_logits(2, 300, 641, 12, quantize=0.5)
    tl, ml = np.array([641, 450]), np.array([300, 211])
    hard, dur, _ = run_cuda(x, tl, ml, cuda_device)
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    assert_same(hard, dur, rh, rd, "T2=641")
    with pytest.raises(_lib.IspError):                     # ISP_MAS_WIDE_MAX_T2
        mas_forward(torch.zeros(1, 2, 16385, device=cuda_device), torch.tensor([16385]), torch.tensor([2]))


@pytest.mark.parametrize("shape", [(3, 900, 1000), (2, 2500, 2100), (2, 700, 4097), (1, 5000, 1024)])
def test_wide_utterances_general_kernel(cuda_device, shape):
    """T2max > 640 (long-form text): isp_mas_cluster.cu up to 1024 tokens, isp_mas_wide.cu beyond, against the oracle, ragged, with ties, path and durations bit-exact,
    also through the path-only entry point."""
    B, T1, T2 = shape
    x = synth.noise_logits(B, T1, T2, 77 + T2, quantize=0.25)
    tl, ml = synth.lengths(B, T2, T1, True, 78 + T2)
    tl[0], ml[0] = T2, T1
    if B > 1:
        tl[1], ml[1] = min(T2, 700), min(T1, 650)           # more tokens than frames: the pure diagonal
    hard, dur, xt = run_cuda(x, tl, ml, cuda_device)
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    assert_same(hard, dur, rh, rd, f"wide {shape}")
    assert np.array_equal(dur.sum(1), ml)
    none, dur2, path = mas_forward(xt, torch.from_numpy(tl), torch.from_numpy(ml), return_path=True, dense=False)
    assert none is None and np.array_equal(dur2.cpu().numpy(), rd)
    ref_path = omas.path_from_hard(rh, ml)
    got = path.cpu().numpy()
    for b in range(B):
        assert np.array_equal(got[b, :ml[b]], ref_path[b, :ml[b]]) and np.all(got[b, ml[b]:] == -1)


@pytest.mark.parametrize("mode", ["auto", "cluster"])
def test_unaligned_and_strided_inputs(cuda_device, mode):
    """Rows that are not 16 B aligned take the non-TMA producer path; padded row strides too (also on the cluster kernel, whose
    tensor map then carries the padded stride)."""
    set_mode(mode)
    rs = np.random.RandomState(3)
    for (B, T1, T2) in [(3, 50, 33), (2, 64, 7), (4, 31, 1), (2, 1, 9)]:
        x = (np.rint(rs.standard_normal((B, T1, T2)) * 4) / 4).astype(np.float32)
        tl = rs.randint(1, T2 + 1, size=B); ml = rs.randint(1, T1 + 1, size=B)
        tl[0], ml[0] = T2, T1
        hard, dur, _ = run_cuda(x, tl, ml, cuda_device)
        rh, rd = omas.b_mas_with_durations(x, tl, ml)
        assert_same(hard, dur, rh, rd, f"shape {(B, T1, T2)}")
    # a strided view: row stride 40 floats for 36 columns, base offset 4 B off alignment
    buf = torch.zeros(3 * 20 * 40 + 1, device=cuda_device)
    x = (np.rint(rs.standard_normal((3, 20, 36)) * 2) / 2).astype(np.float32)
    view = buf[1:].view(3, 20, 40)[:, :, :36]
    view.copy_(torch.from_numpy(x))
    tl, ml = np.array([36, 10, 22]), np.array([20, 20, 5])
    hard, dur = mas_forward(view, torch.from_numpy(tl), torch.from_numpy(ml))
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    assert_same(hard.cpu().numpy(), dur.cpu().numpy(), rh, rd, "strided view")
    # 16 B aligned rows with a padded stride and a padded batch stride: the TMA path with strides that are not the extents
    buf = torch.zeros(3 * 70 * 48 + 64, device=cuda_device)
    x = (np.rint(rs.standard_normal((3, 66, 40)) * 2) / 2).astype(np.float32)
    view = buf[: 3 * 70 * 48].view(3, 70, 48)[:, :66, :40]
    view.copy_(torch.from_numpy(x))
    tl, ml = np.array([40, 17, 33]), np.array([66, 66, 9])
    hard, dur = mas_forward(view, torch.from_numpy(tl), torch.from_numpy(ml))
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    assert_same(hard.cpu().numpy(), dur.cpu().numpy(), rh, rd, "aligned strided view")


def test_reference_entry_points(cuda_device):
    """b_mas (numpy in/out) and cuda_b_mas[grid, block](...) keep the reference call shapes."""
    g = golden("mas_ragged_ties.npz")
    out = b_mas(g["x"].copy(), g["text_len"], g["mel_len"])                  # mas.py:30
    assert isinstance(out, np.ndarray) and out.dtype == np.int16 and np.array_equal(out, g["hard"])
    # alignment.py:321-331, verbatim apart from the import
    attn_logits = torch.from_numpy(g["x"]).to(cuda_device)
    text_len = torch.from_numpy(g["text_len"]).to(cuda_device)
    mel_len = torch.from_numpy(g["mel_len"]).to(cuda_device)
    log_p = attn_logits.clone().detach()
    attn_out = torch.zeros_like(attn_logits, dtype=torch.int16)
    prev_log_p = torch.zeros_like(attn_logits)
    prev_ind = torch.zeros_like(attn_logits, dtype=torch.int16)
    cuda_b_mas[(max(64, attn_logits.shape[0]), 1), (1, 256)](log_p, prev_log_p, prev_ind, attn_out, text_len, mel_len)
    assert np.array_equal(attn_out.cpu().numpy(), g["hard"])
    assert torch.equal(log_p, attn_logits)


@pytest.mark.parametrize("mode", ["auto", "v1", "wide", "cluster"])
def test_out_of_contract_lengths_are_reported(cuda_device, mode):
    set_mode(mode)
    x = torch.zeros(3, 8, 5, device=cuda_device)
    with pytest.raises(_lib.IspError):
        mas_forward(x, torch.tensor([5, 9, 2]), torch.tensor([8, 8, 0]), check_lengths=True)
    hard, dur = mas_forward(x, torch.tensor([5, 4, 2]), torch.tensor([8, 8, 3]), check_lengths=True)
    assert dur.sum(1).tolist() == [8, 8, 3]


def test_cpu_tensor_is_rejected():
    with pytest.raises(_lib.IspError):
        mas_forward(torch.zeros(1, 4, 4), torch.tensor([4]), torch.tensor([4]))


@pytest.mark.parametrize("T2,mode", [(72, "auto"), (72, "v1_bits_global"), (160, "auto"), (160, "v1"), (160, "no_tma"), (160, "pairs_in_turn")])
def test_slots_take_several_utterances(cuda_device, T2, mode):
    """More utterances than 2 x #SMs (the cfg5 sweep's regime): every persistent slot aligns several utterances
    in a row, with one strip (72 tokens) or two (160), bits in shared memory or in the workspace."""
    set_mode(mode)
    B, T1 = 700, 150
    x = synth.noise_logits(B, T1, T2, 21, quantize=0.25)
    tl, ml = synth.lengths(B, T2, T1, True, 22)
    hard, dur, _ = run_cuda(x, tl, ml, cuda_device)
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    assert_same(hard, dur, rh, rd, f"B=700 T2={T2} {mode}")
    assert np.array_equal(dur.sum(1), ml)


def test_path_output_and_binarization_loss(cuda_device):
    """isp_mas_forward_path + isp_bin_loss_sums against the reference formula on the dense tensors
    (tts/models/acoustic/loss.py:97-105), value and gradient."""
    from isp_tts_b200.mas import binarization_loss
    B, T1, T2 = 5, 140, 36
    x = synth.noise_logits(B, T1, T2, 31)
    tl, ml = synth.lengths(B, T2, T1, True, 32)
    xt = torch.from_numpy(x).to(cuda_device)
    mlt = torch.from_numpy(ml).to(cuda_device)
    hard, dur, path = mas_forward(xt, torch.from_numpy(tl), mlt, return_path=True)
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    assert_same(hard.cpu().numpy(), dur.cpu().numpy(), rh, rd, "path run")
    ref_path = omas.path_from_hard(rh, ml).astype(np.int64)
    got = path.cpu().numpy().astype(np.int64)
    for b in range(B):
        assert np.array_equal(got[b, :ml[b]], ref_path[b, :ml[b]]) and np.all(got[b, ml[b]:] == -1)
    # a soft attention with some entries below eps on the path
    soft = torch.softmax(xt * 3.0, dim=2)
    soft = (soft * (soft > 1e-3)).detach().requires_grad_(True)
    loss = binarization_loss(soft, path, mlt)
    loss.backward()
    soft2 = soft.detach().clone().requires_grad_(True)
    ref = -torch.log(torch.clamp(soft2[hard == 1], min=1e-6)).sum() / hard.sum()
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    assert torch.allclose(soft.grad, soft2.grad, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_length_regulator_from_path(cuda_device, dtype):
    """isp_length_regulate (+ backward) against the reference's matmul with the 0/1 matrix built from the cumulated
    durations (tts/models/acoustic/modules/temporal_adaptor.py:420-431)."""
    from isp_tts_b200.consumers import LengthRegulator
    B, T1, T2, C = 4, 150, 40, 48
    x_l = synth.noise_logits(B, T1, T2, 41)
    tl, ml = synth.lengths(B, T2, T1, True, 42)
    mlt = torch.from_numpy(ml).to(cuda_device)
    hard, dur, path = mas_forward(torch.from_numpy(x_l).to(cuda_device), torch.from_numpy(tl), mlt, return_path=True)
    gen = torch.Generator(device=cuda_device).manual_seed(3)
    x = torch.randn((B, T2, C), device=cuda_device, generator=gen).to(dtype).requires_grad_(True)
    out, dec = LengthRegulator()(x, dur, path=path)
    assert torch.equal(dec, mlt)
    # reference: temporal_adaptor.py:420-431
    x2 = x.detach().clone().float().requires_grad_(True)
    reps = (dur.float() + 0.5).long()
    cums = torch.cumsum(torch.nn.functional.pad(reps, (1, 0, 0, 0), value=0.0), dim=1, dtype=torch.float32)[:, None, :]
    r = torch.arange(int(reps.sum(1).max()), device=cuda_device)[None, :, None]
    mult = ((cums[:, :, :-1] <= r) & (cums[:, :, 1:] > r)).float()
    ref = torch.matmul(mult, x2)
    assert out.shape[1] == T1 and ref.shape[1] == int(ml.max())
    assert torch.equal(out[:, :ref.shape[1]].float(), ref.to(dtype).float())           # a gather: exact
    assert out.detach()[:, ref.shape[1]:].abs().sum().item() == 0.0
    w = torch.randn(out.shape, device=cuda_device, generator=gen)
    (out.float() * w).sum().backward()
    (ref * w[:, :ref.shape[1]]).sum().backward()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert torch.allclose(x.grad.float(), x2.grad, rtol=tol, atol=tol * 10)


def test_temporal_averager_from_durations(cuda_device):
    """isp_temporal_average against the reference's running-sum formula (temporal_adaptor.py:447-465), evaluated in
    float64 so that it is exact, and against the same formula in fp32 within its own cancellation error."""
    from isp_tts_b200.consumers import TemporalAverager
    B, T1, T2, C = 6, 400, 90, 2
    x_l = synth.noise_logits(B, T1, T2, 51)
    tl, ml = synth.lengths(B, T2, T1, True, 52)
    mlt = torch.from_numpy(ml).to(cuda_device)
    _, dur = mas_forward(torch.from_numpy(x_l).to(cuda_device), torch.from_numpy(tl), mlt)
    gen = torch.Generator(device=cuda_device).manual_seed(5)
    x = torch.rand((B, C, T1), device=cuda_device, generator=gen) * 200.0 + 80.0       # pitch-like, Hz
    x[torch.rand((B, C, T1), device=cuda_device, generator=gen) < 0.3] = 0.0            # unvoiced frames
    x = x.masked_fill(torch.arange(T1, device=cuda_device)[None, None, :] >= mlt[:, None, None], 0.0)
    out = TemporalAverager()(x, dur)

    def reference(xx):
        F = torch.nn.functional
        ends = torch.cumsum(dur, dim=1).long()
        starts = F.pad(ends[:, :-1], (1, 0))
        nz = F.pad(torch.cumsum(xx != 0.0, dim=2), (1, 0))
        cs = F.pad(torch.cumsum(xx, dim=2), (1, 0))
        dcs = starts[:, None, :].expand(B, C, T2)
        dce = ends[:, None, :].expand(B, C, T2)
        sums = torch.gather(cs, 2, dce) - torch.gather(cs, 2, dcs)
        n = (torch.gather(nz, 2, dce) - torch.gather(nz, 2, dcs)).to(xx.dtype)
        return torch.where(n == 0.0, n, sums / n)

    exact = reference(x.double())
    assert torch.allclose(out.double(), exact, rtol=1e-6, atol=1e-4)
    assert torch.allclose(out, reference(x), rtol=1e-3, atol=0.5)       # the fp32 running sums reach ~1e5 here
    assert (out.masked_select(dur[:, None, :].expand(B, C, T2) == 0) == 0).all()


@pytest.mark.parametrize("T2", [72, 200])
def test_path_only_without_dense_output(cuda_device, T2):
    """isp_mas_forward_path with attn_hard == NULL: same path and durations as the run that also writes the dense tensor."""
    B, T1 = 9, 260
    x = synth.noise_logits(B, T1, T2, 61, quantize=0.5)
    tl, ml = synth.lengths(B, T2, T1, True, 62)
    xt = torch.from_numpy(x).to(cuda_device)
    hard, dur, path = mas_forward(xt, torch.from_numpy(tl), torch.from_numpy(ml), return_path=True)
    none, dur2, path2 = mas_forward(xt, torch.from_numpy(tl), torch.from_numpy(ml), return_path=True, dense=False)
    assert none is None and torch.equal(dur, dur2) and torch.equal(path, path2)
    # the path is the dense tensor's argmax on valid frames
    for b in range(B):
        assert torch.equal(hard[b, :ml[b]].argmax(dim=1).to(torch.int16), path[b, :ml[b]])
    with pytest.raises(ValueError):
        mas_forward(xt, torch.from_numpy(tl), torch.from_numpy(ml), dense=False)


def _realistic_logits(dev, name, batch=None):
    """attn_logits as the hot path produces them (values in about [-20, -5], 41 % of the valid cells on the
    log(1e-6) plateau of the prior): the log-likelihood kernel's own output on the workload's synthetic encodings."""
    from isp_tts_b200.alignment import loglik_forward
    w = synth.WORKLOADS[name]
    tl, ml = synth.workload_lengths(w, batch)
    B = len(tl)
    q, k = synth.encoded_pair(B, w.t1max, w.t2max, w.dim, tl, ml, w.seed + 1)
    soft, logits = loglik_forward(torch.from_numpy(q).to(dev).to(torch.bfloat16), torch.from_numpy(k).to(dev).to(torch.bfloat16),
                                  torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev))
    return logits, tl, ml


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_realistic_logits_against_oracle(cuda_device, name):
    """MAS on the SAME fp32 tensor for both sides: the kernel's own attn_logits into mas_forward and into the oracle."""
    logits, tl, ml = _realistic_logits(cuda_device, name)
    hard, dur = mas_forward(logits, torch.from_numpy(tl), torch.from_numpy(ml))
    torch.cuda.synchronize()
    x = logits.cpu().numpy()
    valid = (np.arange(x.shape[1])[None, :, None] < ml[:, None, None]) & (np.arange(x.shape[2])[None, None, :] < tl[:, None, None])
    plateau = np.isclose(x, np.log(np.float32(1e-6)), atol=3.0) & valid            # prior floor + (S - lse) within a few units
    assert x[valid].min() > -60 and x[valid].max() < 0 and plateau.sum() > 0.2 * valid.sum()
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    assert_same(hard.cpu().numpy(), dur.cpu().numpy(), rh, rd, f"{name} realistic logits")
    assert np.array_equal(dur.cpu().numpy().sum(1), ml)


@pytest.mark.parametrize("shape", [(4, 1500, 384), (2, 3000, 513), (2, 4096, 1024), (3, 700, 260), (5, 37, 1000), (2, 8000, 300)])
def test_cluster_kernel_shapes(cuda_device, shape):
    """isp_mas_cluster.cu (one thread-block cluster per utterance, 1 .. 8 CTAs of 128 tokens): what isp_mas_forward picks for
    641 .. 1024 tokens, forced here onto narrower shapes too.  Ragged with ties; full-length, tiny and token-heavy utterances in the same batch (clusters whose right
    CTAs idle, paths that cannot reach column 0); odd T2max takes the loader's register path; path-only entry point too."""
    B, T1, T2 = shape
    if T2 <= 640:
        _lib.set_option("mas.impl", 4)
    x = synth.noise_logits(B, T1, T2, 177 + T2, quantize=0.25)
    tl, ml = synth.lengths(B, T2, T1, True, 178 + T2)
    tl[0], ml[0] = T2, T1
    tl[1], ml[1] = min(T2, T1 + 40), max(1, min(T1, T2 - 60))     # more tokens than frames: the pure diagonal
    if B > 2:
        tl[2], ml[2] = 3, min(T1, 17)                             # an utterance that lives in the first CTA only
    hard, dur, xt = run_cuda(x, tl, ml, cuda_device)
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    assert_same(hard, dur, rh, rd, f"cluster {shape}")
    _, dur2, path = mas_forward(xt, torch.from_numpy(tl), torch.from_numpy(ml), return_path=True, dense=False)
    assert np.array_equal(dur2.cpu().numpy(), rd)
    ref_path = omas.path_from_hard(rh, ml)
    got = path.cpu().numpy()
    for b in range(B):
        assert np.array_equal(got[b, :ml[b]], ref_path[b, :ml[b]]) and np.all(got[b, ml[b]:] == -1)


@pytest.mark.parametrize("mode", ["auto", "two_slots", "v1", "cluster"])
def test_all_plateau_adversarial(cuda_device, mode):
    """Every valid cell on the prior's floor (one constant): every comparison of the DP is a tie, so the tie rule alone
    decides the path (surplus frames go to the FIRST token); then the same with one better column."""
    set_mode(mode)
    B, T1, T2 = 7, 333, 150
    tl, ml = synth.lengths(B, T2, T1, True, 71)
    tl[1], ml[1] = 150, 149                                      # more tokens than frames: the pure diagonal
    x = np.full((B, T1, T2), np.log(np.float32(1e-6)), dtype=np.float32)
    hard, dur, _ = run_cuda(x, tl, ml, cuda_device)
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    assert_same(hard, dur, rh, rd, "all plateau")
    for b in range(B):
        if ml[b] >= tl[b]:
            assert dur[b, 0] == ml[b] - tl[b] + 1 and np.all(dur[b, 1:tl[b]] == 1)
    x[:, :, 40] += np.float32(0.5)                               # a ridge in one column
    hard, dur, _ = run_cuda(x, tl, ml, cuda_device)
    rh, rd = omas.b_mas_with_durations(x, tl, ml)
    assert_same(hard, dur, rh, rd, "plateau with a ridge")
