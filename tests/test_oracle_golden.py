"""Pins the CPU oracle (oracle/) to outputs of the UNMODIFIED reference.

The fixtures under tests/golden/ were produced by oracle/gen_golden.py, which
ran the reference's own `b_mas` / `mas_width1` (tts/modules/aligner/mas.py)
and `ConvAttention.forward` (tts/models/acoustic/modules/alignment.py:158-208)
in the build container.  MAS: bit-exact.  Log-likelihood: fp32 tolerance below.
"""
import numpy as np
import pytest

from conftest import golden
from isp_tts_b200 import synth
from oracle import loglik as oll
from oracle import mas as omas


def _hard_from_path(path, T2):
    B, T1 = path.shape
    out = np.zeros((B, T1, T2), np.int16)
    b, i = np.nonzero(path >= 0)
    out[b, i, path[b, i]] = 1
    return out


def test_known_answers():
    g = golden("mas_kat.npz")
    x = g["a12_x"]
    hard, dur = omas.b_mas_with_durations(x[None], [3], [5])
    assert np.array_equal(hard[0], g["a12_hard"])
    assert hard[0].argmax(1).tolist() == [0, 1, 1, 2, 2]
    assert dur[0].tolist() == [1, 2, 2]
    # the DP arithmetic itself: the reference leaves the accumulated Q in its input
    assert np.array_equal(omas.accumulate(x), g["a12_Q"])
    for name in ["zeros_10x4", "zeros_3x6", "zeros_1x3", "zeros_5x1", "zeros_1x1", "zeros_7x7"]:
        ref = g[name + "_hard"]
        n, m = ref.shape
        hard, dur = omas.b_mas_with_durations(np.zeros((1, n, m), np.float32), [m], [n])
        assert np.array_equal(hard[0], ref), name
        assert np.array_equal(dur[0], ref.sum(0)), name
    assert omas.b_mas_with_durations(np.zeros((1, 10, 4), np.float32), [4], [10])[1][0].tolist() == [7, 1, 1, 1]
    assert omas.b_mas_with_durations(np.zeros((1, 3, 6), np.float32), [6], [3])[1][0].tolist() == [0, 0, 0, 1, 1, 1]


def test_cfg1_bit_exact():
    g = golden("mas_cfg1.npz")
    w = synth.WORKLOADS["cfg1"]
    assert int(g["seed"]) == w.seed
    x = synth.noise_logits(1, w.t1max, w.t2max, w.seed)
    assert np.array_equal(x, g["x"])                              # the seed regenerates the stored input
    hard, dur = omas.b_mas_with_durations(x, g["text_len"], g["mel_len"])
    assert np.array_equal(omas.path_from_hard(hard, g["mel_len"]), g["path"])
    assert np.array_equal(dur, g["durations"])
    assert np.array_equal(omas.accumulate(x[0])[-1], g["Q_last_row"])
    assert not np.shares_memory(x, hard) and np.array_equal(x, g["x"])   # input not mutated


def test_ragged_ties_bit_exact():
    g = golden("mas_ragged_ties.npz")
    hard, dur = omas.b_mas_with_durations(g["x"], g["text_len"], g["mel_len"])
    assert np.array_equal(hard, g["hard"])
    assert np.array_equal(dur, g["hard"].sum(axis=1))
    assert np.array_equal(dur.sum(axis=1), g["mel_len"])


@pytest.mark.parametrize("tag", ["cfg2_noise", "cfg2_ties", "odd_shapes", "wide", "long"])
def test_seeded_cases_bit_exact(tag):
    g = golden(f"mas_seeded_{tag}.npz")
    B, T1, T2 = int(g["B"]), int(g["T1"]), int(g["T2"])
    tl, ml = synth.lengths(B, T2, T1, bool(g["ragged"]), int(g["seed"]))
    assert np.array_equal(tl, g["text_len"]) and np.array_equal(ml, g["mel_len"])
    x = synth.noise_logits(B, T1, T2, int(g["seed"]), quantize=float(g["quantize"]))
    assert float(x.astype(np.float64).sum()) == float(g["x_checksum"])
    hard, dur = omas.b_mas_with_durations(x, tl, ml)
    assert np.array_equal(hard, _hard_from_path(g["path"], T2))
    assert np.array_equal(dur, g["durations"])


def test_single_thread_equals_parallel():
    x = synth.noise_logits(6, 64, 17, 3, quantize=0.5)
    tl, ml = synth.lengths(6, 17, 64, True, 3)
    a = omas.b_mas(x, tl, ml, nthreads=1)
    b = omas.b_mas(x, tl, ml, nthreads=0)
    assert np.array_equal(a, b)


def test_bad_lengths_rejected():
    with pytest.raises(ValueError):
        omas.b_mas(np.zeros((1, 4, 4), np.float32), [5], [4])
    with pytest.raises(ValueError):
        omas.b_mas(np.zeros((1, 4, 4), np.float32), [4], [0])


# ---- log-likelihood -------------------------------------------------------------------
# fp32 tolerance of the restatement vs the reference (SURVEY.md A.5 measured 1.9e-6 / 1.9e-7)
LOGITS_ATOL = 2e-5
SOFT_ATOL = 2e-6


@pytest.mark.parametrize("tag", ["small", "dim80", "dim128"])
def test_loglik_matches_reference(tag):
    g = golden(f"loglik_{tag}.npz")
    soft, logits, parts = oll.loglik(g["Q"], g["K"], g["text_len"], g["mel_len"], return_parts=True)
    # cells whose prior sits within 1e-5 relative of the hard 1e-4 threshold (alignment.py:35) may flip
    assert np.abs(logits - g["attn_logits"]).max() < LOGITS_ATOL
    assert np.abs(soft - g["attn_soft"]).max() < SOFT_ATOL
    # padded positions: soft is exactly 0, logits are NOT masked
    assert np.all(soft[~parts["mask"]] == 0)
    assert np.all(np.isfinite(logits))
    # MAS on the reference's own logits -> the reference's own path and durations
    hard, dur = omas.b_mas_with_durations(g["attn_logits"], g["text_len"], g["mel_len"])
    assert np.array_equal(hard, g["attn_hard"])
    assert np.array_equal(dur, g["durations"])
    assert np.array_equal(oll.durations_from_hard(hard), dur)


def test_prior_matches_reference_shape_and_floor():
    tl, ml = np.array([5, 3]), np.array([9, 4])
    p = oll.batch_diagonal_prior(tl, ml)
    assert p.shape == (2, 9, 5) and p.dtype == np.float32
    assert np.all(p[1, 4:] == 0) and np.all(p[1, :, 3:] == 0)
    assert np.all((p == 0) | (p >= 1e-4))


def test_consumers_oracle_against_reference_modules():
    """oracle/consumers.py against the reference's own LengthRegulator / TemporalAverager outputs (both routes)."""
    from oracle import consumers as oc
    g = golden("consumers.npz")
    T1 = g["attn_soft"].shape[1]
    out, dec = oc.length_regulate_hard(g["x"], g["durations"], max_len=T1)
    assert np.array_equal(dec, g["lr_hard_len"]) and np.array_equal(out.astype(np.float32), g["lr_hard"])
    out, dec = oc.length_regulate_hard(g["x"], g["float_durations"])
    assert np.array_equal(dec, g["lr_float_len"]) and np.array_equal(out.astype(np.float32), g["lr_float"])
    out, dec = oc.length_regulate_soft(g["x"], g["durations"], g["attn_soft"], max_len=T1)
    assert np.array_equal(dec, g["lr_soft_len"]) and np.allclose(out, g["lr_soft"], rtol=1e-5, atol=1e-5)
    assert np.allclose(oc.temporal_average_hard(g["feat"], g["durations"]), g["avg_hard"], rtol=1e-4, atol=2e-2)   # fp32 running sums
    assert np.allclose(oc.temporal_average_soft(g["feat"], g["attn_soft"]), g["avg_soft"], rtol=1e-4, atol=1e-3)
    # the golden path is the argmax of the reference's hard alignment: durations are its histogram
    for b in range(len(g["mel_len"])):
        p = g["path"][b, :g["mel_len"][b]]
        assert np.array_equal(np.bincount(p, minlength=g["durations"].shape[1]), g["durations"][b])
