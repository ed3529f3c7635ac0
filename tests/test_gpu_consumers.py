"""The duration / alignment consumers against outputs of the reference's own modules
(tests/golden/consumers.npz, written by oracle/gen_golden.py::gen_consumers from tts/models/acoustic/modules/
temporal_adaptor.py) and against the float64 oracle (oracle/consumers.py), on both routes: hard durations and the recipe's
soft alignment."""
import numpy as np
import pytest
import torch

from conftest import golden
from isp_tts_b200.consumers import LengthRegulator, TemporalAverager, path_from_durations
from isp_tts_b200.soft import soft_average, soft_expand
from oracle import consumers as oc

pytestmark = pytest.mark.gpu


def dev_tensors(g, dev, *names):
    return [torch.from_numpy(g[n]).to(dev) for n in names]


def test_length_regulator_calls_of_the_reference(cuda_device):
    """The three ways temporal_adaptor.py calls the module (:300 with and without `alignment`, :325), argument for argument."""
    g = golden("consumers.npz")
    x, dur, soft, path, fdur = dev_tensors(g, cuda_device, "x", "durations", "attn_soft", "path", "float_durations")
    T1 = soft.shape[1]
    lr = LengthRegulator()
    # hard durations, no path given: rebuilt on the device
    out, dec = lr(x, dur, max_len=T1)
    assert np.array_equal(dec.cpu().numpy(), g["lr_hard_len"]) and np.array_equal(out.cpu().numpy(), g["lr_hard"])
    # hard durations with the MAS path (what Aligner-aware callers pass): the same gather
    out2, dec2 = lr(x, dur, max_len=T1, path=path)
    assert torch.equal(out2[:, :out.shape[1]], out) and torch.equal(dec2, dec)
    # predicted (float) durations, rounded like the reference (:423)
    out, dec = lr(x, fdur)
    assert np.array_equal(dec.cpu().numpy(), g["lr_float_len"]) and np.array_equal(out.cpu().numpy(), g["lr_float"])
    # the recipe's soft route: keyword call of temporal_adaptor.py:300
    out, dec = lr(x, dur, max_len=T1, alignment=soft)
    assert np.array_equal(dec.cpu().numpy(), g["lr_soft_len"])
    ref = g["lr_soft"]
    assert out.shape == ref.shape
    assert np.abs(out.cpu().numpy() - ref).max() <= 3e-3 * np.abs(ref).max()          # TF32 products, fp32 accumulate
    ref64, _ = oc.length_regulate_soft(g["x"], g["durations"], g["attn_soft"], max_len=T1)
    assert np.abs(out.cpu().numpy() - ref64).max() <= 3e-3 * np.abs(ref64).max()


def test_path_from_durations(cuda_device):
    g = golden("consumers.npz")
    dur, path = dev_tensors(g, cuda_device, "durations", "path")
    got = path_from_durations(dur, path.shape[1])
    assert torch.equal(got, path)
    z = torch.zeros((2, 5), dtype=torch.int64, device=cuda_device)                  # no frames at all: everything -1
    assert (path_from_durations(z, 7) == -1).all()


def test_temporal_averager_both_routes(cuda_device):
    g = golden("consumers.npz")
    feat, dur, soft = dev_tensors(g, cuda_device, "feat", "durations", "attn_soft")
    av = TemporalAverager()
    hard = av(feat, dur).cpu().numpy()
    assert np.allclose(hard, oc.temporal_average_hard(g["feat"], g["durations"]), rtol=1e-6, atol=1e-4)
    assert np.allclose(hard, g["avg_hard"], rtol=1e-4, atol=2e-2)                   # the reference's fp32 running sums
    got = av(feat, dur, soft).cpu().numpy()                                          # temporal_adaptor.py:340 call shape
    assert got.shape == g["avg_soft"].shape
    assert np.allclose(got, g["avg_soft"], rtol=1e-5, atol=1e-3)
    assert np.allclose(got, oc.temporal_average_soft(g["feat"], g["attn_soft"]), rtol=2e-6, atol=1e-4)
    # rows past mel_len are zero in attn_soft: telling the kernel so must not change anything
    ml = torch.from_numpy(g["mel_len"]).to(cuda_device)
    assert torch.equal(soft_average(feat, soft, row_len=ml), soft_average(feat, soft))


@pytest.mark.parametrize("shape", [(3, 170, 44, 48), (2, 300, 200, 384), (2, 130, 37, 20)])
def test_soft_route_gradients(cuda_device, shape):
    """soft_expand / soft_average forward and backward against autograd on the reference formulas in float64."""
    B, T1, T2, C = shape
    gen = torch.Generator(device="cpu").manual_seed(7)
    tl = torch.tensor([T2, max(1, T2 // 2), 1][:B])
    ml = torch.tensor([T1, max(1, T1 // 3), 2][:B])
    valid = (torch.arange(T1)[None, :, None] < ml[:, None, None]) & (torch.arange(T2)[None, None, :] < tl[:, None, None])
    soft = torch.softmax(torch.randn((B, T1, T2), generator=gen) * 2 - 1e9 * (~valid), dim=2) * valid
    x = torch.randn((B, T2, C), generator=gen)
    feat = torch.rand((B, 2, T1), generator=gen) * 100
    w1 = torch.randn((B, T1, C), generator=gen)
    w2 = torch.randn((B, 2, T2), generator=gen)

    a = soft.to(cuda_device).requires_grad_(True)
    xx = x.to(cuda_device).requires_grad_(True)
    out = soft_expand(a, xx, frame_len=ml, token_len=tl)
    avg = soft_average(feat.to(cuda_device), a, row_len=ml)
    ((out * w1.to(cuda_device)).sum() + (avg * w2.to(cuda_device)).sum()).backward()

    a64 = soft.double().requires_grad_(True)
    x64 = x.double().requires_grad_(True)
    out64 = (x64.transpose(1, 2) @ a64.transpose(1, 2)).transpose(1, 2)                   # temporal_adaptor.py:419
    avg64 = feat.double() @ a64 / (a64.sum(dim=1, keepdim=True) + 1e-5)                   # :447-448
    ((out64 * w1.double()).sum() + (avg64 * w2.double()).sum()).backward()

    def close(got, ref, tol, what):
        err = (got.detach().cpu().double() - ref).abs().max().item() / (ref.abs().max().item() + 1e-30)
        assert err <= tol, f"{what}: {err:.3e}"

    close(out, out64.detach(), 3e-3, "soft_expand")
    close(avg, avg64.detach(), 1e-5, "soft_average")
    close(xx.grad, x64.grad, 3e-3, "d x")
    ga = a.grad.cpu().double()
    ref_ga = a64.grad * valid                          # entries outside the window are constants of the Aligner (masked)
    close(ga * valid, ref_ga, 3e-3, "d alignment")


def test_soft_route_has_no_cpu_fallback():
    with pytest.raises(Exception):
        soft_expand(torch.zeros((1, 8, 4)), torch.zeros((1, 4, 8)))
    with pytest.raises(Exception):
        soft_average(torch.zeros((1, 1, 8)), torch.zeros((1, 8, 4)))
