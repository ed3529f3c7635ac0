"""Forward-sum (CTC) alignment loss: the CUDA path (isp_ctc_forward / isp_ctc_backward through the C ABI) against the oracle
(oracle/ctc.py: the reference's op sequence, tts/models/acoustic/loss.py:41-79, with torch on the CPU in float64).

Tolerances (fp32 kernels, log2-domain recursion with ex2.approx / lg2.approx, variables kept relative to each frame's
largest term): nll within 2e-5 relative + 1e-4 absolute; gradient within 1e-3 of the largest gradient entry of the
utterance (measured: <= 2e-4 up to cfg3 shapes, 8e-4 at 512 tokens x 900 frames, where the approximations' bias adds up
over the most steps; torch's own fp32 CTC differs from float64 by as much).
"""
import numpy as np
import pytest
import torch

from isp_tts_b200 import synth
from isp_tts_b200.ctc import AttentionCTCLoss, ctc_nll
from oracle import ctc as octc

pytestmark = pytest.mark.gpu


def realistic_logits(B, T1, T2, tl, ml, seed):
    """Values like the log-likelihood kernel's output: a diagonal ridge around -5 on a floor near -20, constants at padding."""
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((B, T1, T2)).astype(np.float32) * 1.5 - 14.0
    for b in range(B):
        i = np.arange(ml[b])[:, None] / ml[b]
        j = np.arange(tl[b])[None, :] / tl[b]
        x[b, :ml[b], :tl[b]] += 10.0 * np.exp(-((i - j) ** 2) / (2 * 0.1 ** 2)).astype(np.float32)
        x[b, ml[b]:, :] = -19.1
        x[b, :, tl[b]:] = -19.5
    return x


def run_both(x, tl, ml, dev, blank=-1.0):
    xt = torch.from_numpy(x).to(dev).requires_grad_(True)
    tlt, mlt = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
    nll = ctc_nll(xt, tlt, mlt, blank)
    w = torch.linspace(0.5, 1.5, x.shape[0], device=dev)
    finite = torch.where(torch.isinf(nll), torch.zeros_like(nll), nll)
    (finite * w).sum().backward()
    xo = torch.from_numpy(x).double().requires_grad_(True)
    ref = octc.attention_ctc_loss(xo, torch.from_numpy(tl), torch.from_numpy(ml), blank, reduction="none")
    (ref * w.cpu().double()).sum().backward()
    return nll.detach().cpu().double(), xt.grad.cpu().double(), ref.detach(), xo.grad


@pytest.mark.parametrize("shape", [(3, 40, 9), (4, 150, 40), (2, 300, 70), (2, 64, 130), (1, 200, 200),
                                   (2, 90, 37), (3, 77, 3), (2, 700, 300), (1, 1200, 500), (1, 900, 512)])
def test_ctc_matches_oracle(cuda_device, shape):
    B, T1, T2 = shape
    tl, ml = synth.lengths(B, T2, T1, True, 91 + T2)
    tl = np.minimum(tl, ml)
    x = realistic_logits(B, T1, T2, tl, ml, 92 + T1)
    nll, g, ref, gref = run_both(x, tl, ml, cuda_device)
    assert torch.allclose(nll, ref, rtol=2e-5, atol=1e-4), (nll, ref)
    for b in range(B):
        scale = gref[b].abs().max().item()
        assert (g[b] - gref[b]).abs().max().item() <= 1e-3 * scale, (b, (g[b] - gref[b]).abs().max().item(), scale)
    assert g[0, ml[0]:].abs().sum().item() == 0.0                         # frames past the utterance carry no gradient


def test_ctc_noise_and_degenerate(cuda_device):
    """N(0,1) logits (no ridge: the mass spreads over many paths), one token, T1 == T2 (a single path), and an impossible
    utterance (mel_len < text_len): nll = +inf here, 0 in the reference's zero_infinity loss, and no gradient."""
    B, T1, T2 = 5, 48, 20
    tl = np.array([20, 1, 12, 16, 7], dtype=np.int64)
    ml = np.array([48, 30, 12, 9, 40], dtype=np.int64)                     # utterance 3 is impossible
    x = synth.noise_logits(B, T1, T2, 7)
    nll, g, ref, gref = run_both(x, tl, ml, cuda_device)
    assert torch.isinf(nll[3]) and ref[3].item() == 0.0
    ok = [0, 1, 2, 4]
    assert torch.allclose(nll[ok], ref[ok], rtol=2e-5, atol=1e-4), (nll, ref)
    assert g[3].abs().sum().item() == 0.0
    for b in ok:
        assert (g[b] - gref[b]).abs().max().item() <= 1e-3 * gref[b].abs().max().item()


def test_ctc_module_matches_reference_reduction(cuda_device):
    B, T1, T2 = 6, 120, 30
    tl, ml = synth.lengths(B, T2, T1, True, 5)
    tl = np.minimum(tl, ml)
    x = realistic_logits(B, T1, T2, tl, ml, 6)
    loss = AttentionCTCLoss()(torch.from_numpy(x).to(cuda_device), torch.from_numpy(tl).to(cuda_device), torch.from_numpy(ml).to(cuda_device))
    ref = octc.attention_ctc_loss(torch.from_numpy(x), torch.from_numpy(tl), torch.from_numpy(ml))
    assert abs(loss.item() - ref.item()) <= 2e-5 * abs(ref.item()) + 1e-5


def test_ctc_full_size_cfg3_shapes(cuda_device):
    """BASELINE cfg3 shapes (<= 200 tokens x <= 1000 frames, ragged, one utterance at the maximum), a batch the float64
    oracle finishes in seconds."""
    B, T1, T2 = 8, 1000, 200
    tl, ml = synth.lengths(B, T2, T1, True, 1237)
    x = realistic_logits(B, T1, T2, tl, ml, 1238)
    nll, g, ref, gref = run_both(x, tl, ml, cuda_device)
    assert torch.allclose(nll, ref, rtol=2e-5, atol=1e-4), (nll, ref)
    for b in range(B):
        assert (g[b] - gref[b]).abs().max().item() <= 1e-3 * gref[b].abs().max().item()
