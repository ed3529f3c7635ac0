"""isp_gemm_batched (tcgen05 batched ragged GEMM, csrc/isp_gemm.cu) against float64 matmul: every operand layout (K-major /
MN-major, shared operands), both operand types, ragged lengths, bf16 output, activations, the column statistics and the
implicit convolution against torch's Conv1d."""
import numpy as np
import pytest
import torch

from isp_tts_b200.gemm import bgemm, conv1d_channels_last

pytestmark = pytest.mark.gpu

# bf16 operands are compared at the operands' precision (inputs rounded to bf16 first): what is left is fp32 accumulation
# order.  fp32 operands go through TF32 products (10-bit mantissa): 2e-3 of the row's scale.
TOL = {torch.bfloat16: 2e-5, torch.float16: 2e-5, torch.float32: 3e-3}


def rnd(shape, dev, dtype, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(shape, generator=g).to(dtype).to(dev)


def check(got, ref, dtype, what, out_bf16=False):
    scale = ref.abs().max().item() + 1e-30
    err = (got.double() - ref).abs().max().item() / scale
    tol = TOL[dtype] + (8e-3 if out_bf16 else 0.0)
    assert err <= tol, f"{what}: relative error {err:.3e} > {tol:.1e}"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
@pytest.mark.parametrize("shape", [(3, 200, 128, 1000), (2, 1000, 384, 200), (4, 130, 72, 40), (1, 128, 256, 64), (2, 77, 520, 136)])
def test_layouts_against_float64(cuda_device, dtype, ta, tb, shape):
    B, M, N, K = shape
    a = rnd((B, K, M) if ta else (B, M, K), cuda_device, dtype, 1)
    b = rnd((B, N, K) if tb else (B, K, N), cuda_device, dtype, 2)
    x = a.transpose(1, 2) if ta else a
    y = b.transpose(1, 2) if tb else b
    got = bgemm(x, y, alpha=0.5)
    ref = 0.5 * torch.matmul(x.double(), y.double())
    assert got.shape == ref.shape and got.dtype == torch.float32
    check(got, ref, dtype, f"{shape} ta={ta} tb={tb}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("bn", [0, 64, 128, 192, 256])
def test_tile_widths_and_shared_operand(cuda_device, dtype, bn):
    B, M, N, K = 3, 260, 200, 96
    a = rnd((B, M, K), cuda_device, dtype, 3)
    w = rnd((N, K), cuda_device, dtype, 4)                      # one weight for the whole batch, K contiguous
    got = bgemm(a, w.t(), bn=bn)
    check(got, torch.matmul(a.double(), w.double().t()), dtype, f"shared B, bn={bn}")
    w2 = rnd((K, N), cuda_device, dtype, 5)                     # the same, N contiguous
    got = bgemm(a, w2, bn=bn)
    check(got, torch.matmul(a.double(), w2.double()), dtype, f"shared MN-major B, bn={bn}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_ragged_lengths_and_bf16_output(cuda_device, dtype):
    B, M, N, K = 5, 300, 150, 210
    ml = torch.tensor([300, 1, 129, 128, 40])
    nl = torch.tensor([150, 150, 7, 65, 128])
    kl = torch.tensor([210, 64, 33, 1, 200])
    a = rnd((B, M, K), cuda_device, dtype, 6)
    b = rnd((B, K, N), cuda_device, dtype, 7)
    for i in range(B):                                            # the k_len contract: one operand is zero past it
        b[i, kl[i]:] = 0
    got = bgemm(a, b, m_len=ml, n_len=nl, k_len=kl, out_dtype=torch.bfloat16)
    assert got.dtype == torch.bfloat16 and got.shape == (B, M, N)
    ref = torch.matmul(a.double(), b.double())
    for i in range(B):
        ref[i, ml[i]:] = 0
        ref[i, :, nl[i]:] = 0
        assert got[i, ml[i]:].abs().sum().item() == 0 and got[i, :, nl[i]:].abs().sum().item() == 0
    check(got, ref, dtype, "ragged, bf16 out", out_bf16=True)
    # poisoned padding of A beyond m_len must not leak into valid rows, and NaN there must not appear in C
    a2 = a.clone()
    for i in range(B):
        a2[i, ml[i]:] = float("nan")
    got2 = bgemm(a2, b, m_len=ml, n_len=nl, k_len=kl)
    assert torch.isfinite(got2).all()
    check(got2, ref, dtype, "ragged, poisoned rows")


@pytest.mark.parametrize("act", ["relu", "gelu"])
def test_activation_and_column_statistics(cuda_device, act):
    B, M, N, K = 3, 333, 160, 80
    ml = torch.tensor([333, 200, 31])
    a = rnd((B, M, K), cuda_device, torch.bfloat16, 8)
    w = rnd((N, K), cuda_device, torch.bfloat16, 9) * 0.2
    got, stats = bgemm(a, w.t(), m_len=ml, act=act, out_dtype=torch.bfloat16, col_stats=True)
    z = torch.matmul(a.double(), w.double().t())
    ref = torch.relu(z) if act == "relu" else torch.nn.functional.gelu(z)
    for i in range(B):
        ref[i, ml[i]:] = 0
    check(got, ref, torch.bfloat16, act, out_bf16=True)
    # the statistics are those of the values as stored (bf16), summed over the valid rows
    s = stats.double().sum(dim=1)
    g64 = got.double()
    assert torch.allclose(s[..., 0], g64.sum(dim=1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[..., 1], (g64 * g64).sum(dim=1), rtol=1e-4, atol=1e-2)
    a16, w16 = a.to(torch.float16), w.to(torch.float16)
    got16, stats16 = bgemm(a16, w16.t(), m_len=ml, act=act, out_dtype=torch.float16, col_stats=True)
    assert got16.dtype == torch.float16
    check(got16, ref, torch.float16, act + " fp16", out_bf16=True)
    s = stats16.double().sum(dim=1)
    assert torch.allclose(s[..., 0], got16.double().sum(dim=1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[..., 1], (got16.double() ** 2).sum(dim=1), rtol=1e-4, atol=1e-2)
    got32, stats32 = bgemm(a, w.t(), m_len=ml, act=act, col_stats=True)
    s = stats32.double().sum(dim=1)
    assert torch.allclose(s[..., 0], got32.double().sum(dim=1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[..., 1], (got32.double() ** 2).sum(dim=1), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
@pytest.mark.parametrize("cfg", [(3, 200, 384, 768, 5), (2, 333, 80, 160, 5), (2, 150, 160, 80, 3), (2, 64, 768, 128, 1)])
def test_implicit_convolution_against_conv1d(cuda_device, dtype, cfg):
    """alignment.py:58-62: Conv1d(kernel k, padding (k - 1) / 2, no bias) on the masked input, channels-last here."""
    B, T, Cin, Cout, k = cfg
    lens = torch.tensor([T, max(1, T // 2 + 3), 1][:B])
    x = rnd((B, T, Cin), cuda_device, dtype, 10)
    for i in range(B):
        x[i, lens[i]:] = 0                                          # ConvBlock1D masks its input (alignment.py:75-76)
    w = rnd((Cout, Cin, k), cuda_device, dtype, 11) / (Cin * k) ** 0.5
    got = conv1d_channels_last(x, w.permute(2, 0, 1).contiguous(), lens, out_dtype=torch.float32)
    ref = torch.nn.functional.conv1d(x.double().transpose(1, 2), w.double(), padding=(k - 1) // 2).transpose(1, 2).clone()
    for i in range(B):
        ref[i, lens[i]:] = 0
    check(got, ref, dtype, f"conv {cfg}")


def test_rejects_bad_arguments(cuda_device):
    a = torch.zeros((2, 16, 16), device=cuda_device)
    with pytest.raises(ValueError):
        bgemm(a, a.to(torch.bfloat16))
    with pytest.raises(ValueError):
        bgemm(a, torch.zeros((2, 8, 16), device=cuda_device))
    with pytest.raises(Exception):
        bgemm(torch.zeros((2, 16, 16)), torch.zeros((2, 16, 16)))           # CPU tensors: no fallback
