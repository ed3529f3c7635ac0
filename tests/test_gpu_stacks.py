"""The projection stacks on the sm_100a kernels against the torch restatement of
ConvBlock1D (alignment.py:40-83) evaluated in float64, at the recipe's shapes and at odd ones; the two element-wise
companions on their own; and the whole Aligner with fused stacks against the reference's golden outputs."""
import numpy as np
import pytest
import torch

from conftest import golden
from isp_tts_b200 import Aligner, synth
from isp_tts_b200 import stacks
from isp_tts_b200.gemm import conv1d_channels_last

pytestmark = pytest.mark.gpu


def make(dev, seed=2024, **over):
    hp = dict(synth.RECIPE_HP)
    hp.update(over)
    al = Aligner(**hp).eval()
    if not over:
        al.load_state_dict({k: torch.from_numpy(v) for k, v in synth.recipe_state(seed).items()}, strict=True)
    return al.to(dev)


@pytest.mark.parametrize("layout", ["channels_first", "channels_last"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_prep_channels_last(cuda_device, layout, dtype):
    B, C, T = 3, 80, 333
    lens = torch.tensor([333, 100, 1], device=cuda_device)
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn((B, C, T), generator=g).to(cuda_device)
    src = x if layout == "channels_first" else x.transpose(1, 2).contiguous()
    if C == T:
        pytest.skip("ambiguous")
    out = stacks.prep_channels_last(src, lens, C, dtype)
    ref = (x * (torch.arange(T, device=cuda_device)[None, None, :] < lens[:, None, None])).transpose(1, 2).to(dtype)
    assert out.shape == (B, T, C) and out.dtype == dtype and torch.equal(out, ref)
    # channel padding to a whole 16 B vector is zero-filled
    out = stacks.prep_channels_last(src[:, :77] if layout == "channels_first" else src[:, :, :77].contiguous(), lens, 77, dtype)
    cp = 80
    assert out.shape == (B, T, cp) and torch.equal(out[:, :, :77], ref[:, :, :77]) and out[:, :, 77:].abs().sum().item() == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_conv_block_against_float64(cuda_device, dtype):
    """One ConvBlock1D: conv(k=5) -> GELU -> masked instance norm, re-masked (alignment.py:75-82)."""
    from isp_tts_b200.alignment import ConvBlock1D
    B, T, Cin, Cout = 3, 200, 384, 768
    lens = torch.tensor([200, 97, 3], device=cuda_device)
    torch.manual_seed(5)
    blk = ConvBlock1D(Cin, Cout, kernel_size=5, bias=False, activation="gelu", normalization="instance").to(cuda_device).eval()
    with torch.no_grad():
        blk.norm.weight.normal_(1.0, 0.1)
        blk.norm.bias.normal_(0.0, 0.1)
    x = torch.randn((B, Cin, T), device=cuda_device)
    mask = (torch.arange(T, device=cuda_device)[None, None, :] < lens[:, None, None])
    h = stacks.prep_channels_last(x, lens, Cin, dtype)
    w = stacks._taps(blk.conv, dtype, Cin)
    y, st = conv1d_channels_last(h, w, lens, act="gelu", out_dtype=dtype, col_stats=True)
    y = stacks.instance_norm_apply(y, st, blk.norm, lens)
    # float64 restatement on the operands as the kernel sees them (rounded to the operand type)
    xd = (x * mask).to(dtype).double()
    wd = blk.conv.weight.detach().to(dtype).double()
    z = torch.nn.functional.gelu(torch.nn.functional.conv1d(xd, wd, padding=2))
    if dtype != torch.float32:
        z = z.to(dtype).double()                               # the activation is stored in the 2-byte type before the norm
    n = lens[:, None, None].double()
    mean = (z * mask).sum(2, keepdim=True) / n
    var = (((z * mask - mean) * mask) ** 2).sum(2, keepdim=True) / n
    ref = ((z - mean) / (var + 1e-5).sqrt() * blk.norm.weight.double()[None, :, None] + blk.norm.bias.double()[None, :, None]) * mask
    err = (y.double().transpose(1, 2) - ref).abs().max().item()
    tol = {torch.float32: 3e-2, torch.bfloat16: 6e-2, torch.float16: 8e-3}[dtype]              # on normalised values of up to ~4: TF32 products over K = 1920 / bf16 storage
    assert err <= tol, err
    assert y[1, 97:].abs().sum().item() == 0


@pytest.mark.parametrize("mode,inner,tol", [("fp32", None, 2e-2), ("bf16", torch.bfloat16, 8e-2), ("bf16", torch.float16, 2e-2)])
def test_recipe_stacks_against_torch_path(cuda_device, mode, inner, tol):
    """q, k of the fused stacks against the torch restatement (fp32, same module, fused_stacks off), recipe shapes, ragged."""
    al = make(cuda_device)
    al.attention.gemm_dtype = mode
    al.attention.fused_stacks = True                            # "auto" keeps fp32 mode on the torch ops
    if inner is not None:
        al.attention.stack_dtype = inner
    B, T1, T2 = 4, 500, 120
    tl = np.array([120, 77, 120, 20], np.int64); ml = np.array([500, 333, 129, 100], np.int64)
    mel, txt = synth.recipe_inputs(99, B, T1, T2, tl, ml)
    args = [torch.from_numpy(a).to(cuda_device) for a in (mel, txt, ml, tl)]
    with torch.no_grad():
        q, k = al.attention.encode(*args)
        al.attention.fused_stacks = False
        q0, k0 = al.attention.encode(*args)
    assert q.shape == q0.shape and k.shape == k0.shape
    assert q.dtype == (torch.bfloat16 if mode == "bf16" else torch.float32)
    for got, ref, name in ((q, q0, "q"), (k, k0, "k")):
        err = (got.float() - ref).abs().max().item() / ref.abs().max().item()
        assert err <= tol, f"{name}: {err:.3e}"
    for b in range(B):
        assert q[b, ml[b]:].abs().sum().item() == 0 and k[b, tl[b]:].abs().sum().item() == 0


def test_unsupported_blocks_take_the_torch_path(cuda_device):
    al = make(cuda_device, normalization="batch")
    assert not stacks.fused_supported(al.attention.key_proj, torch.float32)
    al = make(cuda_device, activation="tanh")
    assert not stacks.fused_supported(al.attention.key_proj, torch.float32)
    al = make(cuda_device)
    assert stacks.fused_supported(al.attention.key_proj, torch.bfloat16) and stacks.fused_supported(al.attention.query_proj, torch.float32)
    # with gradients enabled the torch path runs and autograd reaches the parameters
    B, T1, T2 = 2, 64, 16
    tl = np.array([16, 9], np.int64); ml = np.array([64, 40], np.int64)
    mel, txt = synth.recipe_inputs(3, B, T1, T2, tl, ml)
    out = al.train()(torch.from_numpy(mel).to(cuda_device), torch.from_numpy(txt).to(cuda_device),
                     torch.from_numpy(ml).to(cuda_device), torch.from_numpy(tl).to(cuda_device))
    out.attn_logits.mean().backward()
    assert all(p.grad is not None for p in al.parameters())
