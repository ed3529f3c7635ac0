"""Full Aligner.forward drop-in on the GPU against the reference's own four outputs
(tests/golden/loglik_*.npz were produced by the reference Aligner with the same state_dict)."""
import numpy as np
import pytest
import torch

from conftest import golden
from isp_tts_b200 import Aligner, AlignerOutput
from oracle import loglik as oll
from oracle import mas as omas

pytestmark = pytest.mark.gpu


def build(g, dev):
    hp = {k: eval(v) for k, v in zip(g["hp_keys"], g["hp_vals"])}
    al = Aligner.init(config=hp).eval()
    al.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=True)
    return al.to(dev)


@pytest.mark.parametrize("mode", ["auto", "tf32"])
@pytest.mark.parametrize("tag", ["small", "dim80", "dim128"])
def test_forward_matches_reference(cuda_device, tag, mode):
    """mode "auto" outside autocast = the reference's fp32 products (3xTF32 split); "tf32" = the fused single-pass kernel."""
    g = golden(f"loglik_{tag}.npz")
    al = build(g, cuda_device)
    al.attention.gemm_dtype = mode
    mel = torch.from_numpy(g["mel"]).to(cuda_device)
    txt = torch.from_numpy(g["enc_text"]).to(cuda_device)
    ml = torch.from_numpy(g["mel_len"]).to(cuda_device)
    tl = torch.from_numpy(g["text_len"]).to(cuda_device)
    with torch.no_grad():
        out = al(mel, txt, ml, tl)                                   # model.py:138-141 call shape
    assert isinstance(out, AlignerOutput)
    soft, logits, hard, dur = (t.cpu().numpy() for t in out)
    assert soft.dtype == np.float32 and logits.dtype == np.float32 and hard.dtype == np.int16 and dur.dtype == np.int64
    assert hard.shape == logits.shape == soft.shape == g["attn_logits"].shape and dur.shape == g["durations"].shape
    _, _, parts = oll.loglik(g["Q"], g["K"], g["text_len"], g["mel_len"], return_parts=True)
    ok = ~oll.threshold_ambiguous(parts["prior_raw"])
    err = np.abs(logits - g["attn_logits"])
    # (both modes: the projection stacks in front are torch / cuDNN convolutions, TF32 by torch's default exactly as in the reference
    # on a GPU -- the golden comes from a CPU run; tests/test_gpu_loglik.py::test_fp32_faithful_products bounds the contraction itself)
    assert np.all(err[ok] <= 1e-3 * np.abs(g["attn_logits"][ok]) + 1e-4), err[ok].max()
    # MAS is bit-exact GIVEN the logits: oracle on OUR logits == our path
    rh, rd = omas.b_mas_with_durations(logits, g["text_len"], g["mel_len"])
    assert np.array_equal(hard, rh) and np.array_equal(dur, rd)
    assert np.array_equal(dur, hard.sum(axis=1, dtype=np.int64))     # alignment.py:275
    assert np.array_equal(dur.sum(1), g["mel_len"])
    # and the reference's own path differs from ours on at most a few frames (TF32 products)
    frames = int(g["mel_len"].sum())
    moved = int((hard != g["attn_hard"]).any(axis=2).sum())
    assert moved <= max(2, frames // 100), f"{moved} of {frames} frames moved"


def test_autocast_uses_bf16_and_grads_flow(cuda_device):
    g = golden("loglik_dim80.npz")
    al = build(g, cuda_device).train()
    mel = torch.from_numpy(g["mel"]).to(cuda_device)
    txt = torch.from_numpy(g["enc_text"]).to(cuda_device)
    ml = torch.from_numpy(g["mel_len"]).to(cuda_device)
    tl = torch.from_numpy(g["text_len"]).to(cuda_device)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = al(mel, txt, ml, tl)
    assert out.attn_logits.dtype == torch.float32 and out.attn_logits.requires_grad
    loss = -(out.attn_soft[out.attn_hard == 1] + 1e-8).log().mean() + out.attn_logits.mean() * 1e-3   # loss.py:97-105 shape
    loss.backward()
    grads = [p.grad for p in al.parameters()]
    assert all(gr is not None and torch.isfinite(gr).all() for gr in grads)
    assert sum(float(gr.abs().sum()) for gr in grads) > 0


def test_recipe_shape_matches_reference(cuda_device):
    """The Aligner at the recipe's hyper-parameters (recipes/acoustic/core.yaml:150-156: mel 80, text 384, attention_dim 128,
    kernels 5, gelu, instance norm; 1.7 M parameters) against the reference's outputs for the same seeded weights and inputs
    (oracle/gen_golden.py::gen_recipe froze only the outputs; isp_tts_b200.synth regenerates weights and inputs)."""
    from isp_tts_b200 import synth
    g = golden("aligner_recipe.npz")
    seed, B, T1, T2 = int(g["seed"]), int(g["B"]), int(g["T1"]), int(g["T2"])
    tl, ml = g["text_len"], g["mel_len"]
    al = Aligner(**synth.RECIPE_HP).eval()
    assert {k: tuple(v.shape) for k, v in al.state_dict().items()} == synth.RECIPE_SHAPES        # same keys, same shapes
    assert sum(v.numel() for v in al.state_dict().values()) == 1_713_120
    al.load_state_dict({k: torch.from_numpy(v) for k, v in synth.recipe_state(seed).items()}, strict=True)
    al = al.to(cuda_device)
    mel, txt = synth.recipe_inputs(seed + 1, B, T1, T2, tl, ml)
    with torch.no_grad():
        out = al(torch.from_numpy(mel).to(cuda_device), torch.from_numpy(txt).to(cuda_device),
                 torch.from_numpy(ml).to(cuda_device), torch.from_numpy(tl).to(cuda_device))
    soft, logits, hard, dur = (t.cpu().numpy() for t in out)
    ref = g["attn_logits"]
    valid = (np.arange(T1)[None, :, None] < ml[:, None, None]) & (np.arange(T2)[None, None, :] < tl[:, None, None])
    # cells within 2 % of the prior's 1e-4 threshold may fall on either side of it (a jump of ~4.6 in the logit): leave out
    # rows' cells whose reference value and ours differ by that jump, but bound how many there are
    err = np.abs(logits - ref)
    tol = 1e-3 * np.abs(ref) + 2e-3            # TF32 products through two conv stacks
    bad = (err > tol) & valid
    assert bad.sum() <= max(4, valid.sum() // 2000), (int(bad.sum()), float(err[valid].max()))
    assert np.abs(soft - g["attn_soft"].astype(np.float32))[valid].max() < 5e-3
    assert soft[~valid].sum() == 0.0
    rh, rd = omas.b_mas_with_durations(logits, tl, ml)                       # bit-exact given our logits
    assert np.array_equal(hard, rh) and np.array_equal(dur, rd)
    path = hard.argmax(axis=2)
    frames = int(ml.sum())
    moved = sum(int((path[b, :ml[b]] != g["path"][b, :ml[b]]).sum()) for b in range(B))
    assert moved <= max(2, frames // 100), f"{moved} of {frames} frames moved against the reference's path"
    assert np.array_equal(dur.sum(1), ml)
