"""GPU parity of the fused log-likelihood kernel (tcgen05 GEMM + epilogue) against the golden
outputs of the reference's ConvAttention.forward and against the pinned numpy oracle.

Tolerances (north star: "log-likelihood within 1e-3 relative in fp32, bf16 stated separately"):
  fp32 operands (TF32 products, fp32 accumulate):  |d logits| <= 1e-3 * |logits| + 1e-4,  |d soft| <= 1e-3
  bf16 operands:                                   |d logits| <= 5e-3 * |logits| + 2e-2,  |d soft| <= 2e-2
Cells whose prior lies within 2e-5 (relative) of the reference's hard 1e-4 threshold
(alignment.py:35) may land on either side; they are excluded and counted (must be < 0.1 %).
"""
import numpy as np
import pytest
import torch

from conftest import golden
from isp_tts_b200 import _lib, synth
from isp_tts_b200.alignment import loglik_forward
from oracle import loglik as oll

pytestmark = pytest.mark.gpu

TOL = {"fp32": dict(rel=1e-3, abs=1e-4, soft=1e-3), "bf16": dict(rel=5e-3, abs=2e-2, soft=2e-2)}


def run(q, k, tl, ml, dev, dtype="fp32", prior=True):
    td = torch.float32 if dtype == "fp32" else torch.bfloat16
    qt = torch.from_numpy(q).to(dev).to(td)
    kt = torch.from_numpy(k).to(dev).to(td)
    soft, logits = loglik_forward(qt, kt, torch.from_numpy(np.asarray(tl)).to(dev),
                                  torch.from_numpy(np.asarray(ml)).to(dev), attention_prior=prior)
    torch.cuda.synchronize()
    return soft.cpu().numpy(), logits.cpu().numpy()


def compare(soft, logits, ref_soft, ref_logits, ambiguous, tol, what):
    assert np.all(np.isfinite(logits)) and np.all(np.isfinite(soft)), what
    ok = ~ambiguous
    assert ambiguous.mean() < 1e-3, f"{what}: {ambiguous.sum()} cells on the prior threshold"
    err = np.abs(logits - ref_logits)
    bound = tol["rel"] * np.abs(ref_logits) + tol["abs"]
    if not np.all(err[ok] <= bound[ok]):
        bad = np.argwhere((err > bound) & ok)
        b, i, j = bad[0]
        raise AssertionError(f"{what}: {len(bad)} logits out of tolerance; first (b={b}, i={i}, j={j}) "
                             f"cuda={logits[b, i, j]:.6f} ref={ref_logits[b, i, j]:.6f}; max err {err[ok].max():.3e}")
    # a flipped cell changes its row's normaliser: compare soft on rows without ambiguous cells
    rows_ok = ~ambiguous.any(axis=2, keepdims=True)
    serr = np.abs(soft - ref_soft) * rows_ok
    assert serr.max() <= tol["soft"], f"{what}: attn_soft max err {serr.max():.3e}"
    return float(err[ok].max()), float(serr.max())


@pytest.mark.parametrize("tag", ["small", "dim80", "dim128"])
def test_scores_only(cuda_device, tag):
    """The GEMM alone (debug option): scale * Q.K^T against fp64 numpy."""
    g = golden(f"loglik_{tag}.npz")
    _lib.set_option("loglik.debug_scores", 1)
    try:
        _, s = run(g["Q"], g["K"], g["text_len"], g["mel_len"], cuda_device)
    finally:
        _lib.set_option("loglik.debug_scores", 0)
    ref = (g["Q"].astype(np.float64) @ g["K"].astype(np.float64).transpose(0, 2, 1)) * g["Q"].shape[2] ** -0.5
    err = np.abs(s - ref)
    assert err.max() < 2e-3 * max(1.0, np.abs(ref).max()), f"{tag}: max |dS| = {err.max():.3e} (|S| max {np.abs(ref).max():.3f})"


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", ["small", "dim80", "dim128"])
def test_golden_reference_outputs(cuda_device, tag, dtype):
    g = golden(f"loglik_{tag}.npz")
    soft, logits = run(g["Q"], g["K"], g["text_len"], g["mel_len"], cuda_device, dtype)
    _, _, parts = oll.loglik(g["Q"], g["K"], g["text_len"], g["mel_len"], return_parts=True)
    amb = oll.threshold_ambiguous(parts["prior_raw"])
    compare(soft, logits, g["attn_soft"], g["attn_logits"], amb, TOL[dtype], f"{tag}/{dtype}")
    # structure: padded frames and padded tokens carry no soft mass, valid rows sum to 1
    mask = parts["mask"]
    assert np.all(soft[~np.broadcast_to(mask, soft.shape)] == 0)
    rows = np.arange(soft.shape[1])[None, :] < g["mel_len"][:, None]
    assert np.allclose(soft.sum(2)[rows], 1.0, atol=1e-4)


@pytest.mark.parametrize("shape", [(6, 300, 100, 128), (5, 130, 201, 80), (3, 700, 256, 128), (2, 129, 8, 16), (2, 200, 600, 64)])
def test_fp32_faithful_products(cuda_device, shape):
    """precision="fp32": the 3xTF32 split (isp_split_3xtf32 + one TF32 contraction over 3 D + isp_loglik_rows) -- what the
    drop-in runs outside autocast, where the reference's matmul is true fp32.  Scores within 2e-6 of |Q||K| against float64
    (one TF32 product: ~5e-4), attn_logits within 2e-5 relative + 2e-5 of the oracle, 50x tighter than the TF32 tolerance;
    gradients flow through the same precision."""
    B, T1, T2, D = shape
    tl, ml = synth.lengths(B, T2, T1, True, 77 + T2)
    q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, 99 + T1)
    dev = cuda_device
    qt, kt = torch.from_numpy(q).to(dev), torch.from_numpy(k).to(dev)
    from isp_tts_b200.alignment import _scores_fp32, _scores
    s3 = _scores_fp32(qt, kt, None, None).cpu().numpy().astype(np.float64)
    s1 = _scores(qt, kt).cpu().numpy().astype(np.float64)
    ref = q.astype(np.float64) @ k.astype(np.float64).transpose(0, 2, 1)
    norm = np.linalg.norm(q.astype(np.float64), axis=2)[:, :, None] * np.linalg.norm(k.astype(np.float64), axis=2)[:, None, :] + 1e-30
    e3, e1 = float((np.abs(s3 - ref) / norm).max()), float((np.abs(s1 - ref) / norm).max())
    assert e3 < 2e-6 and e3 < e1 / 20, (e3, e1)
    qg, kg = qt.clone().requires_grad_(True), kt.clone().requires_grad_(True)
    soft, logits = loglik_forward(qg, kg, torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev), precision="fp32")
    rs, rl, parts = oll.loglik(q, k, tl, ml, return_parts=True)
    amb = oll.threshold_ambiguous(parts["prior_raw"])
    compare(soft.detach().cpu().numpy(), logits.detach().cpu().numpy(), rs, rl, amb, dict(rel=2e-5, abs=2e-5, soft=2e-5), f"{shape}/fp32 faithful")
    if T2 > 512:
        return                                       # (isp_loglik_backward_ds stops at 512 tokens)
    (soft * soft).sum().backward()
    assert torch.isfinite(qg.grad).all() and torch.isfinite(kg.grad).all() and float(qg.grad.abs().sum()) > 0


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(6, 300, 100, 128), (5, 130, 201, 80), (3, 700, 256, 128), (2, 129, 8, 16)])
def test_against_oracle(cuda_device, shape, dtype):
    B, T1, T2, D = shape
    tl, ml = synth.lengths(B, T2, T1, True, 77 + T2)
    q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, 99 + T1)
    if dtype == "bf16":   # compare at the operands' precision: round first, then both sides see the same numbers
        q = torch.from_numpy(q).to(torch.bfloat16).float().numpy()
        k = torch.from_numpy(k).to(torch.bfloat16).float().numpy()
    soft, logits = run(q, k, tl, ml, cuda_device, dtype)
    rs, rl, parts = oll.loglik(q, k, tl, ml, return_parts=True)
    amb = oll.threshold_ambiguous(parts["prior_raw"])
    tol = TOL["fp32"] if dtype == "fp32" else dict(rel=1e-3, abs=1e-4, soft=1e-3)   # bf16 inputs are exact products
    compare(soft, logits, rs, rl, amb, tol, f"{shape}/{dtype}")


def test_long_text_two_accumulator_chunks(cuda_device):
    """256 < T2max <= 512 uses all 512 TMEM columns (bf16 operands)."""
    B, T1, T2, D = 2, 260, 512, 128
    tl, ml = np.array([512, 300]), np.array([260, 140])
    q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, 5)
    q = torch.from_numpy(q).to(torch.bfloat16).float().numpy()
    k = torch.from_numpy(k).to(torch.bfloat16).float().numpy()
    soft, logits = run(q, k, tl, ml, cuda_device, "bf16")
    rs, rl, parts = oll.loglik(q, k, tl, ml, return_parts=True)
    compare(soft, logits, rs, rl, oll.threshold_ambiguous(parts["prior_raw"]), dict(rel=1e-3, abs=1e-4, soft=1e-3), "T2=512")


def test_without_prior(cuda_device):
    g = golden("loglik_dim80.npz")
    soft, logits = run(g["Q"], g["K"], g["text_len"], g["mel_len"], cuda_device, prior=False)
    rs, rl = oll.loglik(g["Q"], g["K"], g["text_len"], g["mel_len"], attention_prior=False)
    assert np.abs(logits - rl).max() < 2e-3 and np.abs(soft - rs).max() < 1e-3


def test_unsupported_inputs_raise(cuda_device):
    one = torch.ones(1, dtype=torch.int64, device=cuda_device)
    with pytest.raises(_lib.IspError):
        loglik_forward(torch.zeros(1, 8, 16), torch.zeros(1, 8, 16), one.cpu(), one.cpu())   # CPU tensors: no fallback
    with pytest.raises(ValueError):
        loglik_forward(torch.zeros(1, 8, 16, device=cuda_device), torch.zeros(1, 8, 24, device=cuda_device), one, one)


@pytest.mark.parametrize("shape,dtype", [((2, 300, 700, 64), "fp32"), ((2, 260, 1100, 128), "bf16"), ((3, 200, 48, 12), "fp32"),
                                         ((2, 150, 64, 264), "bf16"), ((1, 130, 513, 80), "fp32")])
def test_shapes_beyond_the_fused_kernel(cuda_device, shape, dtype):
    """More than 512 tokens (long-form text), attention_dim that is not a multiple of 8 or exceeds 256: the batched GEMM plus
    the stand-alone row epilogue (isp_loglik_rows) take over -- same tolerances as the fused kernel; MAS follows on the same
    logits (its general kernel beyond 640 tokens), bit-exact against the oracle."""
    from isp_tts_b200.mas import mas_forward
    from oracle import mas as omas
    B, T1, T2, D = shape
    tl, ml = synth.lengths(B, T2, T1, True, 500 + T2)
    q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, 501 + T2)
    if dtype == "bf16":
        q = torch.from_numpy(q).to(torch.bfloat16).float().numpy()
        k = torch.from_numpy(k).to(torch.bfloat16).float().numpy()
    soft, logits = run(q, k, tl, ml, cuda_device, dtype)
    rs, rl, parts = oll.loglik(q, k, tl, ml, return_parts=True)
    tol = TOL["fp32"] if dtype == "fp32" else dict(rel=1e-3, abs=1e-4, soft=1e-3)
    compare(soft, logits, rs, rl, oll.threshold_ambiguous(parts["prior_raw"]), tol, f"general path {shape} {dtype}")
    hard, dur = mas_forward(torch.from_numpy(logits).to(cuda_device), torch.from_numpy(tl), torch.from_numpy(ml))
    rh, rd = omas.b_mas_with_durations(logits, tl, ml)
    assert np.array_equal(hard.cpu().numpy(), rh) and np.array_equal(dur.cpu().numpy(), rd)
    s2, l2 = run(q, k, tl, ml, cuda_device, dtype, prior=False)
    r2s, r2l = oll.loglik(q, k, tl, ml, attention_prior=False)
    assert np.abs(l2 - r2l).max() < 5e-3 and np.abs(s2 - r2s).max() < 2e-3


@pytest.mark.parametrize("T2", [24, 37])
def test_backward_matches_torch_autograd(cuda_device, T2):
    """Gradients of (attn_soft, attn_logits) w.r.t. q, k against a plain torch restatement, on both routes of the backward pass:
    from attn_logits and the saved prior row sums (T2 = 24) and from recomputed scores (T2 = 37: a token axis that needs padding)."""
    B, T1, D = 3, 70, 32
    tl, ml = synth.lengths(B, T2, T1, True, 3)
    qn, kn = synth.encoded_pair(B, T1, T2, D, tl, ml, 4)
    tlt, mlt = torch.from_numpy(tl).to(cuda_device), torch.from_numpy(ml).to(cuda_device)
    q = torch.from_numpy(qn).to(cuda_device).requires_grad_(True)
    k = torch.from_numpy(kn).to(cuda_device).requires_grad_(True)
    soft, logits = loglik_forward(q, k, tlt, mlt)
    w1 = torch.randn_like(soft); w2 = torch.randn_like(logits)
    (soft * w1).sum().add((logits * w2).sum()).backward()
    gq, gk = q.grad.clone(), k.grad.clone()

    q2 = q.detach().clone().requires_grad_(True); k2 = k.detach().clone().requires_grad_(True)
    from isp_tts_b200.alignment import batch_diagonal_prior
    s = D ** -0.5 * torch.matmul(q2, k2.transpose(1, 2))
    prior = batch_diagonal_prior(tlt, mlt, max_text=T2, max_mel=T1)
    lg = torch.log_softmax(s, dim=2) + torch.log(prior + 1e-6)
    km = (torch.arange(T2, device=cuda_device)[None] < tlt[:, None])[:, None, :]
    qm = (torch.arange(T1, device=cuda_device)[None] < mlt[:, None])[:, :, None]
    sf = torch.softmax(lg.masked_fill(~km, -3.4028234663852886e38), dim=2) * (km & qm)
    (sf * w1).sum().add((lg * w2).sum()).backward()
    assert torch.allclose(gq, q2.grad, rtol=2e-3, atol=2e-3), (gq - q2.grad).abs().max()
    assert torch.allclose(gk, k2.grad, rtol=2e-3, atol=2e-3), (gk - k2.grad).abs().max()


def _torch_ds(s, soft, gl, gs, scale, prior):
    """dL/dS by the book (alignment.py:190-206 differentiated), fp64 on the GPU."""
    s, soft = s.double(), soft.double()
    g = torch.zeros_like(s) if gl is None else gl.double().clone()
    if gs is not None:
        g = g + soft * (gs.double() - (gs.double() * soft).sum(2, keepdim=True))
    if prior:
        g = g - torch.softmax(scale * s, dim=2) * g.sum(2, keepdim=True)
    return g * scale


@pytest.mark.parametrize("T2", [200, 24, 37, 512])
@pytest.mark.parametrize("which", ["both", "logits", "soft"])
def test_backward_ds_kernel(cuda_device, T2, which):
    """isp_loglik_backward_ds against the closed-form Jacobians, incl. a token axis that needs padding (37)."""
    from isp_tts_b200.alignment import loglik_backward_ds
    B, T1, D = 3, 130, 64
    tl, ml = synth.lengths(B, T2, T1, True, 7)
    qn, kn = synth.encoded_pair(B, T1, T2, D, tl, ml, 8)
    q, k = torch.from_numpy(qn).to(cuda_device), torch.from_numpy(kn).to(cuda_device)
    tlt, mlt = torch.from_numpy(tl).to(cuda_device), torch.from_numpy(ml).to(cuda_device)
    soft, logits = loglik_forward(q, k, tlt, mlt)
    s = torch.matmul(q, k.transpose(1, 2))
    gen = torch.Generator(device=cuda_device).manual_seed(5)
    gl = torch.randn(soft.shape, device=cuda_device, generator=gen) if which != "soft" else None
    gs = torch.randn(soft.shape, device=cuda_device, generator=gen) if which != "logits" else None
    scale = D ** -0.5
    for prior in (True, False):
        ref = _torch_ds(s, soft.detach(), gl, gs, scale, prior)
        out = loglik_backward_ds(s, soft.detach(), gl, gs, scale, prior)
        assert out.shape == ref.shape and out.dtype == torch.float32
        err = (out.double() - ref).abs().max().item()
        assert err <= 2e-5 * max(1.0, ref.abs().max().item()), (T2, which, prior, err)
        out16 = loglik_backward_ds(s, soft.detach(), gl, gs, scale, prior, out_dtype=torch.bfloat16)
        assert (out16.double() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("shape", [(3, 130, 200, 64), (5, 260, 24, 32), (2, 300, 512, 32), (40, 200, 100, 64)])
@pytest.mark.parametrize("which", ["both", "logits", "soft"])
def test_backward_from_logits_kernel(cuda_device, shape, which):
    """isp_loglik_backward_from_logits (no scores, no attn_soft: attn_logits and the prior's saved row sums) against the
    closed-form Jacobians in fp64 on the TRUE scores, with and without the prior, ragged lengths, padded frames and tokens
    included.  Tolerance: 1e-4 of the largest entry -- attn_logits carries the forward kernel's own rounding (tf32 products are
    NOT in it: the Jacobian is taken at the scores the forward actually used, recovered from its logits)."""
    from isp_tts_b200.alignment import _loglik_cuda, loglik_backward_from_logits
    B, T1, T2, D = shape
    tl, ml = synth.lengths(B, T2, T1, True, 7)
    qn, kn = synth.encoded_pair(B, T1, T2, D, tl, ml, 8)
    q, k = torch.from_numpy(qn).to(cuda_device), torch.from_numpy(kn).to(cuda_device)
    tlt, mlt = torch.from_numpy(tl).to(cuda_device), torch.from_numpy(ml).to(cuda_device)
    gen = torch.Generator(device=cuda_device).manual_seed(5)
    gl = torch.randn((B, T1, T2), device=cuda_device, generator=gen) if which != "soft" else None
    gs = torch.randn((B, T1, T2), device=cuda_device, generator=gen) if which != "logits" else None
    scale = D ** -0.5
    for prior in (True, False):
        soft, logits, rowsum = _loglik_cuda(q, k, tlt, mlt, scale, prior, want_rowsum=True)
        assert rowsum is not None and rowsum.shape == (B, T1)
        # the scores the forward kernel worked with, from its own output (so that its TF32 rounding is not counted as an error of
        # the backward kernel): scale * S = attn_logits - log(prior + 1e-6) + lse, and softmax is invariant to the per-row lse
        if prior:
            from isp_tts_b200.alignment import batch_diagonal_prior
            pr = batch_diagonal_prior(tlt, mlt, max_text=T2, max_mel=T1).double()
            s_eff = (logits.double() - torch.log(pr + 1e-6)) / scale
        else:
            s_eff = logits.double() / scale
        ref = _torch_ds(s_eff, soft, gl, gs, scale, prior)
        out = loglik_backward_from_logits(logits, gl, gs, rowsum, tlt, mlt, scale, prior)
        assert out.shape == ref.shape and out.dtype == torch.float32
        diff = (out.double() - ref).abs()
        # cells whose prior sits within rounding of the 1e-4 threshold may be thresholded differently by torch's fp32 prior
        if prior:
            raw = pr * 0 + batch_diagonal_prior(tlt, mlt, threshold=0.0, max_text=T2, max_mel=T1).double()
            amb = (raw - 1e-4).abs() < 2e-8
            rows_amb = amb.any(dim=2, keepdim=True).expand_as(diff)
            diff = diff.masked_fill(rows_amb, 0.0)
            assert float(rows_amb[:, :, 0].double().mean()) < 0.03
        err = diff.max().item()
        assert err <= 1e-4 * max(1.0, ref.abs().max().item()), (shape, which, prior, err)
        out16 = loglik_backward_from_logits(logits, gl, gs, rowsum, tlt, mlt, scale, prior, out_dtype=torch.bfloat16)
        d16 = (out16.double() - ref).abs()
        if prior:
            d16 = d16.masked_fill(rows_amb, 0.0)
        assert d16.max().item() <= 1e-2 * max(1.0, ref.abs().max().item())


def test_backward_bf16_operands(cuda_device):
    """bf16 Q, K: gradients arrive in bf16 and agree with the fp32 route to bf16 accuracy."""
    B, T1, T2, D = 2, 96, 40, 64
    tl, ml = synth.lengths(B, T2, T1, True, 9)
    qn, kn = synth.encoded_pair(B, T1, T2, D, tl, ml, 10)
    tlt, mlt = torch.from_numpy(tl).to(cuda_device), torch.from_numpy(ml).to(cuda_device)
    grads = {}
    for dt in (torch.float32, torch.bfloat16):
        q = torch.from_numpy(qn).to(cuda_device).to(dt).requires_grad_(True)
        k = torch.from_numpy(kn).to(cuda_device).to(dt).requires_grad_(True)
        soft, logits = loglik_forward(q, k, tlt, mlt)
        torch.manual_seed(0)
        w1 = torch.randn_like(soft); w2 = torch.randn_like(logits)
        (soft * w1).sum().add((logits * w2).sum()).backward()
        assert q.grad.dtype == dt and k.grad.dtype == dt
        grads[dt] = (q.grad.float(), k.grad.float())
    for a, b_ in zip(grads[torch.float32], grads[torch.bfloat16]):
        assert (a - b_).abs().max() <= 0.05 * a.abs().max() + 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stage_operands_ragged_upload(cuda_device, dtype):
    """isp_stage_operands: rows below the lengths arrive bit-exact, padding rows are zeros on the device whatever the
    host tensors hold there (the operand contract of alignment.py:75-76), and the log-likelihood of the staged operands
    equals the one of plainly copied operands."""
    from isp_tts_b200.alignment import stage_operands
    B, T1, T2, D = 5, 300, 70, 128
    tl, ml = synth.lengths(B, T2, T1, True, 77)
    q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, 78)
    qh = torch.from_numpy(q).to(dtype)
    kh = torch.from_numpy(k).to(dtype)
    clean_q, clean_k = qh.clone(), kh.clone()
    # poison the host padding: it must not reach the device
    qh[torch.arange(T1)[None, :] >= torch.from_numpy(ml)[:, None]] = 7.0
    kh[torch.arange(T2)[None, :] >= torch.from_numpy(tl)[:, None]] = -3.0
    qh, kh = qh.pin_memory(), kh.pin_memory()
    tlt, mlt = torch.from_numpy(tl).to(cuda_device), torch.from_numpy(ml).to(cuda_device)
    qd, kd = stage_operands(qh, kh, tlt, mlt)
    torch.cuda.synchronize()
    assert torch.equal(qd.cpu(), clean_q) and torch.equal(kd.cpu(), clean_k)
    s1, l1 = loglik_forward(qd, kd, tlt, mlt)
    s2, l2 = loglik_forward(clean_q.to(cuda_device), clean_k.to(cuda_device), tlt, mlt)
    assert torch.equal(s1, s2) and torch.equal(l1, l2)
    with pytest.raises(ValueError):
        stage_operands(clean_q, clean_k, tlt, mlt)            # not pinned


def _full_size_against_oracle(dev, name, dtype, chunk):
    """The kernel on a BASELINE.json workload at full size, compared utterance chunk by utterance chunk with the numpy
    oracle (oracle/loglik.py) so that the CPU side runs in seconds."""
    w = synth.WORKLOADS[name]
    tl, ml = synth.workload_lengths(w)
    q, k = synth.encoded_pair(w.batch, w.t1max, w.t2max, w.dim, tl, ml, w.seed + 1)
    if dtype == "bf16":                # compare at the operands' precision (see test_against_oracle)
        q = torch.from_numpy(q).to(torch.bfloat16).float().numpy()
        k = torch.from_numpy(k).to(torch.bfloat16).float().numpy()
    soft, logits = run(q, k, tl, ml, dev, dtype)
    tol = TOL["fp32"] if dtype == "fp32" else dict(rel=1e-3, abs=1e-4, soft=1e-3)
    worst = 0.0
    for b0 in range(0, w.batch, chunk):
        sl = slice(b0, min(b0 + chunk, w.batch))
        rs, rl, parts = oll.loglik(q[sl], k[sl], tl[sl], ml[sl], return_parts=True)
        amb = oll.threshold_ambiguous(parts["prior_raw"])
        e, _ = compare(soft[sl], logits[sl], rs, rl, amb, tol, f"{name}/{dtype} utterances {b0}..")
        worst = max(worst, e)
    # size-independent structure: soft mass only inside each window, valid rows sum to one
    rows = np.arange(w.t1max)[None, :] < ml[:, None]
    assert np.allclose(soft.sum(2)[rows], 1.0, atol=2e-4)
    assert soft[~rows].sum() == 0.0
    return worst


def test_full_size_cfg3_bf16_against_oracle(cuda_device):
    """BASELINE.json configs[2], the bench workload itself: batch 256 ragged x 1000 x 200, bf16 operands."""
    _full_size_against_oracle(cuda_device, "cfg3", "bf16", 16)


def test_full_size_cfg4_against_oracle(cuda_device):
    """BASELINE.json configs[3]: 512 tokens x 4096 frames, batch 16 (both TMEM accumulator chunks, 32 frame tiles per utterance)."""
    _full_size_against_oracle(cuda_device, "cfg4", "bf16", 1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B", [5, 1300])
def test_unpack_operands_packed_upload(cuda_device, dtype, B):
    """isp_unpack_operands: packed valid rows (one plain DMA each) -> padded operands, bit-exact rows, zero padding; the batch of
    1300 exercises the chunked scan of the offsets."""
    from isp_tts_b200.alignment import pack_rows, unpack_operands
    T1, T2, D = (300, 70, 128) if B < 100 else (40, 12, 16)
    tl, ml = synth.lengths(B, T2, T1, True, 91)
    q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, 92)
    qh, kh = torch.from_numpy(q).to(dtype), torch.from_numpy(k).to(dtype)
    qp, kp = pack_rows(qh, ml).pin_memory(), pack_rows(kh, tl).pin_memory()
    assert qp.shape == (int(ml.sum()), D) and kp.shape == (int(tl.sum()), D)
    qd, kd = unpack_operands(qp.to(cuda_device, non_blocking=True), kp.to(cuda_device, non_blocking=True),
                             torch.from_numpy(tl), torch.from_numpy(ml), T1, T2)
    torch.cuda.synchronize()
    assert torch.equal(qd.cpu(), qh) and torch.equal(kd.cpu(), kh)          # synth's padding rows are zeros
