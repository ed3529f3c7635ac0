"""isp_align_forward: the log-likelihood kernel and the MAS kernel linked through per-utterance ready counts (the MAS kernel starts
under the first kernel's last wave).  Results must be those of the two separate calls bit for bit, and the oracle's."""
import numpy as np
import pytest
import torch

from isp_tts_b200 import _lib, synth
from isp_tts_b200.alignment import _ALIGN_WS, _align_cuda, _loglik_cuda, align_forward, loglik_forward
from isp_tts_b200.mas import mas_forward
from oracle import mas as omas

pytestmark = pytest.mark.gpu


def _inputs(dev, B, T1, T2, D, seed, dtype=torch.bfloat16, ragged=True):
    tl, ml = synth.lengths(B, T2, T1, ragged, seed)
    q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, seed + 10)
    return (torch.from_numpy(q).to(dev).to(dtype), torch.from_numpy(k).to(dev).to(dtype),
            torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev), tl, ml)


def _separate(q, k, tlt, mlt, scale):
    soft, logits = _loglik_cuda(q, k, tlt, mlt, scale, True)
    hard, dur, path = mas_forward(logits, tlt, mlt, return_path=True)
    return soft, logits, hard, dur, path


# (B, T1, T2, D): one SM's worth, cfg3, more utterances than SMs x 2 (CTAs in waves), beyond the self-ranking range (plan kernel:
# plain sequence), wider than the second MAS kernel (round-1 kernel: plain sequence), one utterance
SHAPES = [(4, 300, 64, 128), (256, 1000, 200, 128), (400, 500, 120, 64), (600, 300, 80, 64), (3, 1500, 400, 128), (1, 130, 9, 32)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_linked_equals_separate_calls(cuda_device, shape, dtype):
    B, T1, T2, D = shape
    if dtype == torch.float32 and B * T1 * T2 > 30e6:
        pytest.skip("one operand type is enough at the large shapes")
    q, k, tlt, mlt, tl, ml = _inputs(cuda_device, B, T1, T2, D, 7 + B, dtype)
    scale = D ** -0.5
    ref = _separate(q, k, tlt, mlt, scale)
    _ALIGN_WS.clear()
    for it in range(3):                                   # first call: fresh workspace (memset in front); then the cached, clean one
        out = _align_cuda(q, k, tlt, mlt, scale, True, return_path=True)
        torch.cuda.synchronize()
        for name, a, b in zip(("soft", "logits", "hard", "durations", "path"), ref, out):
            assert torch.equal(a, b), f"{name} differs on call {it}"
    rh, rd = omas.b_mas_with_durations(out[1].cpu().numpy(), tl, ml)
    assert np.array_equal(out[2].cpu().numpy(), rh) and np.array_equal(out[3].cpu().numpy(), rd)


def test_linked_back_to_back_different_batches(cuda_device):
    """The same cached workspace serves different batches of one shape one after the other without a host sync in between:
    the ready counts a call leaves behind must be clean before the next call's first kernel increments them."""
    B, T1, T2, D = 64, 700, 150, 128
    sets = [_inputs(cuda_device, B, T1, T2, D, 40 + i) for i in range(4)]
    refs = [_separate(s[0], s[1], s[2], s[3], D ** -0.5) for s in sets]
    _ALIGN_WS.clear()
    outs = [_align_cuda(s[0], s[1], s[2], s[3], D ** -0.5, True, return_path=True) for s in sets for _ in range(2)]
    torch.cuda.synchronize()
    for i, out in enumerate(outs):
        for a, b in zip(refs[i // 2], out):
            assert torch.equal(a, b)
    assert len(_ALIGN_WS) == 1


def test_dirty_workspace_without_the_clean_flag(cuda_device):
    """C ABI: a workspace full of garbage and flags = 0 -- the call clears its counters itself."""
    B, T1, T2, D = 20, 400, 90, 64
    q, k, tlt, mlt, tl, ml = _inputs(cuda_device, B, T1, T2, D, 3)
    ref = _separate(q, k, tlt, mlt, D ** -0.5)
    lib = _lib.load()
    nb = lib.isp_align_workspace_bytes(B, T1, T2, D, _lib.ISP_DTYPE_BF16)
    ws = torch.full((nb,), 0x5A, dtype=torch.uint8, device=cuda_device)
    logits = torch.empty((B, T1, T2), dtype=torch.float32, device=cuda_device)
    soft = torch.empty_like(logits)
    hard = torch.empty((B, T1, T2), dtype=torch.int16, device=cuda_device)
    dur = torch.empty((B, T2), dtype=torch.int64, device=cuda_device)
    rowsum = torch.empty((B, T1), dtype=torch.float32, device=cuda_device)
    ref_rowsum = _loglik_cuda(q, k, tlt, mlt, D ** -0.5, True, want_rowsum=True)[2]
    valid = torch.arange(T1, device=cuda_device)[None] < mlt[:, None]
    st = torch.cuda.current_stream().cuda_stream
    for flags in (0, _lib.ISP_ALIGN_WS_CLEAN, _lib.ISP_ALIGN_WS_CLEAN):
        rc = lib.isp_align_forward(q.data_ptr(), k.data_ptr(), _lib.ISP_DTYPE_BF16, tlt.data_ptr(), mlt.data_ptr(), B, T1, T2, D, D ** -0.5, 1,
                                   logits.data_ptr(), soft.data_ptr(), hard.data_ptr(), dur.data_ptr(), None, rowsum.data_ptr(), ws.data_ptr(), nb, flags, st)
        assert rc == 0, lib.isp_last_error()
        torch.cuda.synchronize()
        assert torch.equal(logits, ref[1]) and torch.equal(soft, ref[0]) and torch.equal(hard, ref[2]) and torch.equal(dur, ref[3])
        assert torch.equal(rowsum[valid], ref_rowsum[valid])          # the prior's row sums for the backward pass (valid frames)
        assert lib.isp_mas_status(ws.data_ptr(), st) == 0
    assert lib.isp_align_forward(q.data_ptr(), k.data_ptr(), _lib.ISP_DTYPE_BF16, tlt.data_ptr(), mlt.data_ptr(), B, T1, T2, D, D ** -0.5, 1,
                                 logits.data_ptr(), soft.data_ptr(), hard.data_ptr(), dur.data_ptr(), None, None, ws.data_ptr(), nb - 1, 0, st) != 0


def test_linked_inside_a_cuda_graph(cuda_device):
    B, T1, T2, D = 48, 600, 128, 128
    q, k, tlt, mlt, tl, ml = _inputs(cuda_device, B, T1, T2, D, 17)
    ref = _separate(q, k, tlt, mlt, D ** -0.5)
    cap = torch.cuda.Stream(device=cuda_device)
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        _align_cuda(q, k, tlt, mlt, D ** -0.5, True)       # this stream's workspace exists (and is clean) before the capture
    cap.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=cap):
        out = _align_cuda(q, k, tlt, mlt, D ** -0.5, True, return_path=True)
    for _ in range(3):
        for t in out:
            t.zero_()
        g.replay()
        torch.cuda.synchronize()
        for a, b in zip(ref, out):
            assert torch.equal(a, b)


def test_align_forward_autograd_matches_loglik_forward(cuda_device):
    B, T1, T2, D = 6, 260, 70, 64
    q, k, tlt, mlt, tl, ml = _inputs(cuda_device, B, T1, T2, D, 23, torch.float32)
    g_soft = torch.randn(B, T1, T2, device=cuda_device)
    g_logits = torch.randn(B, T1, T2, device=cuda_device)
    grads = []
    for fn in (loglik_forward, align_forward):
        qq, kk = q.clone().requires_grad_(True), k.clone().requires_grad_(True)
        out = fn(qq, kk, tlt, mlt)
        (out[0] * g_soft).sum().add((out[1] * g_logits).sum()).backward()
        grads.append((qq.grad, kk.grad, out))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
    hard, dur = grads[1][2][2], grads[1][2][3]
    assert not hard.requires_grad and not dur.requires_grad
    rh, rd = omas.b_mas_with_durations(grads[1][2][1].detach().cpu().numpy(), tl, ml)
    assert np.array_equal(hard.cpu().numpy(), rh) and np.array_equal(dur.cpu().numpy(), rd)
