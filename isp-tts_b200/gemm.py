"""Batched, ragged GEMM on the sm_100a tensor cores (isp_gemm_batched, csrc/isp_gemm.cu) for torch tensors.

`bgemm(x, y)` is `x @ y` for x (batch, M, K) and y (batch, K, N).  Either argument may be a transposed VIEW (or a 2-D tensor /
an expanded one shared by the whole batch): the strides say whether the contraction index or the row / column index is
contiguous, and the kernel loads either form with TMA -- no copy is made.  There is no fallback: on anything but a B200
the call raises.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

__all__ = ["bgemm", "conv1d_channels_last", "ACT"]

ACT = {None: 0, "linear": 0, "none": 0, "relu": 1, "gelu": 2}


def _dt(t: torch.Tensor) -> int:
    return _lib.dtype_code(t.dtype)


def _operand(t: torch.Tensor, cd: int):
    """(tensor, mn_major, ld, batch_stride) of a (batch, r, c) operand whose contraction index is dimension `cd` (1 or 2).
    K-major = the contraction index is contiguous, MN-major = the other one is; anything else (or strides that are not
    whole 16 B units) is copied once into an aligned row-major buffer."""
    if t.dim() == 2:
        t = t.unsqueeze(0)
    od = 3 - cd
    esz = t.element_size()

    def layout(x):
        if x.stride(cd) == 1:
            mn, ld = 0, x.stride(od)
        elif x.stride(od) == 1:
            mn, ld = 1, x.stride(cd)
        else:
            return None
        bs = x.stride(0) if x.shape[0] > 1 else 0
        if ld < x.shape[cd if mn == 0 else od] or (ld * esz) % 16 or (bs * esz) % 16 or x.data_ptr() % 16:
            return None
        return mn, ld, bs

    lay = layout(t)
    if lay is None:
        pad = (-t.shape[2]) % (16 // esz)
        t = torch.nn.functional.pad(t, (0, pad))[:, :, :t.shape[2]] if pad else t.contiguous()
        lay = layout(t)
    return (t,) + lay


def _len_ptr(v, dev):
    if v is None:
        return None, None
    v = v.to(device=dev, dtype=torch.int64).contiguous()
    return v, v.data_ptr()


def _launch(desc: "_lib.GemmDesc", dev):
    lib = _lib.load()
    with torch.cuda.device(dev):
        rc = lib.isp_gemm_batched(ctypes.byref(desc), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "isp_gemm_batched")


def bgemm(x: torch.Tensor, y: torch.Tensor, *, out_dtype: torch.dtype = torch.float32, alpha: float = 1.0,
          m_len=None, n_len=None, k_len=None, act=None, col_stats: bool = False, bn: int = 0, out: torch.Tensor | None = None,
          trace: torch.Tensor | None = None):
    """act(alpha * x @ y): x (batch, M, K), y (batch, K, N) -> (batch, M, N) in `out_dtype` (fp32 accumulate).

    m_len / n_len / k_len: optional (batch,) lengths; rows / columns of the result past them are zeros and the contraction
    stops at k_len (one operand must be zero from there on).  col_stats=True also returns the per-slab column sums of the
    result and of its square, (batch, 4 * ceil(M / 128), N, 2) fp32."""
    dev = x.device
    _lib.require_device(dev)
    if x.dtype != y.dtype:
        raise ValueError("x and y must have the same dtype")
    xo, a_mn, lda, a_b = _operand(x, 2)
    yo, b_mn, ldb, b_b = _operand(y, 1)
    batch = max(xo.shape[0], yo.shape[0])
    M, K = xo.shape[1], xo.shape[2]
    if yo.shape[1] != K:
        raise ValueError(f"shape mismatch: {tuple(x.shape)} @ {tuple(y.shape)}")
    N = yo.shape[2]
    esz_c = 4 if out_dtype == torch.float32 else 2
    ldc = (N * esz_c + 15) // 16 * 16 // esz_c
    if out is None:
        buf = torch.empty((batch, M, ldc), dtype=out_dtype, device=dev)
    else:
        if out.dtype != out_dtype or tuple(out.shape) != (batch, M, N) or out.stride(2) != 1 or out.stride(1) < N:
            raise ValueError("out must be (batch, M, N) in out_dtype with a unit last stride")
        buf, ldc = out, out.stride(1)
    stats = torch.zeros((batch, 4 * ((M + 127) // 128), N, 2), dtype=torch.float32, device=dev) if col_stats else None
    keep = [_len_ptr(m_len, dev), _len_ptr(n_len, dev), _len_ptr(k_len, dev)]
    d = _lib.GemmDesc()
    d.a, d.b, d.c = xo.data_ptr(), yo.data_ptr(), buf.data_ptr()
    d.m_len, d.n_len, d.k_len = keep[0][1], keep[1][1], keep[2][1]
    d.col_stats = stats.data_ptr() if stats is not None else None
    d.lda, d.ldb, d.ldc = lda, ldb, ldc
    d.a_batch, d.b_batch, d.c_batch = a_b, b_b, (buf.stride(0) if batch > 1 else M * ldc)
    d.b_tap_stride = 0
    d.batch, d.M, d.N, d.K = batch, M, N, K
    d.dtype_ab, d.dtype_c = _dt(xo), _lib.dtype_code(out_dtype)
    d.a_mn_major, d.b_mn_major = a_mn, b_mn
    d.taps, d.tap_shift, d.act, d.bn, d.skip_padding = 1, 0, ACT[act], bn, 0
    d.alpha = float(alpha)
    d.trace = trace.data_ptr() if trace is not None else None
    _launch(d, dev)
    res = buf if out is not None else (buf[:, :, :N] if ldc != N else buf)
    return (res, stats) if col_stats else res


def conv1d_channels_last(x: torch.Tensor, w_taps: torch.Tensor, lengths=None, *, act=None, out_dtype: torch.dtype = torch.bfloat16,
                         col_stats: bool = False, bn: int = 0, stages: int = 0):
    """'same' Conv1d without bias on a channels-last activation, as an implicit GEMM (alignment.py:58-62):
    x (B, T, Cin), w_taps (k, Cout, Cin) = conv.weight.permute(2, 0, 1), rows of x past `lengths` must be zero ->
    act(conv) (B, T, Cout), rows past `lengths` zero; with col_stats the masked column sums for the instance norm (slabs of
    128-row tiles that lie wholly past `lengths` are left undefined: isp_instance_norm_apply does not read them)."""
    dev = x.device
    _lib.require_device(dev)
    if x.dtype != w_taps.dtype:
        raise ValueError("x and w_taps must have the same dtype")
    esz = x.element_size()
    if x.stride(2) != 1 or (x.stride(1) * esz) % 16 or (x.stride(0) * esz) % 16 or x.data_ptr() % 16:
        x = x.contiguous()
        if (x.shape[2] * esz) % 16:
            raise ValueError("Cin * element size must be a multiple of 16 B")
    w_taps = w_taps.contiguous()
    k, Cout, Cin = w_taps.shape
    if Cin != x.shape[2] or k % 2 != 1:
        raise ValueError("w_taps must be (odd k, Cout, Cin)")
    if (Cin * esz) % 16:
        raise ValueError("Cin * element size must be a multiple of 16 B")
    B, T, _ = x.shape
    esz_c = 4 if out_dtype == torch.float32 else 2
    ldc = (Cout * esz_c + 15) // 16 * 16 // esz_c
    buf = torch.empty((B, T, ldc), dtype=out_dtype, device=dev)
    # not zeroed: only the slabs of tiles below ceil(len / 128) are written, and isp_instance_norm_apply reads exactly those
    stats = torch.empty((B, 4 * ((T + 127) // 128), Cout, 2), dtype=torch.float32, device=dev) if col_stats else None
    lens, lens_ptr = _len_ptr(lengths, dev)
    d = _lib.GemmDesc()
    d.a, d.b, d.c = x.data_ptr(), w_taps.data_ptr(), buf.data_ptr()
    d.m_len, d.n_len, d.k_len = lens_ptr, None, None
    d.col_stats = stats.data_ptr() if stats is not None else None
    d.lda, d.ldb, d.ldc = x.stride(1), Cin, ldc
    d.a_batch, d.b_batch, d.c_batch = x.stride(0), 0, T * ldc
    d.b_tap_stride = Cout * Cin
    d.batch, d.M, d.N, d.K = B, T, Cout, Cin
    d.dtype_ab, d.dtype_c = _dt(x), _lib.dtype_code(out_dtype)
    d.a_mn_major, d.b_mn_major = 0, 0
    d.taps, d.tap_shift, d.act, d.bn, d.skip_padding = k, -(k // 2), ACT[act], bn, 0
    d.alpha = 1.0
    d.stages = stages
    _launch(d, dev)
    res = buf[:, :, :Cout] if ldc != Cout else buf
    return (res, stats) if col_stats else res
