"""Timing of isp_gemm_batched on the cfg3 shapes of rows f-1 / f-3 (CUDA-graph replay), with and without ragged lengths."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isp_tts_b200 import synth
from isp_tts_b200.gemm import bgemm
dev = torch.device("cuda:0")
B, T1, T2, D, C = 256, 1000, 200, 128, 384
tl, ml = synth.lengths(B, T2, T1, True, 1237)
tld, mld = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
q = torch.randn((B, T1, D), device=dev).to(torch.bfloat16)
k = torch.randn((B, T2, D), device=dev).to(torch.bfloat16)
ds = torch.randn((B, T1, T2), device=dev).to(torch.bfloat16)
soft = torch.rand((B, T1, T2), device=dev)
x = torch.randn((B, T2, C), device=dev)

def gms(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        g.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3

cases = {
    "scores ragged (fills)": lambda: bgemm(q, k.transpose(1, 2), m_len=mld, n_len=tld),
    "scores all tiles": lambda: bgemm(q, k.transpose(1, 2)),
    "scores bf16 out all tiles": lambda: bgemm(q, k.transpose(1, 2), out_dtype=torch.bfloat16),
    "dQ (k_len)": lambda: bgemm(ds, k, out_dtype=torch.bfloat16, k_len=tld),
    "dK (k_len)": lambda: bgemm(ds.transpose(1, 2), q, out_dtype=torch.bfloat16, k_len=mld),
    "soft expand ragged": lambda: bgemm(soft, x, m_len=mld, k_len=tld),
    "soft expand all": lambda: bgemm(soft, x),
    "torch scores": lambda: torch.bmm(q, k.transpose(1, 2), out_dtype=torch.float32),
    "torch dQ": lambda: torch.matmul(ds, k),
    "torch dK": lambda: torch.matmul(ds.transpose(1, 2), q),
    "torch soft expand": lambda: torch.matmul(soft, x),
}
for bn in ([0] + [int(a) for a in sys.argv[1:]]):
    for name, fn in cases.items():
        if bn and name.startswith("torch"):
            continue
        if bn:
            import functools
            base = fn
            # rebuild the call with a forced tile width
            src = {"scores ragged (fills)": lambda: bgemm(q, k.transpose(1, 2), m_len=mld, n_len=tld, bn=bn),
                   "scores all tiles": lambda: bgemm(q, k.transpose(1, 2), bn=bn),
                   "soft expand ragged": lambda: bgemm(soft, x, m_len=mld, k_len=tld, bn=bn),
                   "soft expand all": lambda: bgemm(soft, x, bn=bn)}
            if name not in src:
                continue
            fn = src[name]
        try:
            print(f"bn={bn:3d} {name:28s} {gms(fn):8.1f} us", flush=True)
        except Exception as exc:
            print(f"bn={bn:3d} {name:28s} failed: {str(exc)[:100]}", flush=True)
