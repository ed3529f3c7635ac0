"""Timing of the key convolution (384 -> 768, k = 5, GELU + norm sums) at the cfg3 shape for ring depths / tile widths."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isp_tts_b200 import synth
from isp_tts_b200.gemm import conv1d_channels_last
dev = torch.device("cuda:0")
B, T2 = 256, 200
tl, ml = synth.lengths(B, T2, 1000, True, 1237)
tld = torch.from_numpy(tl).to(dev)
x = torch.randn((B, T2, 384), device=dev).to(torch.float16)
x = x * (torch.arange(T2, device=dev)[None, :, None] < tld[:, None, None])
w = (torch.randn((5, 768, 384), device=dev) * 0.02).to(torch.float16)

def gms(fn, reps=10):
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

rows = int(((tl + 127) // 128 * 128).sum())
fl = 2 * rows * 768 * 1920
for bn in (256, 192, 128):
    for st in (2, 3, 4):
        try:
            t = gms(lambda: conv1d_channels_last(x, w, tld, act="gelu", out_dtype=torch.float16, col_stats=True, bn=bn, stages=st))
            print(f"bn={bn} stages={st}: {t:7.1f} us  {fl / t / 1e6:6.0f} TFLOP/s executed", flush=True)
        except Exception as exc:
            print(f"bn={bn} stages={st}: {str(exc)[:120]}", flush=True)
