"""Debug aid: MAS kernels side by side (cluster = isp_mas_cluster.cu, v1 = isp_mas.cu, wide = isp_mas_wide.cu) on full-length and ragged
batches:  python tools/masc_compare.py [B,T1,T2 ...]   (wrapper time included: ~15 us)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from isp_tts_b200 import _lib, synth
from isp_tts_b200.mas import mas_forward

dev = torch.device("cuda:0")


def t(shape, impl, ragged):
    B, T1, T2 = shape
    _lib.set_option("mas.impl", impl)
    x = torch.from_numpy(synth.noise_logits(B, T1, T2, 3)).to(dev)
    if ragged:
        tl, ml = synth.lengths(B, T2, T1, True, 9)
    else:
        tl, ml = np.full(B, T2, np.int64), np.full(B, T1, np.int64)
    tl, ml = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
    for _ in range(3):
        mas_forward(x, tl, ml)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(10):
        s.record(); mas_forward(x, tl, ml); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    return min(ts)


shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(16, 4096, 512), (64, 4096, 512), (16, 4096, 640), (16, 2000, 384), (8, 4096, 1024), (16, 2000, 700)]
kernels = [k for k in os.environ.get("KERNELS", "cluster,v1,wide").split(",")]
for shape in shapes:
    for ragged in (False, True):
        r = {}
        for name, impl in (("cluster", 4), ("v1", 1), ("wide", 3)):
            if name not in kernels or (name == "v1" and shape[2] > 640):
                continue
            try:
                r[name] = round(t(shape, impl, ragged), 1)
            except Exception:
                r[name] = "n/a"
        print(shape, "ragged" if ragged else "dense", r, flush=True)
