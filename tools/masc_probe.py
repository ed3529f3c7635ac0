"""Debug aid: phase timing of the cluster MAS kernel (isp_mas_cluster.cu, option masc.trace) for a workload.

    python tools/masc_probe.py cfg4
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from isp_tts_b200 import _lib, synth


def run(name):
    lib = _lib.load()
    _lib.set_option("mas.impl", 4)
    _lib.set_option("masc.trace", 1)
    _lib.set_option("masc.dbg", int(os.environ.get("MASC_DBG", "0")))
    dev = torch.device("cuda:0")
    if "," in name:                                  # "B,T1,T2": full-length utterances of that shape
        B, T1, T2 = (int(v) for v in name.split(","))
        tl, ml = np.full(B, T2, np.int64), np.full(B, T1, np.int64)
        x = torch.from_numpy(synth.noise_logits(B, T1, T2, 5)).to(dev)
    else:
        w = synth.WORKLOADS[name]
        tl, ml = synth.workload_lengths(w)
        x = torch.from_numpy(synth.noise_logits(w.batch, w.t1max, w.t2max, w.seed)).to(dev)
    tlt, mlt = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
    B, T1, T2 = x.shape
    nc = (T2 + 127) // 128
    hard = torch.empty((B, T1, T2), dtype=torch.int16, device=dev)
    dur = torch.empty((B, T2), dtype=torch.int64, device=dev)
    wsb = lib.isp_mas_workspace_bytes(B, T1, T2)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    times = []
    for it in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = lib.isp_mas_forward(x.data_ptr(), x.stride(0), x.stride(1), 1, tlt.data_ptr(), mlt.data_ptr(), B, T1, T2,
                                 hard.data_ptr(), dur.data_ptr(), ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream)
        e.record()
        torch.cuda.synchronize()
        assert rc == 0, lib.isp_last_error()
        times.append(s.elapsed_time(e) * 1e3)
    off = 256 + ((B * T1 * 2 + 15) & ~15)
    tr = ws[off:off + B * nc * 32 * 8].view(torch.int64).cpu().numpy().reshape(B, nc, 32)
    t0 = tr[:, :, 0].min()
    print(f"{name}: kernel {min(times):.1f} us (median {np.median(times):.1f}), {B} clusters of {nc}")
    k = int(np.argmax(ml * 10000 + tl))
    for b in sorted({0, k}):
        print(f" utterance {b}: {ml[b]} frames x {tl[b]} tokens")
        for c in range(nc):
            r = tr[b, c]
            us = lambda i: (r[i] - t0) / 1e3 if r[i] else float('nan')
            print(f"  cta {c}: start {us(0):6.1f} setup {us(1):6.1f} | strips end {us(2):6.1f} {us(3):6.1f} loader {us(9):6.1f} fill {us(14):6.1f} | maps {us(4):6.1f} hops {us(5):6.1f} | sync {us(6):6.1f} walk {us(7):6.1f} end {us(8):6.1f} | waits (kcyc) logits {r[10]/1e3:.1f} {r[12]/1e3:.1f} neighbour {r[11]/1e3:.1f} {r[13]/1e3:.1f}")


for name in sys.argv[1:] or ["cfg4"]:
    run(name)
