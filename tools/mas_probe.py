"""Debug aid: phase timing of the MAS kernel (SM clocks of utterance 0) for a workload.

    python tools/mas_probe.py cfg2 cfg3 cfg3d cfg4        # env: PROBE_RING=0,64  PROBE_SLOTS=0,1,2
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from isp_tts_b200 import _lib, synth


def run(name, ring=0, slots=0, dbg=0, bits_global=0):
    w = synth.WORKLOADS[name]
    lib = _lib.load()
    _lib.set_option("mas.ring_rows", ring)
    _lib.set_option("mas.slots", slots)
    _lib.set_option("mas.dbg", dbg)
    _lib.set_option("mas.bits_global", bits_global)
    dev = torch.device("cuda:0")
    tl, ml = synth.workload_lengths(w)
    # utterance 0 is the one that is probed: make it the longest so that the chain is visible
    k = int(np.argmax(ml * 10000 + tl))
    tl[[0, k]] = tl[[k, 0]]
    ml[[0, k]] = ml[[k, 0]]
    x = torch.from_numpy(synth.noise_logits(w.batch, w.t1max, w.t2max, w.seed)).to(dev)
    tlt, mlt = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
    B, T1, T2 = x.shape
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
    pr = ws[64:192].view(torch.int64).cpu().numpy()
    n = int(ml[0])
    print(f"{name} bits_global={bits_global} ring={ring} slots={slots} dbg={dbg}: kernel {min(times):.1f} us (median {np.median(times):.1f}); utterance 0 ({n} x {int(tl[0])}): "
          f"forward {pr[1]-pr[0]} cyc ({(pr[1]-pr[0])/(n+31):.1f}/step), barrier {pr[2]-pr[1]}, "
          f"backtrack {pr[3]-pr[2]} cyc ({(pr[3]-pr[2])/n:.1f}/row); waiting for logits {pr[4]} cyc, for the neighbour strip {pr[5]} cyc; loader: {pr[7]} cyc of which waiting for free stages {pr[6]}; steady chunks: {pr[9]} x {pr[8]/max(pr[9],1):.0f} cyc; backtrack waiting for converters {pr[10]} cyc, per block: windows+prefetch {pr[11]/((n+31)//32):.0f} + outputs {pr[13]/((n+31)//32):.0f} + chain {pr[12]/((n+31)//32):.0f} cyc, waiting for the zero-fill + writing the ones {pr[3]-pr[14]} cyc", flush=True)


rings = [int(c) for c in os.environ.get("PROBE_RING", "0").split(",")]
slots = [int(c) for c in os.environ.get("PROBE_SLOTS", "0").split(",")]
for name in sys.argv[1:] or ["cfg2", "cfg3"]:
    for r in rings:
        for sl in slots:
            for d in [int(c) for c in os.environ.get("PROBE_DBG", "0").split(",")]:
                for c in [int(c) for c in os.environ.get("PROBE_BITS_GLOBAL", "0").split(",")]:
                    run(name, r, sl, d, c)
