"""Debug aid: phase timing of isp_mas2.cu (SM clocks of the longest utterance) and kernel time for a workload.

    python tools/mas2_probe.py cfg2 cfg3 cfg3d      # env: PROBE_IMPL=0,1  PROBE_SLOTS=0,1,2

The phase counters are a BUILD option of the kernel (compiled in, they cost the shipped kernel 3 %):
    bash tools/ab_build.sh probe isp_mas2.cu -DISP_MAS2_PROBE=1 && ISP_TTS_B200_LIB=tools/_ab/lib_probe.so python tools/mas2_probe.py cfg3
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from isp_tts_b200 import _lib, synth


EXTRA = {"w64": synth.Workload("w64: 32 x (64 tok x 1000 fr) dense", 32, 64, 1000, False, 99),
         "w128": synth.Workload("w128: 32 x (128 tok x 1000 fr) dense", 32, 128, 1000, False, 98)}


def run(name, impl=0, slots=0, minpair=0, batch=None, dbg=0):
    w = synth.WORKLOADS.get(name) or EXTRA[name]
    lib = _lib.load()
    _lib.set_option("mas.impl", impl)
    _lib.set_option("mas.slots", slots)
    _lib.set_option("mas2.min_pair_stages", minpair)
    _lib.set_option("mas.dbg", dbg)
    _lib.set_option("mas2.fill_us", int(os.environ.get("PROBE_FILL_US", "0")))
    _lib.set_option("mas2.pace", int(os.environ.get("PROBE_PACE", "-1")))
    dev = torch.device("cuda:0")
    tl, ml = synth.workload_lengths(w, batch)
    B = len(tl)
    x = torch.from_numpy(synth.noise_logits(B, w.t1max, w.t2max, w.seed)).to(dev)
    tlt, mlt = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
    B, T1, T2 = x.shape
    hard = torch.empty((B, T1, T2), dtype=torch.int16, device=dev)
    dur = torch.empty((B, T2), dtype=torch.int64, device=dev)
    wsb = lib.isp_mas_workspace_bytes(B, T1, T2)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    times = []
    stream = torch.cuda.current_stream().cuda_stream
    for it in range(4):
        reps = 1 if it < 2 else 20
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):         # back to back: the host's launch work hides under the previous launch
            rc = lib.isp_mas_forward(x.data_ptr(), x.stride(0), x.stride(1), 1, tlt.data_ptr(), mlt.data_ptr(), B, T1, T2,
                                     hard.data_ptr(), dur.data_ptr(), ws.data_ptr(), wsb, stream)
            assert rc == 0, lib.isp_last_error()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e) * 1e3 / reps)
    pr = ws[64:256].view(torch.int64).cpu().numpy()
    k = int(np.argmax(ml * 10000 + tl))
    n = int(ml[k])
    valid = int((tl * ml).sum())
    alg = 4 * valid + 2 * B * T1 * T2 + 8 * B * T2 + 16 * B
    t = min(times[2:])
    msg = f"{name} B={B} impl={impl} slots={slots} minpair={minpair} dbg={dbg}: kernel {t:.1f} us (median {np.median(times[2:]):.1f}) = {alg / t / 1e3:.0f} GB/s"
    if impl != 1:
        msg += (f"; longest ({n} x {int(tl[k])}): forward {pr[1]-pr[0]} cyc ({(pr[1]-pr[0])/(n+31):.1f}/step), maps ready +{pr[2]-pr[1]}, "
                f"hops+expansion+outputs {pr[3]-pr[2]} cyc; strip 0 waited {pr[4]} cyc for logits, {pr[5]} for its neighbours; "
                f"last strip ends +{pr[10]-pr[1]}; transposer ends +{pr[6]-pr[1]} (waited {pr[7]} for strips, {pr[8]} for the mapper); mapper waited {pr[9]}, computed {pr[19]}; "
                f"plan kernel -> main kernel entry {(pr[17]-pr[18])/1e3:.1f} us, entry -> sweep starts {(pr[15]-pr[17])/1e3:.1f} us, sweep start -> done {(pr[16]-pr[15])/1e3:.1f} us "
                f"({(pr[3]-pr[0])/max(pr[16]-pr[15],1):.3f} cycles/ns); hops {pr[11]-pr[2]}, expansion {pr[12]-pr[11]}, durations + fill wait {pr[13]-pr[12]}, ones {pr[3]-pr[13]}")
    print(msg, flush=True)
    if dbg & 64:
        off = 256 + ((B * 4 + 15) & ~15)
        off += (B + 15) & ~15
        tr = ws[off:off + 32 * B].view(torch.int64).cpu().numpy().reshape(B, 4)
        t0 = tr[:, 0].min()
        order = np.argsort(-(ml * 1024 + tl), kind="stable")
        rank_of = np.empty(B, np.int64); rank_of[order] = np.arange(B)
        by_cta = {}
        for b in range(B):
            by_cta.setdefault(int(tr[b, 2]) >> 8, []).append(b)
        rows = []
        for b in order:
            mates = [x for x in by_cta[int(tr[b, 2]) >> 8] if x != b]
            mate = f"with rank {rank_of[mates[0]]:3d} ({int(ml[mates[0]])} x {int(tl[mates[0]])})" if mates else "alone"
            rows.append((int(rank_of[b]), int(ml[b]), int(tl[b]), (tr[b, 0] - t0) / 1e3, (tr[b, 1] - t0) / 1e3, int(tr[b, 2]) >> 8, (int(tr[b, 2]) >> 4) & 15, int(tr[b, 2]) & 1, int(tr[b, 3]), mate))
        print("  slowest utterances: rank, frames x tokens, start -> end us, CTA, slot, in turn, ring stages, partner")
        for r in sorted(rows, key=lambda r: -r[4])[:24]:
            print("   rank %3d: %4d x %3d  %6.1f -> %6.1f  cta %3d slot %d turn %d stages %2d  %s" % r)
        print("   last end %.1f us; ends after 40 us: %d; median end %.1f" % (max(r[4] for r in rows), sum(r[4] > 40 for r in rows), float(np.median([r[4] for r in rows]))))


for name in sys.argv[1:] or ["cfg2", "cfg3"]:
    for impl in [int(c) for c in os.environ.get("PROBE_IMPL", "0").split(",")]:
        for sl in [int(c) for c in os.environ.get("PROBE_SLOTS", "0").split(",")]:
            for mp in [int(c) for c in os.environ.get("PROBE_MINPAIR", "0").split(",")]:
                for bt in [int(c) for c in os.environ.get("PROBE_BATCH", "0").split(",")]:
                    for dbg in [int(c) for c in os.environ.get("PROBE_DBG", "0").split(",")]:
                        run(name, impl, sl, mp, bt or None, dbg)
