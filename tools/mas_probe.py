"""Debug aid: phase timing of the MAS kernel for one full-size utterance (SM clocks)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isp_tts_b200 import _lib, synth

def run(name, cols=0, ring=0, prod=0, dbg=0):
    w = synth.WORKLOADS[name]
    lib = _lib.load()
    _lib.set_option("mas.cols_per_lane", cols); _lib.set_option("mas.ring_rows", ring); _lib.set_option("mas.producer", prod); _lib.set_option("mas.dbg", dbg)
    dev = torch.device("cuda:0")
    tl, ml = synth.workload_lengths(w)
    x = torch.from_numpy(synth.noise_logits(w.batch, w.t1max, w.t2max, w.seed)).to(dev)
    tlt, mlt = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
    B, T1, T2 = x.shape
    hard = torch.empty((B, T1, T2), dtype=torch.int16, device=dev)
    dur = torch.empty((B, T2), dtype=torch.int64, device=dev)
    wsb = lib.isp_mas_workspace_bytes(B, T1, T2)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    for it in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = lib.isp_mas_forward(x.data_ptr(), x.stride(0), x.stride(1), 1, tlt.data_ptr(), mlt.data_ptr(), B, T1, T2,
                                 hard.data_ptr(), dur.data_ptr(), ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream)
        e.record(); torch.cuda.synchronize()
        assert rc == 0
    pr = ws[64:96].view(torch.int64).cpu().numpy()
    print(f"{name} cols={cols} ring={ring} prod={prod} dbg={dbg}: kernel {s.elapsed_time(e)*1e3:.1f} us; forward {pr[1]-pr[0]} cyc "
          f"({(pr[1]-pr[0])/T1:.1f}/row), barrier {pr[2]-pr[1]}, backtrack {pr[3]-pr[2]} cyc ({(pr[3]-pr[2])/T1:.1f}/row)")

cols_list = [int(c) for c in os.environ.get("PROBE_COLS", "8,4").split(",")]
prod_list = [int(c) for c in os.environ.get("PROBE_PRODS", "3,2,1").split(",")]
for name in sys.argv[1:] or ["cfg2", "cfg3"]:
    for cols in cols_list:
        for prod in prod_list:
            for dbg in [int(d) for d in os.environ.get("PROBE_DBG", "0").split(",")]:
                run(name, cols, 64, prod, dbg)
