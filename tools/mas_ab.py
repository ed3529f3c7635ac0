"""Debug aid: time isp_mas_forward of a given build of the library (A/B between builds on one box).
    python tools/mas_ab.py path/to/lib.so cfg3 cfg2"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isp_tts_b200 import synth

lib = ctypes.CDLL(os.path.abspath(sys.argv[1]))
vp, ci, c64, csz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t
lib.isp_mas_workspace_bytes.argtypes = [ci, ci, ci]; lib.isp_mas_workspace_bytes.restype = csz
lib.isp_mas_forward.argtypes = [vp, c64, c64, c64, vp, vp, ci, ci, ci, vp, vp, vp, csz, vp]
dev = torch.device("cuda:0")
for name in sys.argv[2:]:
    w = synth.WORKLOADS[name]
    tl, ml = synth.workload_lengths(w, None)
    B = len(tl)
    x = torch.from_numpy(synth.noise_logits(B, w.t1max, w.t2max, w.seed)).to(dev)
    tlt, mlt = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
    _, T1, T2 = x.shape
    hard = torch.empty((B, T1, T2), dtype=torch.int16, device=dev); dur = torch.empty((B, T2), dtype=torch.int64, device=dev)
    wsb = lib.isp_mas_workspace_bytes(B, T1, T2); ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for it in range(6):
        reps = 1 if it < 2 else 20
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            assert lib.isp_mas_forward(x.data_ptr(), x.stride(0), x.stride(1), 1, tlt.data_ptr(), mlt.data_ptr(), B, T1, T2, hard.data_ptr(), dur.data_ptr(), ws.data_ptr(), wsb, st) == 0
        e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3 / reps)
    print(f"{os.path.basename(sys.argv[1])} {name}: {min(ts[2:]):.1f} us (median {np.median(ts[2:]):.1f})", flush=True)
