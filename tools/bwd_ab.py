"""Debug aid: time isp_loglik_backward_from_logits (and the scores-route kernel) at a workload's shapes.  ISP_TTS_B200_LIB selects the build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isp_tts_b200 import synth
from isp_tts_b200.alignment import _loglik_cuda, loglik_backward_from_logits

dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["cfg3"]:
    w = synth.WORKLOADS[name]
    tl, ml = synth.workload_lengths(w, None)
    B, T1, T2, D = len(tl), w.t1max, w.t2max, w.dim
    q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, 5)
    qd, kd = torch.from_numpy(q).to(dev).bfloat16(), torch.from_numpy(k).to(dev).bfloat16()
    tlt, mlt = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
    soft, logits, rowsum = _loglik_cuda(qd, kd, tlt, mlt, D ** -0.5, True, want_rowsum=True)
    gl, gs = torch.randn_like(logits), torch.randn_like(logits)
    for label, a, b in (("both", gl, gs), ("logits only", gl, None), ("soft only", None, gs)):
        ts = []
        for it in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                ds = loglik_backward_from_logits(logits, a, b, rowsum, tlt, mlt, D ** -0.5, True, out_dtype=torch.bfloat16)
            e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 100)
        print(f"{os.environ.get('ISP_TTS_B200_LIB', 'default')} {name} {label}: {min(ts):.1f} us", flush=True)
