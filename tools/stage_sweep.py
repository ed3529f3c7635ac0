import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from isp_tts_b200 import _lib, synth
from isp_tts_b200.alignment import stage_operands
w = synth.WORKLOADS["cfg3"]; dev = torch.device("cuda:0")
tl, ml = synth.workload_lengths(w)
q, k = synth.encoded_pair(w.batch, w.t1max, w.t2max, w.dim, tl, ml, w.seed)
qh = torch.from_numpy(q).bfloat16().pin_memory(); kh = torch.from_numpy(k).bfloat16().pin_memory()
tlt, mlt = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
qd = torch.empty(qh.shape, dtype=qh.dtype, device=dev); kd = torch.empty(kh.shape, dtype=kh.dtype, device=dev)
valid = int((tl.sum() + ml.sum()) * w.dim * 2)
def t(f):
    for _ in range(3): f()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(10): f()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / 10
a = t(lambda: (qd.copy_(qh, non_blocking=True), kd.copy_(kh, non_blocking=True)))
print("padded DMA: %.3f ms, %.1f GB/s" % (a, (qh.numel() + kh.numel()) * 2 / a / 1e6))
for ctas in [8, 16, 32, 64, 128, 296]:
    _lib.set_option("stage.ctas", ctas)
    a = t(lambda: stage_operands(qh, kh, tlt, mlt, out_q=qd, out_k=kd))
    print("staged ctas=%d: %.3f ms, %.1f GB/s over PCIe" % (ctas, a, valid / a / 1e6))
