"""Debug aid: per-utterance start / end of the MAS kernel inside isp_align_forward (globaltimer), relative to the first start."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isp_tts_b200 import _lib, synth
from isp_tts_b200.alignment import _align_cuda

w = synth.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
dev = torch.device("cuda:0")
tl, ml = synth.workload_lengths(w, None)
B, T1, T2, D = len(tl), w.t1max, w.t2max, w.dim
q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, 5)
qd, kd = torch.from_numpy(q).to(dev).bfloat16(), torch.from_numpy(k).to(dev).bfloat16()
tlt, mlt = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
lib = _lib.load()
_lib.set_option("mas.dbg", 64)
for kv in sys.argv[2:]:
    a, b = kv.split("="); _lib.set_option(a, int(b))
logits = torch.empty((B, T1, T2), dtype=torch.float32, device=dev); soft = torch.empty_like(logits)
hard = torch.empty((B, T1, T2), dtype=torch.int16, device=dev); dur = torch.empty((B, T2), dtype=torch.int64, device=dev)
nb = lib.isp_align_workspace_bytes(B, T1, T2, D, 1)
ws = torch.zeros(nb, dtype=torch.uint8, device=dev)
_lib.set_option("align.trace", 1)
roff = ((lib.isp_mas_workspace_bytes(B, T1, T2) + 255) & ~255) + 4 * ((B + 1) & ~1)
st = torch.cuda.current_stream().cuda_stream
for it in range(4):
    ws[roff + 8:roff + 16] = 255          # first-start stamp: atomicMin
    ws[roff + 16:roff + 24] = 0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rc = lib.isp_align_forward(qd.data_ptr(), kd.data_ptr(), 1, tlt.data_ptr(), mlt.data_ptr(), B, T1, T2, D, D ** -0.5, 1, logits.data_ptr(),
                               soft.data_ptr(), hard.data_ptr(), dur.data_ptr(), None, None, ws.data_ptr(), nb, 1, st)
    assert rc == 0, lib.isp_last_error()
    e.record(); torch.cuda.synchronize()
print(f"{w.name}: align call {s.elapsed_time(e)*1e3:.1f} us")
stamps = ws[roff:roff + 24].view(torch.int64).cpu().numpy()
off = 256 + ((B * 4 + 15) & ~15); off += (B + 15) & ~15
tr = ws[off:off + 32 * B].view(torch.int64).cpu().numpy().reshape(B, 4)
t0 = tr[:, 0].min()
order = np.argsort(-(ml * 1024 + tl), kind="stable")
for r, b in enumerate(order):
    if r % 6 == 0 or r < 8:
        print(f"  rank {r:3d}: {int(ml[b]):4d} x {int(tl[b]):3d}  b={b:3d}  start {(tr[b,0]-t0)/1e3:6.1f} -> end {(tr[b,1]-t0)/1e3:6.1f}  ({(tr[b,1]-tr[b,0])/1e3:5.1f})  cta {int(tr[b,2])>>8} slot {(int(tr[b,2])>>4)&15} stages {int(tr[b,3])}")
late = np.argsort(-tr[:, 1])[:6]
rank_of = np.empty(B, np.int64); rank_of[order] = np.arange(B)
for b in late:
    print(f"  late: rank {rank_of[b]:3d} {int(ml[b]):4d} x {int(tl[b]):3d} start {(tr[b,0]-t0)/1e3:6.1f} -> end {(tr[b,1]-t0)/1e3:6.1f} cta {int(tr[b,2])>>8} slot {(int(tr[b,2])>>4)&15} in-turn {int(tr[b,2])&1} stages {int(tr[b,3])}")
print(f"  log-likelihood kernel: first CTA start {(stamps[1]-t0)/1e3:.1f}, last CTA end {(stamps[2]-t0)/1e3:.1f}; MAS origin {(stamps[0]-t0)/1e3:.1f} (relative to the first MAS sweep start)")
print(f"  first start 0, last start {(tr[:,0].max()-t0)/1e3:.1f}, last end {(tr[:,1].max()-t0)/1e3:.1f}")
