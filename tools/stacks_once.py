"""One fused pass of the projection stacks at the cfg3 recipe shape after a warm-up: the command ncu lists launches of."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isp_tts_b200 import Aligner, synth
dev = torch.device("cuda:0")
al = Aligner(**synth.RECIPE_HP).eval().to(dev)
al.attention.gemm_dtype = "bf16"
B, T1, T2 = 256, 1000, 200
tl, ml = synth.lengths(B, T2, T1, True, 1236)
mel, txt = synth.recipe_inputs(5, B, T1, T2, tl, ml)
args = [torch.from_numpy(a).to(dev) for a in (mel, txt, ml, tl)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
with torch.no_grad():
    for _ in range(n):
        q, k = al.attention.encode(*args)
torch.cuda.synchronize()
print("ok", tuple(q.shape), tuple(k.shape), q.dtype)
