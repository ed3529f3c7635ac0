import sys; sys.path.insert(0, ".")
import numpy as np, torch
from isp_tts_b200 import synth
from isp_tts_b200.mas import mas_forward
dev = torch.device("cuda:0")
w = synth.WORKLOADS["cfg4"]
tl, ml = synth.workload_lengths(w)
x = torch.from_numpy(synth.noise_logits(w.batch, w.t1max, w.t2max, w.seed)).to(dev)
tl, ml = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
for _ in range(4):
    mas_forward(x, tl, ml)
torch.cuda.synchronize()
