import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from isp_tts_b200 import synth
from isp_tts_b200.alignment import _align_cuda, _loglik_cuda
from isp_tts_b200.mas import mas_forward
dev = torch.device("cuda:0")
for (B, T1, T2, D, seed) in [(4, 300, 64, 128, 1), (256, 1000, 200, 128, 2), (32, 1000, 200, 128, 3), (600, 400, 100, 64, 4), (3, 2000, 400, 128, 5)]:
    tl, ml = synth.lengths(B, T2, T1, True, seed)
    q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, seed + 10)
    qd, kd = torch.from_numpy(q).to(dev).bfloat16(), torch.from_numpy(k).to(dev).bfloat16()
    tlt, mlt = torch.from_numpy(tl).to(dev), torch.from_numpy(ml).to(dev)
    for it in range(3):
        soft0, logits0 = _loglik_cuda(qd, kd, tlt, mlt, D ** -0.5, True)
        hard0, dur0, path0 = mas_forward(logits0, tlt, mlt, return_path=True)
        soft, logits, hard, dur, path = _align_cuda(qd, kd, tlt, mlt, D ** -0.5, True, return_path=True)
        torch.cuda.synchronize()
        ok = all(torch.equal(a, b) for a, b in [(soft0, soft), (logits0, logits), (hard0, hard), (dur0, dur), (path0, path)])
        print(B, T1, T2, D, it, "ok" if ok else "MISMATCH", flush=True)
        assert ok
