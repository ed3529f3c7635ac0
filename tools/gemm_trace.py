"""Phase timing inside gemm_kernel (debug): SM-clock stamps per CTA for the scores shape of row f-1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isp_tts_b200.gemm import bgemm
dev = torch.device("cuda:0")
B, T1, T2, D = 256, 1000, 200, 128
q = torch.randn((B, T1, D), device=dev).to(torch.bfloat16)
k = torch.randn((B, T2, D), device=dev).to(torch.bfloat16)
ncta = 8 * B
for out_dtype in (torch.float32, torch.bfloat16):
    tr = torch.zeros((ncta, 8), dtype=torch.int64, device=dev)
    for _ in range(3):
        bgemm(q, k.transpose(1, 2), out_dtype=out_dtype, trace=tr)
    torch.cuda.synchronize()
    t = tr.cpu().numpy().astype(np.float64)
    d = np.diff(t, axis=1)
    names = ["setup (barriers, tmap prefetch, TMEM alloc, sync)", "first stage issued", "first stage landed", "all MMAs issued",
             "accumulator complete (seen by epilogue)", "epilogue warp 0 done", "dealloc + exit"]
    print(str(out_dtype), "median cycles per phase over", ncta, "CTAs; total", np.median(t[:, 7] - t[:, 0]))
    for i, n in enumerate(names):
        print(f"   {n:52s} median {np.median(d[:, i]):9.0f}   p90 {np.percentile(d[:, i], 90):9.0f}")
