"""Prints the error of isp_gemm_batched against float64 for every operand layout (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isp_tts_b200.gemm import bgemm
dev = torch.device("cuda:0")
torch.manual_seed(0)
for dtype in (torch.bfloat16, torch.float32):
    for (B, M, N, K) in [(1, 128, 128, 32), (1, 128, 128, 64), (2, 200, 128, 1000), (2, 130, 72, 40)]:
        for ta in (False, True):
            for tb in (False, True):
                a = torch.randn((B, K, M) if ta else (B, M, K)).to(dtype).to(dev)
                b = torch.randn((B, N, K) if tb else (B, K, N)).to(dtype).to(dev)
                x = a.transpose(1, 2) if ta else a
                y = b.transpose(1, 2) if tb else b
                try:
                    got = bgemm(x, y)
                    torch.cuda.synchronize()
                except Exception as e:
                    print(dtype, (B, M, N, K), ta, tb, "EXC", str(e)[:200]); continue
                ref = torch.matmul(x.double(), y.double())
                err = (got.double() - ref).abs().max().item() / ref.abs().max().item()
                nz = (got != 0).float().mean().item()
                print(str(dtype)[6:], (B, M, N, K), "A", "MN" if ta else "K ", "B", "K " if tb else "MN", f"err {err:.2e} nonzero {nz:.2f}",
                      "got", [round(v, 2) for v in got[0, 0, :4].tolist()], "ref", [round(v, 2) for v in ref[0, 0, :4].tolist()], flush=True)
