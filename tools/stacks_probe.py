"""Accuracy of the bf16 fused route of the whole Aligner against the reference's recipe golden, and timing of the stacks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isp_tts_b200 import Aligner, synth
dev = torch.device("cuda:0")
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "aligner_recipe.npz"))
seed, B, T1, T2 = int(g["seed"]), int(g["B"]), int(g["T1"]), int(g["T2"])
tl, ml = g["text_len"], g["mel_len"]
al = Aligner(**synth.RECIPE_HP).eval()
al.load_state_dict({k: torch.from_numpy(v) for k, v in synth.recipe_state(seed).items()}, strict=True)
al = al.to(dev)
mel, txt = synth.recipe_inputs(seed + 1, B, T1, T2, tl, ml)
args = [torch.from_numpy(a).to(dev) for a in (mel, txt, ml, tl)]
valid = (np.arange(T1)[None, :, None] < ml[:, None, None]) & (np.arange(T2)[None, None, :] < tl[:, None, None])
for mode, fused, inner in (("fp32", False, None), ("fp32", True, None), ("bf16", False, None), ("bf16", True, torch.bfloat16), ("bf16", True, torch.float16)):
    al.attention.gemm_dtype = mode
    al.attention.fused_stacks = fused
    if inner is not None:
        al.attention.stack_dtype = inner
    print(inner, end=" ")
    with torch.no_grad():
        out = al(*args)
    lg = out.attn_logits.cpu().numpy()
    err = np.abs(lg - g["attn_logits"])[valid]
    path = out.attn_hard.cpu().numpy().argmax(2)
    moved = sum(int((path[b, :ml[b]] != g["path"][b, :ml[b]]).sum()) for b in range(B))
    print(mode, "fused" if fused else "torch", f"logits err max {err.max():.4f} mean {err.mean():.5f} p99 {np.percentile(err, 99):.4f}; frames moved {moved} of {int(ml.sum())}", flush=True)

# timing at the cfg3 shape
B, T1, T2 = 256, 1000, 200
tl, ml = synth.lengths(B, T2, T1, True, 1236)
mel, txt = synth.recipe_inputs(5, B, T1, T2, tl, ml)
args = [torch.from_numpy(a).to(dev) for a in (mel, txt, ml, tl)]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for mode, fused, inner in (("bf16", True, torch.float16), ("bf16", True, torch.bfloat16), ("bf16", False, None), ("fp32", True, None), ("fp32", False, None)):
    al.attention.gemm_dtype = mode
    al.attention.fused_stacks = fused
    if inner is not None:
        al.attention.stack_dtype = inner
    print(inner, end=" ")
    ts = []
    for it in range(6):
        ev[0].record()
        with torch.no_grad():
            if not fused and mode == "bf16":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    q, k = al.attention.encode(*args)
            else:
                q, k = al.attention.encode(*args)
        ev[1].record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(ev[0].elapsed_time(ev[1]))
    fl = 2 * B * (T2 * (768 * 384 * 5 + 128 * 768) + T1 * (160 * 80 * 5 + 80 * 160 * 5 + 128 * 80))
    print(mode, "fused" if fused else "torch", f"stacks {np.mean(ts):.3f} ms  ({fl / np.mean(ts) / 1e9:.1f} TFLOP/s on padded flops {fl / 1e9:.1f} G)", flush=True)
