#!/usr/bin/env python
"""BASELINE.json configs[4]: batch-size sweep with the cfg3 length law, utterances sharded across the ranks.

    python tools/sweep.py                       # one GPU, B = 64 .. 4096
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py --gpus N

Prints one line per global batch: utterances/s, kernel times and HBM fractions (bench.py does the measuring).
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--batches", default="64,128,256,512,1024,2048,4096")
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()

for gb in [int(x) for x in args.batches.split(",")]:
    per = max(1, gb // args.gpus)
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", "3",
           "--batch", str(per), "--no-cpu", "--no-backward"]
    if args.gpus > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533"] + cmd[1:]
    r = subprocess.run(cmd, capture_output=True, text=True)
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        print(f"B={gb}: failed\n{r.stderr[-400:]}", flush=True)
        continue
    d = json.loads(lines[-1])
    k = d["kernels"]
    ll = next(v for n, v in k.items() if n.startswith("isp_loglik"))
    ms = next(v for n, v in k.items() if n.startswith("isp_mas"))
    print(f"global batch {gb:5d} on {args.gpus} GPU(s) ({per}/GPU): {d['value']:12.0f} utt/s  step {d['ms_per_step']*1e3:8.1f} us  "
          f"loglik {ll['ms']*1e3:7.1f} us ({ll['hbm_frac']*100:4.1f}% HBM)  mas {ms['ms']*1e3:7.1f} us ({ms['hbm_frac']*100:4.1f}% HBM)  "
          f"e2e {d['e2e']['value']:10.0f} utt/s", flush=True)
