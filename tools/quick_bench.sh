#!/bin/bash
# quick bench: cfg3 + cfg3d summary
for w in cfg3 cfg3d; do python bench.py --workload $w --no-cpu --no-backward 2>/dev/null > gpurun_out/q_$w.json || python bench.py --workload $w > gpurun_out/q_$w.json 2>/dev/null; done
python - <<PY
import json
for w in ["cfg3","cfg3d"]:
    d=json.loads(open(f"gpurun_out/q_{w}.json").read().strip().splitlines()[-1]); print(w, round(d["value"]), round(d["ms_per_step"],4), {k[:10]:(round(v["ms"]*1e3,1),round(v["hbm_frac"],3)) for k,v in d["kernels"].items()})
PY
