#!/usr/bin/env python
"""Condense an `ncu --set full` report into the handful of numbers DESIGN.md / bench.py quote.

    python tools/ncu_summary.py gpurun_out/prof_mas.ncu-rep [more.ncu-rep ...] > profiles/rNN_xxx.md

Runs `ncu -i <rep> --page raw --csv` (no GPU needed) and prints one markdown table per
profiled launch.  `--json` prints {kernel: dram bytes per launch} instead (profiles/traffic.json).
"""
from __future__ import annotations

import csv
import io
import json
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor HMMA sub-pipe %"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("smsp__cycles_active.avg", "SMSP cycles active (avg)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
]


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit)
    return None if mult is None else float(val.replace(",", "")) * mult


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    as_json = "--json" in sys.argv
    traffic = {}
    for rep in args:
        hdr, units, rows = read(rep)
        col = {h: i for i, h in enumerate(hdr)}
        if not as_json:
            print(f"## {rep.split('/')[-1]}\n")
        for n, r in enumerate(rows):
            name = r[col["Kernel Name"]]
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            traffic.setdefault(name, []).append((rd or 0) + (wr or 0))
            if as_json:
                continue
            print(f"### launch {n}: `{name}`  grid {r[col['Grid Size']]} block {r[col['Block Size']]}\n")
            print("| metric | value |\n|---|---|")
            for key, label in WANT:
                if key in col:
                    print(f"| {label} (`{key}`) | {r[col[key]]} {units[col[key]]} |")
            if rd is not None and wr is not None:
                print(f"| **DRAM traffic per launch** | {(rd + wr) / 1e6:.1f} MB |")
            print()
    if as_json:
        print(json.dumps({k: sum(v) / len(v) for k, v in traffic.items()}, indent=1))


if __name__ == "__main__":
    main()
