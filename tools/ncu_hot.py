#!/usr/bin/env python
"""Top stall locations of a kernel from `ncu -i rep --page source --csv` output.  usage: ncu_hot.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
items = []; tot = {s: 0 for s in stalls}
for k, r in enumerate(rows[2:]):
    try: n = int(r[ix['# Samples']])
    except Exception: continue
    d = {s: int(r[ix[s]] or 0) for s in stalls}
    for s in stalls: tot[s] += d[s]
    items.append((k, n, r[ix['Source']], d, r[ix['Instructions Executed']]))
T = sum(x[1] for x in items) or 1
print(rows[0][1]); print("total samples", T)
print({k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
for k, n, src, d, ie in sorted(items, key=lambda x: -x[1])[:N]:
    top = [(a.replace('stall_', ''), b) for a, b in sorted(d.items(), key=lambda kv: -kv[1])[:2] if b]
    print(f"{k:5d} {n:6d} {100*n/T:5.1f}% exec={ie:>8} {src[:64]:64s} {top}")
