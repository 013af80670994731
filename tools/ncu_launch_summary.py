#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: tools/ncu_launch_summary.py launches.csv [skip_first_n_launches] > profiles/rNN_launches.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.OrderedDict()
print("kernel,grid,block,time_us")
for r in rows:
    name = r[4].split("(")[0]
    t = float(r[-1]) / 1e3
    print(f'{name},"{r[8]}","{r[7]}",{t:.1f}')
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print("\n# aggregated\nkernel,launches,total_ms,share_pct")
for k, (n, t) in agg.items():
    print(f"{k},{n},{t / 1e3:.3f},{100 * t / tot:.2f}")
print(f"TOTAL,{sum(a[0] for a in agg.values())},{tot / 1e3:.3f},100.00")
