#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall-sample shares from an ncu report.
usage: tools/ncu_source_top.py report.ncu-rep kernel_name [top_n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, data = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; ii = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ii:
        continue
    if r[0].strip().isdigit():       # a CUDA source line (SASS rows under it have an empty first column)
        key = (cur_file, int(r[0]))
        try:
            data[key] = [int(r[ii] or 0), int(r[sm] or 0), r[1].strip()]
        except ValueError:
            pass
tot_i = sum(v[0] for v in data.values()) or 1
tot_s = sum(v[1] for v in data.values()) or 1
print(f"kernel {kern}: {tot_i} warp instructions, {tot_s} samples")
for (f, ln), v in sorted(data.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[1] / tot_s * 100:6.2f}% smp {v[0] / tot_i * 100:6.2f}% inst  {f}:{ln:<5d} {v[2][:110]}")
