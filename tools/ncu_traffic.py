#!/usr/bin/env python
"""DRAM bytes per K7/K8 launch pair (k_sweep_score + k_sweep_update) from an `ncu --set full` report -> profiles/r02_traffic.json,
keyed by bench.py's workload name (bench.py reads it for `roofline.traffic`).
usage: tools/ncu_traffic.py report.ncu-rep <workload name> <W> <H> <S>"""
import csv, json, os, subprocess, sys
rep, workload, W, H, S = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
acc = {"k_sweep_score": [0, 0.0, 0.0], "k_sweep_update": [0, 0.0, 0.0]}
for r in rows[2:]:
    for k in acc:
        if r[ki].startswith(k + "("):
            acc[k][0] += 1; acc[k][1] += float(r[ri]) * scale[units[ri]]; acc[k][2] += float(r[wi]) * scale[units[wi]]
n = min(v[0] for v in acc.values())
assert n > 0, "no sweep launches in the report"
read = sum(v[1] / v[0] for v in acc.values()); write = sum(v[2] / v[0] for v in acc.values())
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
doc = json.load(open(path)) if os.path.exists(path) else {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per K7/K8 launch PAIR (k_sweep_score + k_sweep_update, "
                                                           "averaged over the launches captured) from ncu --set full, keyed by bench.py workload name; written by tools/ncu_traffic.py"}
doc[workload] = {"k_sweep_pair": {"dram_read_bytes": int(read), "dram_write_bytes": int(write), "algorithmic_bytes": int(W * H * (218 + 4 * S) / 2),
                                  "scratch_bytes_written_and_read_once": int(((W * ((H + 1) // 2)) * (9 * S + 9) * 4)), "launch_pairs_averaged": n,
                                  "source": os.path.basename(rep)}}
json.dump(doc, open(path, "w"), indent=1)
print(json.dumps(doc[workload]))
