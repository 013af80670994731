#!/usr/bin/env python
"""Key metrics of every kernel in an ncu report (--set full), as text for profiles/.
usage: tools/ncu_summary.py report.ncu-rep > profiles/rNN_xxx_ncu.txt"""
import csv, subprocess, sys
KEYS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
launch__registers_per_thread launch__shared_mem_per_block_dynamic launch__occupancy_limit_shared_mem launch__occupancy_limit_registers
sm__warps_active.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed
lts__throughput.avg.pct_of_peak_sustained_elapsed l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct
smsp__thread_inst_executed_per_inst_executed.ratio smsp__inst_executed.sum""".split()
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
stall = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("-----")
    for k in ["Kernel Name", "Block Size", "Grid Size"] + KEYS:
        if k in d:
            print(f"{k:86s}{d[k]} {units[hdr.index(k)]}")
    top = sorted(((float(d[h]), h) for h in stall if d[h] not in ("", "n/a")), reverse=True)[:6]
    for v, h in top:
        print(f"{h:86s}{v:.3f} inst")
