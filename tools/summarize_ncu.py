#!/usr/bin/env python
"""Summaries of ncu outputs for profiles/ (read here, no GPU needed):
  summarize_ncu.py launches <launches.csv>          per-kernel share of the launch list
  summarize_ncu.py kernel <report.ncu-rep> [regex]  headline metrics + hottest source lines
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def launches(path):
    rows = list(csv.reader(open(path)))
    for k, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, k + 1
            break
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(lambda: [0, 0.0])
    for r in rows[start:]:
        if len(r) > iv:
            try:
                v = float(r[iv].replace(",", ""))
            except ValueError:
                continue
            n = re.sub(r"\(.*", "", r[ik])[-70:]
            d[n][0] += 1
            d[n][1] += v
    tot = sum(v[1] for v in d.values())
    print("# gpu__time_duration.sum per kernel (ncu launch list: cold-cache, serialised -> compare shares)")
    for n, (c, t) in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / tot * 100:6.2f}%  {t / 1e6:10.3f} ms  launches {c:<5d} {n}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__inst_executed_op_shared_atom.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def kernel(path, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("## kernel:", vals[hdr.index("Kernel Name")][:120])
        for i, h in enumerate(hdr):
            if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                print(f"{h:95s} {vals[i]:>18s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, out = None, []
    for r in csv.reader(src.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 7 and r[0].isdigit() and r[2] == "-":
            out.append((cur, int(r[0]), r[1].strip(), int(r[7]), int(r[4])))
    tot = sum(o[3] for o in out) or 1
    print("## hottest source lines: % of executed warp instructions, stall samples")
    for o in sorted(out, key=lambda o: -o[3])[:top]:
        print(f"{o[3] / tot * 100:5.1f}%  stalls {o[4]:>7d}  {o[0]}:{o[1]:<4d} {o[2][:100]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2])
