#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full --import-source on) as text for profiles/:
the metrics the roofline argument needs and the hottest source lines (share of executed
warp instructions, stall samples).  Usage: ncu_summary.py file.ncu-rep [n_lines] > out.txt"""
import csv
import io
import subprocess
import sys

KEEP = (
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__t_sector_hit_rate.pct",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__inst_executed_op_shared_atom.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg",
)


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, check=True).stdout


def main():
    rep = sys.argv[1]
    n_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print("## kernel:", d.get("Kernel Name"))
        for k in hdr:
            if k in KEEP or "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                print("%-104s %s %s" % (k, d[k], u[k]))
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    # find the header row of the source table
    hi = next((i for i, r in enumerate(src) if "Source" in r and any("Instructions Executed" in c for c in r)), None)
    if hi is None:
        return
    h = src[hi]
    ci = h.index("Source")
    cx = next(i for i, c in enumerate(h) if c == "# Instructions Executed" or c == "Instructions Executed")
    cs = next((i for i, c in enumerate(h) if c.startswith("Warp Stall Sampling (All")), None)
    cf = next((i for i, c in enumerate(h) if c in ("File", "File Name")), None)
    cl = next((i for i, c in enumerate(h) if c in ("Line", "#")), None)
    rows = []
    tot = 0
    for r in src[hi + 1:]:
        if len(r) <= cx:
            continue
        try:
            n = int(r[cx] or 0)
        except ValueError:
            continue
        tot += n
        try:
            st = int(r[cs]) if cs is not None else 0
        except ValueError:
            st = 0
        rows.append((n, st, r[ci].strip(), r[cf] if cf is not None else "", r[cl] if cl is not None else ""))
    rows.sort(reverse=True)
    print("## hottest source lines: % of executed warp instructions, stall samples")
    for n, st, text, f, l in rows[:n_lines]:
        print("%5.1f%%  stalls %7d  %s:%s  %s" % (100. * n / max(tot, 1), st, f.split("/")[-1], l, text[:110]))


if __name__ == "__main__":
    main()
