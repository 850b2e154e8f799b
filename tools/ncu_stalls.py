#!/usr/bin/env python
"""Stall breakdown (per issued instruction) and the headline utilisation metrics of one .ncu-rep.
Usage: tools/ncu_stalls.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed_op_branch.sum"]
for rep in sys.argv[1:]:
    out = (open(rep).read() if rep.endswith(".csv") else
           subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        d = dict(zip(h, v))
        un = dict(zip(h, u))
        print("==", rep, d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print(f"  {k:70s} {d[k]:>16s} {un[k]}")
        st = [(float(d[k].replace(",", "")), k) for k in h if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
        tot = sum(x for x, _ in st)
        print(f"  stalls per issue (sum {tot:.2f}):")
        for x, k in sorted(st, reverse=True)[:12]:
            print(f"    {k.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {x:.3f}")
