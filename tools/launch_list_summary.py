#!/usr/bin/env python
"""Markdown summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (per kernel: launches, total
time, share).  Usage: tools/launch_list_summary.py gpurun_out/x_launches_bench.csv "title" > profiles/x.md"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    if r is hdr or r[ik] == "Kernel Name":
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    name = r[ik].split("(lompc::")[0].split("(bimpc::")[0][:70]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"# {sys.argv[2]}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv` (raw: `%s`; per-launch times are "
      "cold-cache and serialised, shares matter).\n" % sys.argv[1].split("/")[-1])
print("| launches | total us | share | kernel |\n|---|---|---|---|")
for k in sorted(tot, key=tot.get, reverse=True)[:16]:
    print(f"| {cnt[k]} | {tot[k]:.1f} | {100 * tot[k] / total:.1f}% | `{k}` |")
