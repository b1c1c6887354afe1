#!/usr/bin/env python
"""Summarise an .ncu-rep (one `--set full` capture) into profiles/<name>.md and
profiles/ncu_traffic.json (per-kernel DRAM bytes per launch, read by bench.py as roofline.traffic).

  python tools/ncu_summary.py gpurun_out/prof_r01_all.ncu-rep profiles/r01_ncu_summary.md cfg3_1M_1024
"""
import csv
import json
import os
import subprocess
import sys

rep, out_md, workload = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
     ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp inst"),
     ("lts__t_bytes.sum", "L2 bytes")]
STAGE = {"project_sh": "project_sh", "multisplit<count>": "project_sh", "scan_order": "scan_order", "scatter": "scatter", "sort_big": "tile_sort",
         "sort_small": "tile_sort", "composite_fwd": "composite_fwd", "composite_bwd": "composite_bwd",
         "project_bwd": "project_bwd"}


def val(r, m):
    if m not in col:
        return None, ""
    v = r[col[m]].replace(",", "")
    try:
        return float(v), units[col[m]]
    except ValueError:
        return None, ""


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


lines = [f"# ncu summary: {os.path.basename(rep)} ({workload}; `ncu --set full --clock-control none`)", "",
         "Per-launch numbers under the profiler are cold-cache and serialised: compare SHARES, not absolutes.", "",
         "| kernel | " + " | ".join(n for _, n in M) + " |", "|---|" + "---|" * len(M)]
traffic = {}
seen = set()
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].split("::")[-1].strip()
    short = name.replace("_kernel", "").replace("void ", "").split("<")[0]
    if short == "multisplit":       # one template, two kernels: <0> counts tile hits (project stage), <1> scatters the keys
        short = "multisplit<scatter>" if "<1" in name else "multisplit<count>"
    if short in seen:
        continue
    seen.add(short)
    cells = []
    for m, _ in M:
        v, u = val(r, m)
        if v is None:
            cells.append("-")
        elif "bytes" in m:
            cells.append(f"{to_bytes(v, u)/1e6:.1f} MB")
        elif "time" in m:
            cells.append(f"{to_us(v, u):.1f} us")
        elif m.endswith("inst_executed.sum"):
            cells.append(f"{v/1e6:.1f} M")
        else:
            cells.append(f"{v:.1f}")
    lines.append(f"| {short} | " + " | ".join(cells) + " |")
    rd, ru = val(r, "dram__bytes_read.sum")
    wr, wu = val(r, "dram__bytes_write.sum")
    for k, st in STAGE.items():
        if k in short and rd is not None:
            traffic[st] = traffic.get(st, 0) + int(to_bytes(rd, ru) + to_bytes(wr, wu))
open(out_md, "w").write("\n".join(lines) + "\n")
tp = os.path.join(os.path.dirname(out_md), "ncu_traffic.json")
allt = json.load(open(tp)) if os.path.exists(tp) else {}
allt[workload] = traffic
json.dump(allt, open(tp, "w"), indent=1, sort_keys=True)
print("\n".join(lines))
print(traffic)
