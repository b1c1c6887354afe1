#!/usr/bin/env python
"""Per-source-line hotspots of one kernel from an .ncu-rep captured with --import-source on:
joins ncu's per-SASS-instruction counters with nvdisasm's line table of the built library.

  python tools/ncu_hotspots.py <report.ncu-rep> <kernel-regex> <cubin-name e.g. binning> [top N]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern, unit = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
BY_STALL = os.environ.get("BY_STALL") == "1"   # rank lines by stall samples instead of instructions
so = os.path.join(ROOT, "dreamscene_b200", "libb200gsr.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", f"{unit}.sm_100a.cubin", so], cwd=tmp, check=True, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f"{unit}.sm_100a.cubin")], capture_output=True,
                     text=True).stdout
# per function: offset -> (file, line)
funcs, cur, line = {}, None, None
for l in dis.splitlines():
    m = re.match(r"^(_Z\w+):\s*$", l)
    if m:
        cur = funcs.setdefault(m.group(1), {}); line = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m and cur is not None and line:
        cur[int(m.group(1), 16)] = line
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
tables, t = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        t = {"name": r[1], "rows": []}; tables.append(t); continue
    if t is not None:
        t["rows"].append(r)
t = tables[0]
h = t["rows"][0]
ai, ii, si = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
ins = [(int(r[ai], 16), int(r[ii]), int(r[si])) for r in t["rows"][1:] if len(r) > ii and r[ai].startswith("0x")]
base = ins[0][0]
short = re.sub(r"\(.*", "", t["name"]).split("::")[-1].split("<")[0]
cands = [f for f in funcs if short in f]
# pick the function whose instruction count matches best
table = min(cands, key=lambda f: abs(len(funcs[f]) - len(ins))) if cands else None
off2line = funcs.get(table, {})
agg, samp = collections.Counter(), collections.Counter()
for a, n, s in ins:
    ln = off2line.get(a - base, ("?", 0)); agg[ln] += n; samp[ln] += s
tot, ts = sum(agg.values()), sum(samp.values())
srcs = {}
print(f"kernel {t['name'][:80]}\n total warp instructions {tot}, stall samples {ts}")
order = samp.most_common(top) if BY_STALL else agg.most_common(top)
for ln, _ in order:
    n = agg[ln]
    if ln[0] not in srcs:
        p = os.path.join(ROOT, "dreamscene_b200", "csrc", ln[0])
        srcs[ln[0]] = open(p).read().splitlines() if os.path.exists(p) else []
    txt = srcs[ln[0]][ln[1] - 1].strip()[:88] if 0 < ln[1] <= len(srcs[ln[0]]) else ""
    print(f"{n:>11} {100*n/tot:5.1f}%  stall {100*samp[ln]/max(ts,1):5.1f}%  {ln[0]}:{ln[1]:<4} {txt}")
