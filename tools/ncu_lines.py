#!/usr/bin/env python
"""Attribute ncu warp-stall samples to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]
    (NCU_KERNEL=<regex> when the demangled name in the report differs from the mangled substring in the cubin)

ncu's SASS page gives samples per instruction address; `nvdisasm -gi` gives the source line (incl. inlining) of every SASS
offset of the same cubin (built with -lineinfo).  The first SASS row of the report is offset 0 of the kernel.
"""
import csv
import os
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + os.environ.get("NCU_KERNEL", kname)],
                     capture_output=True, text=True).stdout   # reports with several kernels: keep the one asked for
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ai, si, ni = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
sass = [(int(r[ai], 16), r[si], int(r[ni] or 0), r) for r in rows[2:] if len(r) > ni and r[ai].startswith("0x")]
base = sass[0][0]

dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout
lines, cur, inside = {}, None, False
for ln in dis.splitlines():
    if ln.startswith(".text.") or ".section\t.text." in ln or ln.strip().startswith("//--------------------- .text."):
        inside = kname in ln
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', ln)
        cur = (m.group(1).split("/")[-1], int(m.group(2)), tuple((f.split("/")[-1], int(l)) for f, l in inl))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur:
        lines[int(m.group(1), 16)] = cur

agg, agg_outer, total = defaultdict(lambda: [0, defaultdict(int)]), defaultdict(int), 0
for addr, text, n, r in sass:
    loc = lines.get(addr - base)
    total += n
    key = loc[:2] if loc else ("?", 0)
    agg[key][0] += n
    for h, i in stall_cols.items():
        try:
            agg[key][1][h] += int(r[i] or 0)
        except ValueError:
            pass
    outer = (loc[2][-1] if loc and loc[2] else key)
    agg_outer[outer] += n
print(f"total samples {total}")
print("== by innermost source line ==")
for key, (n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{100 * n / total:5.1f}%  {key[0]}:{key[1]:<5d} {tops}")
print("== by outermost (kernel-body) line ==")
for key, n in sorted(agg_outer.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{100 * n / total:5.1f}%  {key[0]}:{key[1]}")
