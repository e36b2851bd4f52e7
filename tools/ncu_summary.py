#!/usr/bin/env python
"""Summarise ncu output for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [kernel-regex] > profiles/rNN_kernel.txt

`launches` reads the CSV of `ncu --metrics gpu__time_duration.sum --csv`; `full` reads a `--set full` report through
`ncu -i ... --page raw --csv` and prints the metrics the roofline / design discussion uses.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}
    for r in rows[1:]:
        a = agg[r[ki]]
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'count':>6} {'total ms':>11} {'avg ms':>9} {'share':>7}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[0]:6d} {v[1]:11.3f} {v[1] / v[0]:9.3f} {100 * v[1] / tot:6.1f}%  {k[:150]}")


def full(path, pattern=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for vals in rows[2:]:
        if pattern and not re.search(pattern, vals[ki]):
            continue
        print(f"## {vals[ki][:160]}")
        m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
        for k in KEYS:
            if k in m and m[k][0] != "":
                print(f"{k:75s} {m[k][0]:>18s} {m[k][1]}")
        st = []
        for h, (v, u) in m.items():
            g = STALL.match(h)
            if g and v:
                try:
                    st.append((float(v.replace(",", "")), g.group(1) or g.group(2)))
                except ValueError:
                    pass
        if st:
            print("warp stall reasons (cycles per issued instruction, top 8):")
            for v, n in sorted(st, reverse=True)[:8]:
                print(f"    {n:30s} {v:8.3f}")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
