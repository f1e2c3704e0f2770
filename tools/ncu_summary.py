#!/usr/bin/env python3
"""Key counters of `ncu --page raw --csv` exports (gpurun_out/full_*.raw.csv) + stall / instruction mix of the source page.
python tools/ncu_summary.py gpurun_out/full_c2c_split_4096 [...]   -> markdown on stdout"""
import csv
import gzip
import sys
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg",
]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    d = {n: (v, u) for n, u, v in zip(names, units, vals)}
    return d


def source(path):
    stall, mix = Counter(), Counter()
    try:
        rows = list(csv.reader(gzip.open(path, "rt")))
    except Exception:
        return stall, mix
    hdr = next((i for i, r in enumerate(rows) if "Source" in r and any("Sampling" in c for c in r)), None)
    if hdr is None:
        return stall, mix
    names = rows[hdr]
    isrc = names.index("Source")
    iexec = next((i for i, c in enumerate(names) if c.startswith("# Warp Instructions Executed") or c == "Instructions Executed"), None)
    stall_cols = [(i, c) for i, c in enumerate(names) if c.startswith("stall_")]
    for r in rows[hdr + 1:]:
        if len(r) <= isrc:
            continue
        op = r[isrc].split()[0] if r[isrc].split() else ""
        if op.startswith("@"):
            op = (r[isrc].split() + ["", ""])[1]
        op = op.split(".")[0]
        if iexec is not None:
            try:
                mix[op] += int(float(r[iexec] or 0))
            except ValueError:
                pass
        for i, c in stall_cols:
            try:
                stall[c[6:]] += int(float(r[i] or 0))
            except ValueError:
                pass
    return stall, mix


for base in sys.argv[1:]:
    d = raw(base + ".raw.csv")
    print(f"## `{d['Kernel Name'][0][:150]}`  ({base.split('full_')[-1]})\n")
    print("| metric | value |\n|---|---|")
    for k in KEYS:
        if k in d:
            print(f"| {k} | {d[k][0]} {d[k][1]} |")
    try:
        tr = float(d["dram__bytes_read.sum"][0]) + float(d["dram__bytes_write.sum"][0])
        print(f"| **DRAM traffic** | {tr:.4g} {d['dram__bytes_read.sum'][1]} |")
    except Exception:
        pass
    stall, mix = source(base + ".source.csv.gz")
    if stall:
        tot = sum(stall.values()) or 1
        print("| warp-state samples | " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in stall.most_common(7)) + " |")
    if mix:
        tot = sum(mix.values()) or 1
        print("| instruction mix (warp-level) | " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in mix.most_common(9)) + " |")
    print()
