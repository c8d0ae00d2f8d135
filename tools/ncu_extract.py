#!/usr/bin/env python
"""Extract the metrics quoted in profiles/ from an .ncu-rep (ncu --page raw / --page source as CSV).
usage: python tools/ncu_extract.py report.ncu-rep [units_per_launch] -> prints a JSON summary (first kernel)"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def sass(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    ops = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr):
            break  # next kernel
        src = r[idx["Source"]].strip()
        n = int(r[idx["Instructions Executed"]] or 0)
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)*)", src)
        op = m.group(2) if m else src[:10]
        base = op.split(".")[0]
        if base in ("LDS", "STS", "LDL", "STL"):
            base = ".".join(op.split(".")[:2]) if "." in op else op
        if base == "IMAD" and ".MOV" in op:
            base = "IMAD.MOV"
        ops[base] += n
    return ops


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    hdr, units_row, rows = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    r = rows[0]

    def f(name):
        v = r[idx[name]].replace(",", "") if name in idx else ""
        try:
            return float(v)
        except ValueError:
            return v
    out = {"report": rep, "kernel": r[idx["Kernel Name"]].split("(")[0]}
    for k in ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
              "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sass__inst_executed_local_loads",
              "sass__inst_executed_local_stores", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"):
        if k in idx:
            out[k] = f(k)
            out[k + ".unit"] = units_row[idx[k]]
    stalls = {}
    for h in hdr:
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
        if m and r[idx[h]] not in ("", "n/a"):
            stalls[m.group(1)] = round(float(r[idx[h]].replace(",", "")), 3)
    out["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    ops = sass(rep)
    tot = sum(ops.values())
    if units:
        out["warp_instructions_per_32_units"] = round(tot / (units / 32.0), 1)
        out["opcode_mix_per_32_units"] = {k: round(v / (units / 32.0), 1) for k, v in ops.most_common(18)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
