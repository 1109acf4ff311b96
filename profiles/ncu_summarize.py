#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw + source pages) into text: key metrics, stall reasons,
per-opcode executed-instruction histogram, hottest SASS lines.
usage: ncu_summarize.py report.ncu-rep [pixels_per_launch]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pixels = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg", "lts__t_bytes.sum"]
print("== metrics (per launch) ==")
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print("%-70s %-12s %s" % (w, units[i], " ".join(r[i] for r in data)))
print("== warp stall reasons (issue-stalled per issue-active) ==")
for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
        print("%-28s %s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), data[0][i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))[2:]
tot = 0
byop = collections.Counter()
seen = set()
lines = []
for r in rows:
    if len(r) < 6 or not r[0].startswith("0x"):
        continue
    if r[0] in seen:
        break
    seen.add(r[0])
    n = int(r[5])
    tot += n
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
    op = m.group(2) if m else r[1]
    byop[op.split(".")[0]] += n
    lines.append((n, int(r[4]), r[1].strip()))
print("== executed warp-instructions: %d total%s ==" % (tot, (" = %.2f thread-instr/pixel" % (tot * 32 / pixels)) if pixels else ""))
for op, n in byop.most_common(28):
    print("%-12s %6.2f%%%s" % (op, 100.0 * n / tot, ("  %.2f /pixel" % (n * 32 / pixels)) if pixels else ""))
print("== SASS lines with most stall samples ==")
for n, smp, txt in sorted(lines, key=lambda t: -t[1])[:25]:
    print("%8d samples %12d exec  %s" % (smp, n, txt))
