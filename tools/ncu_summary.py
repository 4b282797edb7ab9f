#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): per-kernel headline metrics + SASS opcode mix.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [voxels]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
vox = float(sys.argv[2]) if len(sys.argv) > 2 else 2 * 192 ** 3
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_elapsed.avg.per_second",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic"]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("=" * 100)
    print(name)
    for k in KEYS:
        if k in hdr:
            print(f"  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    st = sorted(((float(r[hdr.index(k)] or 0), k) for k in STALL), reverse=True)[:8]
    print("  top stalls (warps per issue-active cycle): " + ", ".join(
        f"{k.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, k in st))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern = []
cur = None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 5:
        cur["rows"].append(r)
seen = set()
for k in kern:
    if k["name"] in seen or not k["rows"]:
        continue
    seen.add(k["name"])
    h = k["hdr"]
    iE, iS, iSrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    ops, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in k["rows"]:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iSrc])
        full = m.group(2) if m else "?"
        op = full.split(".")[0]
        key = full if op in ("LDS", "STS", "LDG", "STG", "LDGSTS") else op
        ops[key] += int(r[iE]); samp[key] += int(r[iS]); tot += int(r[iE])
    print("-" * 100)
    print(f"{k['name']}: {tot} warp-instr, {tot * 32 / vox:.1f} thread-instr/voxel (voxels={vox:.0f})")
    for o, c in ops.most_common(24):
        print(f"   {o:18s} {c * 32 / vox:8.1f}/voxel   stall-samples {samp[o]}")
