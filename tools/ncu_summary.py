#!/usr/bin/env python
"""Key counters + top warp-stall sites of a `ncu --set full --import-source on` capture:  python tools/ncu_summary.py file.ncu-rep [kernel index]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
k = vals[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_issue_stalled_long_scoreboard.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{units[i]}] = {k[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if rows:
    h = rows[0]
    def col(name):
        for i, c in enumerate(h):
            if c.strip().startswith(name):
                return i
        return None
    ci, cs, ce = col("Source"), col("# Samples") or col("Warp Stall Sampling (All"), col("Instructions Executed")
    if ci is not None and cs is not None:
        items = []
        for r in rows[1:]:
            try:
                items.append((float(r[cs]), r[ci], r[ce] if ce is not None else ""))
            except (ValueError, IndexError):
                pass
        tot = sum(x[0] for x in items)
        print("\n# top stall sites (SASS, warp-state samples), total", tot)
        for s, ins, ex in sorted(items, reverse=True)[:14]:
            print(f"  {int(s):6d} {ins[:90]:90s} ex={ex}")
