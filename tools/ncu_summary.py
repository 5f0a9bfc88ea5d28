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
hi = next((i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r), None)
if hi is not None:
    import collections
    h = rows[hi]
    ci, cs, ce = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    items = []
    for r in rows[hi + 1:]:
        try:
            items.append((float(r[cs]), r[ci].strip(), float(r[ce]), {h[i]: float(r[i]) for i in stall_cols if float(r[i]) > 0}))
        except (ValueError, IndexError):
            pass
    tot, totex = sum(x[0] for x in items), sum(x[2] for x in items)
    print(f"\n# top stall sites (SASS, warp-state samples): {int(tot)} samples, {int(totex)} warp-instructions executed")
    for s, ins, ex, st in sorted(items, key=lambda x: -x[0])[:16]:
        top = ", ".join(f"{k} {int(v)}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
        print(f"  {int(s):6d} {ins[:72]:72s} ex={int(ex):8d}  [{top}]")
    mix = collections.Counter()
    for s, ins, ex, st in items:
        parts = ins.split()
        if parts:
            op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            mix[op.split(".")[0]] += ex
    print("\n# instruction mix (warp-instructions executed)")
    print("  " + "  ".join(f"{op} {100 * n / totex:.1f}%" for op, n in mix.most_common(16)))
