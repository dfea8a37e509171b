#!/usr/bin/env python
"""Summarise an ncu report: headline metrics + the SASS lines with the most stall samples.
usage: tools/ncu_top.py report.ncu-rep [ntop]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for r in rows[2:]:
    d = dict(zip(h, zip(r, u)))
    print("==", d.get("Kernel Name", ("?",))[0][:90])
    for k in d:
        if any(k == x or k.startswith(x + ".") and k.count(".") <= x.count(".") + 0 for x in KEYS) or k in KEYS:
            print(f"  {k:85s} {d[k][0]:>18s} {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
h = rows[hi]
ai, si, wi, ii = h.index("Address"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
data = [r for r in rows[hi + 1:] if len(r) == len(h) and (r[wi] or '0').isdigit()]   # the header repeats per kernel
tot = sum(int(r[wi] or 0) for r in data)
print("total samples", tot)
for r in sorted(data, key=lambda r: -int(r[wi] or 0))[:ntop]:
    print(f"{int(r[wi] or 0):7d} {100*int(r[wi] or 0)/max(tot,1):5.1f}% {r[ii]:>10s} {r[ai][-5:]} {r[si][:100]}")
