#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of the tcgen05 MLP kernels of ONE bench.py training step:
    ncu --set full --clock-control none --import-source on -k regex:mlp_tc --launch-skip 18 -c 6 \\
        -o gpurun_out/ncu_step python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline
    python tools/ncu_step.py gpurun_out/ncu_step.ncu-rep profiles/rNN_ncu_mlp_step.json
"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = {"gpu__time_duration.sum": "time_ms", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
        "sm__cycles_elapsed.max": "sm_cycles", "launch__grid_size": "grid", "launch__registers_per_thread": "regs"}
units = dict(zip(h, rows[1]))
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}   # -> ms
kernels, total = [], 0.0
for r in rows[2:]:
    d = dict(zip(h, r))
    k = {"kernel": d["Kernel Name"].split("(")[0]}
    for m, name in want.items():
        if m in d and d[m] != "":
            v = float(d[m].replace(",", ""))
            if name.startswith("dram_r") or name.startswith("dram_w"):
                v *= scale.get(units[m], 1.0)
            if name == "time_ms":
                v *= tscale.get(units[m], 1.0)
            k[name] = v
    k["dram_bytes"] = k.get("dram_read", 0.0) + k.get("dram_write", 0.0)
    total += k["dram_bytes"]
    kernels.append(k)
json.dump({"what": "ncu --set full, tcgen05 MLP kernels of one bench.py training step (N_rand=1024 x 5 rays, 64+64)",
           "dram_bytes_per_step": total, "kernels": kernels}, open(out, "w"), indent=1)
for k in kernels:
    print(k)
print("dram bytes / step", total)
