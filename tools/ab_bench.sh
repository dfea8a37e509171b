#!/usr/bin/env bash
# Same-box A/B of two builds of liblushnerf_b200.so (box-to-box variance in the pool is ~4 %, larger than most
# kernel-level changes).  Build the baseline first, copy it aside, rebuild with the change, then on the GPU box:
#     cp lush_nerf_b200/liblushnerf_b200.so lib_old.so      # baseline (git-ignored, travels with the snapshot)
#     ... edit, python -m lush_nerf_b200.build ...
#     gpurun --timeout 900 -- 'bash tools/ab_bench.sh 2'
# Runs alternate new / old so that both see the same thermal state; LUSH_LIB_PATH selects the library at import.
set -u
N=${1:-2}
mkdir -p gpurun_out
for i in $(seq 1 "$N"); do
  timeout 200 python bench.py --no-cpu-baseline --no-render > "gpurun_out/ab_new_$i.log" 2>/dev/null
  LUSH_LIB_PATH="$PWD/lib_old.so" timeout 200 python bench.py --no-cpu-baseline --no-render > "gpurun_out/ab_old_$i.log" 2>/dev/null
done
python - "$N" <<'PY'
import json, sys
n = int(sys.argv[1])
for v in ("new", "old"):
    for i in range(1, n + 1):
        try:
            d = json.loads(open(f"gpurun_out/ab_{v}_{i}.log").read().strip().splitlines()[-1])
            pk = d["roofline"]["per_kernel"]
            print(v, round(d["value"]), round(d["ms_per_step"], 3), [round(x["ms_per_step"], 3) for x in pk.values()])
        except Exception as e:
            print(v, i, "no result:", e)
PY
