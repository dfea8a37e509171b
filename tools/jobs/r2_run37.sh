# round 2, GPU job 37 (final numbers of the round, after the vectorised wgrad flush): smoke, default bench (CPU baseline + render), reference arm, launch list,
# ncu --set full of the six chain-kernel launches of one step summarised on the box (the report exceeds the merge limit)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_37.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2_tests_37.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_37.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke_37.log
timeout 900 python bench.py > gpurun_out/r2_bench_37.json 2> gpurun_out/r2_bench_37.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_37.json 2> gpurun_out/r2_bench_ref_37.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_37.json").read().strip().splitlines()[-1])
print(round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), [round(x["ms_per_step"], 3) for x in d["roofline"]["per_kernel"].values()], d["roofline"]["frac"], d["clocks"])
print(d["render"]); print(d["cpu_baseline"])
r = json.loads(open("gpurun_out/r2_bench_ref_37.json").read().strip().splitlines()[-1]); print("ref", r["value"], r.get("cpu_baseline"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_37.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch37.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_tc --launch-skip 18 -c 6 -o /tmp/r2_ncu_step37 python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_step37.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_step.py /tmp/r2_ncu_step37.ncu-rep gpurun_out/r2_ncu_mlp_step37.json > gpurun_out/r2_ncu_mlp_step37.txt 2>&1
python tools/ncu_top.py /tmp/r2_ncu_step37.ncu-rep > gpurun_out/r2_ncu_mlp_step37_top.txt 2>&1
python tools/ncu_source_top.py /tmp/r2_ncu_step37.ncu-rep --top 14 > gpurun_out/r2_ncu_mlp_step37_source.txt 2>&1
wc -c gpurun_out/r2_ncu_mlp_step37*
