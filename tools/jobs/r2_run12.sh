# round 2, GPU job 12: evidence with the final build: smoke(), tests, default bench line, reference arm, stress sweep,
# ncu full capture of the MLP kernels of one step + launch list
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_tests_12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_12.log
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_12.log | tail -10
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference arm rc=$?"
timeout 600 python bench.py --mode stress > gpurun_out/r2_stress_final.json 2> gpurun_out/r2_stress_final.err; echo "stress rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_tc --launch-skip 18 -c 6 -o gpurun_out/r2_ncu_step python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_step.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch12.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "traffic", d["roofline"]["traffic"], "render", d["render"]["ms_per_frame"], d["clocks"], d["cpu_baseline"]["value"])
r = json.loads(open("gpurun_out/r2_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
