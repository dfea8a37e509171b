# round 2, GPU job 39: last check of the tree as committed (main library): smoke, full GPU suite, one bench line
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_39.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke_39.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_39.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_39.log | tail -6
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_39.json 2> gpurun_out/r2_bench_39.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_39.json").read().strip().splitlines()[-1])
print(round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), [round(x["ms_per_step"], 3) for x in d["roofline"]["per_kernel"].values()], d["roofline"]["frac"], d["clocks"])
print(d["render"]["ms_per_frame"], d["render"]["mlp_tflops_per_gpu"])
PY
