# round 2, GPU job 15: plan steps read from shared memory instead of indexed kernel parameters: tests, A/B
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
LUSH_LIB_PATH=$PWD/lib_exp_ssteps.so timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_15.log
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_15.log | tail -10
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2 3; do
  LUSH_LIB_PATH=$PWD/lib_exp_ssteps.so timeout 300 $B > gpurun_out/r2_ab15_ss$i.json 2>/dev/null
  timeout 300 $B > gpurun_out/r2_ab15_main$i.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("ss1", "main1", "ss2", "main2", "ss3", "main3"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab15_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
LUSH_LIB_PATH=$PWD/lib_exp_ssteps.so timeout 200 python tools/bench_kernels.py mlp 2>&1 | head -1
timeout 200 python tools/bench_kernels.py mlp 2>&1 | head -1
