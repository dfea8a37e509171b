# round 2, GPU job 7: head gradients folded into the wgrad T job: tests, A/B, driver with the brighter scene
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LUSH_LIB_PATH=$PWD/lib_exp_r2g.so
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_tests_7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_7.log
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_7.log | tail -20
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2; do
  timeout 300 $B > gpurun_out/r2_ab7_new$i.json 2>/dev/null
  LUSH_LIB_PATH=$PWD/lib_exp_r2f.so timeout 300 $B > gpurun_out/r2_ab7_prev$i.json 2>/dev/null
  LUSH_LIB_PATH=$PWD/lib_old.so timeout 300 $B > gpurun_out/r2_ab7_old$i.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("new1", "prev1", "old1", "new2", "prev2", "old2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab7_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
timeout 900 python tools/run_reference_driver.py --run --iters 20 --out gpurun_out/r2_driver > gpurun_out/r2_driver.log 2>&1; echo "driver rc=$?"
tail -1 gpurun_out/r2_driver.log | cut -c1-900
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_7.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch7.log 2>&1; echo "ncu rc=$?"
