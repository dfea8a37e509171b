# round 2, GPU job 3: side kernels (head gradients, fused ray gradients) + T trick: tests, same-box A/B against the
# previous build, the unmodified reference driver through the shim
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LUSH_LIB_PATH=$PWD/lib_exp_r2b.so
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_tests_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_3.log
grep -E "passed|failed|^FAILED|^ERROR|tc vs|config[01] tc|fused-cdf det frame|2-rank|^E  " gpurun_out/r2_tests_3.log | tail -40
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2; do
  timeout 300 $B > gpurun_out/r2_ab3_new$i.json 2>/dev/null
  LUSH_LIB_PATH=$PWD/lib_old.so timeout 300 $B > gpurun_out/r2_ab3_old$i.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("new1", "old1", "new2", "old2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab3_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
timeout 900 python tools/run_reference_driver.py --run --iters 20 --out gpurun_out/r2_driver > gpurun_out/r2_driver.log 2>&1; echo "driver rc=$?"
tail -4 gpurun_out/r2_driver.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_3.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch3.log 2>&1; echo "ncu rc=$?"
