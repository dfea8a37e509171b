# round 2, GPU job 10: T^T-only flush + 3 units per CTA: tests, A/B against the previous candidates and the round-start build
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LUSH_LIB_PATH=$PWD/lib_exp_final2.so
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_10.log
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_10.log | tail -20
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2 3; do
  timeout 300 $B > gpurun_out/r2_ab10_new$i.json 2>/dev/null
  LUSH_LIB_PATH=$PWD/lib_exp_r2f.so timeout 300 $B > gpurun_out/r2_ab10_r2f$i.json 2>/dev/null
  LUSH_LIB_PATH=$PWD/lib_old.so timeout 300 $B > gpurun_out/r2_ab10_old$i.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("new1", "r2f1", "old1", "new2", "r2f2", "old2", "new3", "r2f3", "old3"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab10_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_10.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch10.log 2>&1; echo "ncu list rc=$?"
