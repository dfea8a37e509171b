# round 2, GPU job 14: stash address arithmetic hoisted out of the chunk loops (forward, dgrad): tests, A/B
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
LUSH_LIB_PATH=$PWD/lib_exp_hoist.so timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_14.log
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_14.log | tail -10
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2 3; do
  LUSH_LIB_PATH=$PWD/lib_exp_hoist.so timeout 300 $B > gpurun_out/r2_ab14_hoist$i.json 2>/dev/null
  timeout 300 $B > gpurun_out/r2_ab14_main$i.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("hoist1", "main1", "hoist2", "main2", "hoist3", "main3"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab14_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
