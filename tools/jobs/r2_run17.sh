# round 2, GPU job 17: dgrad stash offload (the prologue warps copy every dZ from tensor memory to the gradient stash):
# MLP / parity tests on the side build, then a same-box A/B against the main library.
# Also confirms the main library on the tests added last (index batches, check_numerics).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_17_main.log 2>&1; echo "pytest(main,train) rc=$?"
tail -3 gpurun_out/r2_tests_17_main.log
export LUSH_LIB_PATH=$PWD/lib_exp_dgoff.so
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_17.log
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_17.log | tail -10
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2; do
  timeout 300 $B > gpurun_out/r2_ab17_dgoff$i.json 2>/dev/null
  LUSH_LIB_PATH= timeout 300 $B > gpurun_out/r2_ab17_main$i.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("dgoff1", "main1", "dgoff2", "main2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab17_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
