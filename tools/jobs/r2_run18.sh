# round 2, GPU job 18: dgrad stash offload, second attempt (barrier observations first, head-gradient load prefetched)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LUSH_LIB_PATH=$PWD/lib_exp_dgoff.so
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2; do
  timeout 300 $B > gpurun_out/r2_ab18_dgoff$i.json 2> gpurun_out/r2_ab18_dgoff$i.err; echo "dgoff rc=$?"
  LUSH_LIB_PATH= timeout 300 $B > gpurun_out/r2_ab18_main$i.json 2>/dev/null
done
tail -5 gpurun_out/r2_ab18_dgoff1.err
python - <<'PY'
import json
for v in ("dgoff1", "main1", "dgoff2", "main2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab18_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_18.log
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_18.log | tail -10
