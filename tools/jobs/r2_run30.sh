# round 2, GPU job 30: head_grads with rolling prefetch (every stash word's register refilled for the next 32-row block
# as soon as it is converted): tests, launch list (before: 137 + 86 us), same-box A/B
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LUSH_LIB_PATH=$PWD/lib_exp_hg.so
timeout 900 python -m pytest tests/test_mlp_tc_gpu.py tests/test_parity_configs_gpu.py tests/test_train_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_30.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_30.log | tail -10
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_30.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch30.log 2>&1; echo "ncu list rc=$?"
grep -E "head_grads" gpurun_out/r2_launches_30.csv | tail -4 | awk -F'","' '{print $5, $NF}' | cut -c1-40,120-
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2; do
  timeout 300 $B > gpurun_out/r2_ab30_hg$i.json 2> gpurun_out/r2_ab30_hg$i.err; echo "hg rc=$?"
  LUSH_LIB_PATH=$PWD/lib_exp_m4.so timeout 300 $B > gpurun_out/r2_ab30_m4$i.json 2>/dev/null
done
python - <<'PY'
import json
for i in (1, 2):
  for v in ("hg", "m4"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab30_{v}{i}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, i, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"]["reasons"])
    except Exception as e:
        print(v, i, "no result:", e)
PY
