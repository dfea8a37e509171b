# round 2, GPU job 36: wgrad flush with four-float reductions (red.global.add.v4.f32) where the output row is 16-byte aligned
# head_grads launch: tests, same-box A/B against main, launch list
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
LUSH_LIB_PATH=$PWD/lib_exp_rv4.so timeout 900 python -m pytest tests/test_mlp_tc_gpu.py tests/test_parity_configs_gpu.py tests/test_train_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_36.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_36.log | tail -10
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2 3; do
  LUSH_LIB_PATH=$PWD/lib_exp_rv4.so timeout 300 $B > gpurun_out/r2_ab36_rv4$i.json 2> gpurun_out/r2_ab36_rv4$i.err; echo "rv4 rc=$?"
  timeout 300 $B > gpurun_out/r2_ab36_main$i.json 2>/dev/null
done
python - <<'PY'
import json
for i in (1, 2, 3):
  for v in ("rv4", "main"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab36_{v}{i}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, i, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"]["reasons"])
    except Exception as e:
        print(v, i, "no result:", e)
PY
LUSH_LIB_PATH=$PWD/lib_exp_rv4.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_36.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch36.log 2>&1; echo "ncu list rc=$?"
grep -E "wgrad" gpurun_out/r2_launches_36.csv | tail -4 | awk -F'","' '{print $5, $NF}' | cut -c1-40,120-
