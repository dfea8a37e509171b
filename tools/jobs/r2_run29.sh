# round 2, GPU job 29: fused noise MLP with explicitly double-buffered weight batches (the compiler had serialised the
# weight loads through one register quad): tests, bench (its "noise MLP" line: 0.155-0.160 ms before), render frame
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py tests/test_parity_configs_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_29.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_29.log | tail -10
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2; do
  timeout 300 $B > gpurun_out/r2_ab29_new$i.json 2> gpurun_out/r2_ab29_new$i.err; echo "new rc=$?"
  LUSH_LIB_PATH=$PWD/lib_exp_m4.so timeout 300 $B > gpurun_out/r2_ab29_m4$i.json 2>/dev/null
done
python - <<'PY'
import json
for i in (1, 2):
  for v in ("new", "m4"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab29_{v}{i}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, i, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"]["reasons"])
    except Exception as e:
        print(v, i, "no result:", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_29.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch29.log 2>&1; echo "ncu list rc=$?"
grep -E "small_mlp" gpurun_out/r2_launches_29.csv | tail -4 | cut -d, -f5,12- | cut -c1-60,200-
timeout 300 python tools/ncu_render_target.py
