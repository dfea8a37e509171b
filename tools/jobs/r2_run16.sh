# round 2, GPU job 16: persistent ray_grad_reduce (weights staged once per block, 9-shuffle column sums): tests, launch list
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LUSH_LIB_PATH=$PWD/lib_exp_small.so
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_16.log
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_16.log | tail -10
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_16.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch16.log 2>&1; echo "ncu list rc=$?"
grep -E "ray_grad_reduce|enc_bwd|head_grads" gpurun_out/r2_launches_16.csv | tail -6 | cut -d, -f5,12-
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2; do
  timeout 300 $B > gpurun_out/r2_ab16_small$i.json 2>/dev/null
  LUSH_LIB_PATH= timeout 300 $B > gpurun_out/r2_ab16_main$i.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("small1", "main1", "small2", "main2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab16_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
