# round 2, GPU job 9: final build of the round: tests, A/B (noise branch on a side stream vs inline, vs round-start build),
# ncu evidence (full capture of the MLP kernels of one step, launch list, stress kernels)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LUSH_LIB_PATH=$PWD/lib_exp_final.so
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_tests_9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_9.log
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_9.log | tail -20
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2; do
  timeout 300 $B > gpurun_out/r2_ab9_side$i.json 2>/dev/null
  LUSH_NOISE_STREAM=0 timeout 300 $B > gpurun_out/r2_ab9_inline$i.json 2>/dev/null
  LUSH_NOISE_STREAM=0 LUSH_LIB_PATH=$PWD/lib_old.so timeout 300 $B > gpurun_out/r2_ab9_old$i.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("side1", "inline1", "old1", "side2", "inline2", "old2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab9_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_tc --launch-skip 18 -c 6 -o gpurun_out/r2_ncu_step python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_step.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch9.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"composite|sample_" --launch-skip 5 -c 5 -o gpurun_out/r2_ncu_stress python tools/ncu_stress_target.py > gpurun_out/r2_ncu_stress.log 2>&1; echo "ncu stress rc=$?"
ls -la gpurun_out/*.ncu-rep
