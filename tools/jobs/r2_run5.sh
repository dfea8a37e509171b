# round 2, GPU job 5: current build (T trick, LPT wgrad schedule, tiled fv_post, recurrence in enc_bwd, leaner compositing):
# tests, same-box A/B against the round-start build, stress sweep
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LUSH_LIB_PATH=$PWD/lib_exp_r2d.so
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_tests_5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_5.log
grep -E "passed|failed|^FAILED|^ERROR|tc vs|config[01] tc|fused-cdf det frame|2-rank|^E  " gpurun_out/r2_tests_5.log | tail -40
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2; do
  timeout 300 $B > gpurun_out/r2_ab5_new$i.json 2>/dev/null
  LUSH_LIB_PATH=$PWD/lib_old.so timeout 300 $B > gpurun_out/r2_ab5_old$i.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("new1", "old1", "new2", "old2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab5_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
timeout 600 python bench.py --mode stress > gpurun_out/r2_stress_5.json 2> gpurun_out/r2_stress_5.err; echo "stress rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_stress_5.json").read().strip().splitlines()[-1])
    for k in d["kernels"]:
        print(f'{k["kernel"]:58s} S={k["S"]:3d} {k["us"]:8.1f} us {k["GBs"]:8.1f} GB/s {k["frac_of_hbm"]:.2f}')
except Exception as e:
    print("stress: no result", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_5.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch5.log 2>&1; echo "ncu rc=$?"
