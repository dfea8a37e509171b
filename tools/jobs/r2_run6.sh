# round 2, GPU job 6: fv_post v3 + compositing occupancy floor: tests, A/B, stress; black-frame diagnosis of the driver run
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LUSH_LIB_PATH=$PWD/lib_exp_r2f.so
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_tests_6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_6.log
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_6.log | tail -20
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2; do
  timeout 300 $B > gpurun_out/r2_ab6_new$i.json 2>/dev/null
  LUSH_LIB_PATH=$PWD/lib_old.so timeout 300 $B > gpurun_out/r2_ab6_old$i.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("new1", "old1", "new2", "old2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab6_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
timeout 600 python bench.py --mode stress > gpurun_out/r2_stress_6.json 2> gpurun_out/r2_stress_6.err; echo "stress rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_stress_6.json").read().strip().splitlines()[-1])
    for k in d["kernels"]:
        if k["kernel"].startswith("composite"):
            print(f'{k["kernel"]:58s} S={k["S"]:3d} {k["us"]:8.1f} us {k["GBs"]:8.1f} GB/s {k["frac_of_hbm"]:.2f}')
except Exception as e:
    print("stress: no result", e)
PY
LUSH_NUMERIC_GUARD=sync timeout 900 python tools/run_reference_driver.py --run --iters 20 --keep-ckpt --out gpurun_out/r2_driver > gpurun_out/r2_driver.log 2>&1; echo "driver rc=$?"
tail -3 gpurun_out/r2_driver.log | cut -c1-600
CK=$(ls gpurun_out/r2_driver/logs/shim/*.tar | tail -1)
timeout 300 python tools/debug_render_stats.py $CK > gpurun_out/r2_debug_render_ckpt.log 2>&1; tail -26 gpurun_out/r2_debug_render_ckpt.log
rm -f gpurun_out/r2_driver/logs/shim/*.tar
