# round 2, GPU job 1: full GPU test suite on the new build (3 epilogue warpgroups), same suite on the previous build if
# that fails, same-box A/B of the two builds through bench.py, stress sweep, launch list
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -2; nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_tests_new.log 2>&1; RC=$?; echo "pytest rc=$RC" >> gpurun_out/r2_tests_new.log
tail -12 gpurun_out/r2_tests_new.log
if [ $RC -ne 0 ]; then
  LUSH_LIB_PATH=$PWD/lib_old.so timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_old.log 2>&1; echo "pytest(old) rc=$?" >> gpurun_out/r2_tests_old.log
  tail -12 gpurun_out/r2_tests_old.log
fi
timeout 600 python bench.py > gpurun_out/r2_bench_new.json 2> gpurun_out/r2_bench_new.err; echo "bench(new) rc=$?"
LUSH_LIB_PATH=$PWD/lib_old.so timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_old.json 2> gpurun_out/r2_bench_old.err; echo "bench(old) rc=$?"
timeout 300 python bench.py --no-cpu-baseline --no-render > gpurun_out/r2_bench_new2.json 2>/dev/null
LUSH_LIB_PATH=$PWD/lib_old.so timeout 300 python bench.py --no-cpu-baseline --no-render > gpurun_out/r2_bench_old2.json 2>/dev/null
python - <<'PY'
import json
for v in ("new", "old", "new2", "old2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        r = d.get("render") or {}
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3), "e2e", round(d["e2e"]["value"]),
              [round(x["ms_per_step"], 3) for x in pk.values()], "render", r.get("ms_per_frame"), d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
tail -3 gpurun_out/r2_bench_new.err
timeout 600 python bench.py --mode stress > gpurun_out/r2_stress_1.json 2> gpurun_out/r2_stress_1.err; echo "stress rc=$?"
tail -3 gpurun_out/r2_stress_1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_1.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu rc=$?"
