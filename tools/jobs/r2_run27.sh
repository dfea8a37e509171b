# round 2, GPU job 27: 128-bit stash loads in head_grads / enc_bwd / ray_grad_reduce (were four 32-bit loads per lane):
# full tests, same-box A/B against the build before (lib_exp_m4.so), launch list of the side kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_27.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_27.log | tail -10
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2 3; do
  timeout 300 $B > gpurun_out/r2_ab27_new$i.json 2> gpurun_out/r2_ab27_new$i.err; echo "new rc=$?"
  LUSH_LIB_PATH=$PWD/lib_exp_m4.so timeout 300 $B > gpurun_out/r2_ab27_m4$i.json 2>/dev/null
done
python - <<'PY'
import json
for i in (1, 2, 3):
  for v in ("new", "m4"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab27_{v}{i}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, i, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"]["reasons"])
    except Exception as e:
        print(v, i, "no result:", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_27.csv python bench.py --steps 2 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_launch27.log 2>&1; echo "ncu list rc=$?"
grep -E "head_grads|enc_bwd|ray_grad_reduce" gpurun_out/r2_launches_27.csv | tail -6 | cut -d, -f5,12-
