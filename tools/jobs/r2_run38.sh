# round 2, GPU job 38: wgrad flush with four-float reductions on EVERY row alignment (leading scalars up to the first
# 16-byte boundary): full tests on the side build, same-box A/B against main (aligned rows only)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
LUSH_LIB_PATH=$PWD/lib_exp_rv5.so timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_38.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_38.log | tail -6
B="python bench.py --no-render --no-cpu-baseline"
for i in 1 2; do
  LUSH_LIB_PATH=$PWD/lib_exp_rv5.so timeout 300 $B > gpurun_out/r2_ab38_rv5$i.json 2> gpurun_out/r2_ab38_rv5$i.err; echo "rv5 rc=$?"
  timeout 300 $B > gpurun_out/r2_ab38_main$i.json 2>/dev/null
done
python - <<'PY'
import json
for i in (1, 2):
  for v in ("rv5", "main"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab38_{v}{i}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, i, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"]["reasons"])
    except Exception as e:
        print(v, i, "no result:", e)
PY
