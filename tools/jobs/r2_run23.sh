# round 2, GPU job 23: ReLU mask words stored once per step, coalesced ([warpgroup][row][16 B]); A/B against the pair-store build (m2) and main
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-render --no-cpu-baseline"
LUSH_LIB_PATH=$PWD/lib_exp_m4.so timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_23.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_23.log | tail -10
for i in 1 2 3; do
  LUSH_LIB_PATH=$PWD/lib_exp_m4.so timeout 300 $B > gpurun_out/r2_ab23_m4$i.json 2> gpurun_out/r2_ab23_m4$i.err; echo "m4 rc=$?"
  LUSH_LIB_PATH=$PWD/lib_exp_m2.so timeout 300 $B > gpurun_out/r2_ab23_main$i.json 2>/dev/null
done
python - <<'PY'
import json
for i in (1, 2, 3):
  for v in ("m4", "main")  # "main" = the m2 build here:
    try:
        d = json.loads(open(f"gpurun_out/r2_ab23_{v}{i}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, i, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"]["reasons"])
    except Exception as e:
        print(v, i, "no result:", e)
PY
