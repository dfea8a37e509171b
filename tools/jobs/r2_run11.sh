# round 2, GPU job 11: encoder with angle doubling between exact octaves (tc forward prologue): tests, A/B incl. render
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LUSH_LIB_PATH=$PWD/lib_exp_enc.so
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_tests_11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_11.log
grep -E "passed|failed|^FAILED|^ERROR|^E  |tc vs fp16|config[01] tc" gpurun_out/r2_tests_11.log | tail -20
unset LUSH_LIB_PATH
B="python bench.py --no-cpu-baseline"
for i in 1 2; do
  LUSH_LIB_PATH=$PWD/lib_exp_enc.so timeout 300 $B > gpurun_out/r2_ab11_enc$i.json 2>/dev/null
  timeout 300 $B > gpurun_out/r2_ab11_main$i.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("enc1", "main1", "enc2", "main2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab11_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        r = d.get("render") or {}
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], "render", r.get("ms_per_frame"), (r.get("clocks") or {}).get("sm_mhz"), d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
LUSH_LIB_PATH=$PWD/lib_exp_enc.so timeout 200 python tools/bench_kernels.py mlp 2>&1 | head -1
timeout 200 python tools/bench_kernels.py mlp 2>&1 | head -1
