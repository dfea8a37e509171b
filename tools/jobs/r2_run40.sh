# round 2, GPU job 40: row tiles of the fused noise MLP: forward 4 rows per CTA (A), and backward 8 (B), against 8 / 16
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
LUSH_LIB_PATH=$PWD/lib_exp_tmB.so timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -m gpu -q -p no:cacheprovider -k "small_mlp or train or noise" > gpurun_out/r2_tests_40.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r2_tests_40.log
B="python bench.py --no-render --no-cpu-baseline --steps 10"
LUSH_LIB_PATH=$PWD/lib_exp_tmA.so timeout 200 $B > gpurun_out/r2_ab40_A.json 2>/dev/null
LUSH_LIB_PATH=$PWD/lib_exp_tmB.so timeout 200 $B > gpurun_out/r2_ab40_B.json 2>/dev/null
timeout 200 $B > gpurun_out/r2_ab40_main.json 2>/dev/null
python - <<'PY'
import json
for v in ("A", "B", "main"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab40_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3), [round(x["ms_per_step"], 3) for x in pk.values()])
    except Exception as e:
        print(v, "no result:", e)
PY
