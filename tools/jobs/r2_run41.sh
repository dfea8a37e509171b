# round 2, GPU job 41: row tiles of the fused noise MLP, second pass: forward / backward 2 / 4 (C) and 4 / 4 (D) against 4 / 8 (B)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in C D; do
  LUSH_LIB_PATH=$PWD/lib_exp_tm$v.so timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -m gpu -q -p no:cacheprovider -k "small_mlp or train or noise" > gpurun_out/r2_tests_41$v.log 2>&1; echo "pytest $v rc=$?"
done
B="python bench.py --no-render --no-cpu-baseline --steps 10"
for v in C D B; do
  LUSH_LIB_PATH=$PWD/lib_exp_tm$v.so timeout 200 $B > gpurun_out/r2_ab41_$v.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("C", "D", "B"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab41_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3), [round(x["ms_per_step"], 3) for x in pk.values()])
    except Exception as e:
        print(v, "no result:", e)
PY
