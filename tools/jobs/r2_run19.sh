# round 2, GPU job 19: stash offload on the d_ready protocol: forward only / dgrad only / both, same-box A/B against main
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-render --no-cpu-baseline"
LUSH_LIB_PATH=$PWD/lib_exp_fwdoff.so timeout 600 python -m pytest tests/test_parity_configs_gpu.py tests/test_mlp_tc_gpu.py tests/test_train_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_tests_19.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_tests_19.log
for i in 1 2; do
  for v in fwdoff dgoff2 off2; do
    LUSH_LIB_PATH=$PWD/lib_exp_$v.so timeout 300 $B > gpurun_out/r2_ab19_$v$i.json 2> gpurun_out/r2_ab19_$v$i.err; echo "$v rc=$?"
  done
  timeout 300 $B > gpurun_out/r2_ab19_main$i.json 2>/dev/null
done
python - <<'PY'
import json
for i in (1, 2):
  for v in ("fwdoff", "dgoff2", "off2", "main"):
    try:
        d = json.loads(open(f"gpurun_out/r2_ab19_{v}{i}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, i, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"]["reasons"])
    except Exception as e:
        print(v, i, "no result:", e)
PY
