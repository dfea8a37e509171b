# round 2, GPU job 2: full GPU test suite, what-bounds-the-chain experiments (no weight stream / no stash stores),
# stress sweep, the unmodified reference driver through the shim
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_tests_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_2.log
grep -E "passed|failed|^FAILED|^ERROR|tc vs|config[01] tc|fused-cdf det frame|2-rank" gpurun_out/r2_tests_2.log | tail -30
export LUSH_BENCH_ALLOW_NONFINITE=1
B="python bench.py --no-render --no-cpu-baseline --steps 10"
timeout 300 $B > gpurun_out/r2_exp_base.json 2>/dev/null
LUSH_CHAIN_DEBUG=1 timeout 300 $B > gpurun_out/r2_exp_noweights.json 2>/dev/null
LUSH_LIB_PATH=$PWD/lib_exp_nostash.so timeout 300 $B > gpurun_out/r2_exp_nostash.json 2>/dev/null
LUSH_CHAIN_DEBUG=1 LUSH_LIB_PATH=$PWD/lib_exp_nostash.so timeout 300 $B > gpurun_out/r2_exp_neither.json 2>/dev/null
timeout 300 $B > gpurun_out/r2_exp_base2.json 2>/dev/null
unset LUSH_BENCH_ALLOW_NONFINITE
python - <<'PY'
import json
for v in ("base", "noweights", "nostash", "neither", "base2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_exp_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
timeout 200 python tools/bench_kernels.py mlp > gpurun_out/r2_exp_fwdonly_base.jsonl 2>&1
LUSH_CHAIN_DEBUG=1 timeout 200 python tools/bench_kernels.py mlp > gpurun_out/r2_exp_fwdonly_noweights.jsonl 2>&1
cat gpurun_out/r2_exp_fwdonly_base.jsonl gpurun_out/r2_exp_fwdonly_noweights.jsonl
timeout 600 python bench.py --mode stress > gpurun_out/r2_stress_2.json 2> gpurun_out/r2_stress_2.err; echo "stress rc=$?"
tail -3 gpurun_out/r2_stress_2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_stress_2.json").read().strip().splitlines()[-1])
    for k in d["kernels"]:
        print(f'{k["kernel"]:58s} S={k["S"]:3d} {k["us"]:8.1f} us {k["GBs"]:8.1f} GB/s {k["frac_of_hbm"]:.2f}')
except Exception as e:
    print("stress: no result", e)
PY
timeout 900 python tools/run_reference_driver.py --run --iters 20 --out gpurun_out/r2_driver > gpurun_out/r2_driver.log 2>&1; echo "driver rc=$?"
tail -5 gpurun_out/r2_driver.log
