# round 2, GPU job 13: is the cost of the stash stores their instructions or their DRAM traffic?
# main build vs the same store instructions into an L2-resident 64 KB window per CTA vs no stores at all
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LUSH_BENCH_ALLOW_NONFINITE=1
B="python bench.py --no-render --no-cpu-baseline --steps 10"
for i in 1 2; do
  timeout 300 $B > gpurun_out/r2_exp13_main$i.json 2>/dev/null
  LUSH_LIB_PATH=$PWD/lib_exp_stashlocal.so timeout 300 $B > gpurun_out/r2_exp13_local$i.json 2>/dev/null
  LUSH_LIB_PATH=$PWD/lib_exp_nostash.so timeout 300 $B > gpurun_out/r2_exp13_nostash$i.json 2>/dev/null
done
python - <<'PY'
import json
for v in ("main1", "local1", "nostash1", "main2", "local2", "nostash2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_exp13_{v}.json").read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              [round(x["ms_per_step"], 3) for x in pk.values()], d["clocks"])
    except Exception as e:
        print(v, "no result:", e)
PY
