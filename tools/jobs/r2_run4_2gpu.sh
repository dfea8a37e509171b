# round 2, GPU job 4 (2 GPUs): NCCL equivalence test, one-graph step with the collectives inside, strong scaling, render
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LUSH_LIB_PATH=$PWD/lib_exp_r2d.so
timeout 200 python tools/debug_render_stats.py > gpurun_out/r2_debug_render.log 2>&1; tail -24 gpurun_out/r2_debug_render.log
nvidia-smi -L
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_tests_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_dist.log
grep -E "passed|failed|skipped|^FAILED|^ERROR|2-rank|^E  " gpurun_out/r2_tests_dist.log | tail
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
for mode in single split none; do
  timeout 240 $TR bench.py --gpus 2 --no-cpu-baseline --graph $mode > gpurun_out/r2_2gpu_weak_$mode.json 2> gpurun_out/r2_2gpu_weak_$mode.err; echo "weak $mode rc=$?"
done
timeout 240 $TR bench.py --gpus 2 --no-cpu-baseline --no-render --scaling strong > gpurun_out/r2_2gpu_strong.json 2> gpurun_out/r2_2gpu_strong.err; echo "strong rc=$?"
timeout 240 python bench.py --no-cpu-baseline > gpurun_out/r2_1gpu_ref.json 2>/dev/null
python - <<'PY'
import json
for v in ("2gpu_weak_single", "2gpu_weak_split", "2gpu_weak_none", "2gpu_strong", "1gpu_ref"):
    try:
        d = json.loads(open(f"gpurun_out/r2_{v}.json").read().strip().splitlines()[-1])
        r = d.get("render") or {}
        print(v, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3), "graph", d["config"].get("cuda_graph"),
              "sync", d["replicas_in_sync"], "scaling", d["scaling"], "render ms/frame", r.get("ms_per_frame"), r.get("error"))
    except Exception as e:
        print(v, "no result:", e)
PY
tail -5 gpurun_out/r2_2gpu_weak_single.err
