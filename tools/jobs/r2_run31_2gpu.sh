# round 2, GPU job 31 (2 GPUs): final build -- NCCL equivalence test, one-graph weak step with render, strong scaling
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_tests_dist31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_dist31.log
grep -E "passed|failed|skipped|^FAILED|^ERROR|2-rank|^E  " gpurun_out/r2_tests_dist31.log | tail
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 300 $TR bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2_2gpu31_weak.json 2> gpurun_out/r2_2gpu31_weak.err; echo "weak rc=$?"
timeout 240 $TR bench.py --gpus 2 --no-cpu-baseline --no-render --scaling strong > gpurun_out/r2_2gpu31_strong.json 2> gpurun_out/r2_2gpu31_strong.err; echo "strong rc=$?"
timeout 240 python bench.py --no-cpu-baseline --no-render > gpurun_out/r2_1gpu31_ref.json 2>/dev/null
python - <<'PY'
import json
for v in ("2gpu31_weak", "2gpu31_strong", "1gpu31_ref"):
    try:
        d = json.loads(open(f"gpurun_out/r2_{v}.json").read().strip().splitlines()[-1])
        r = d.get("render") or {}
        print(v, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3), "graph", d["config"].get("cuda_graph"),
              "sync", d["replicas_in_sync"], "scaling", d["scaling"], "render ms/frame", r.get("ms_per_frame"), r.get("error"))
    except Exception as e:
        print(v, "no result:", e)
PY
tail -3 gpurun_out/r2_2gpu31_weak.err
