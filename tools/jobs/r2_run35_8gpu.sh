# round 2, GPU job 35 (8 GPUs): the final build, weak scaling with the NCCL all-reduces inside one CUDA graph (+ render leg)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 400 $TR bench.py --gpus 8 --no-cpu-baseline --render-frames 8 > gpurun_out/r2_8gpu35_weak.json 2> gpurun_out/r2_8gpu35_weak.err; echo "weak rc=$?"
timeout 200 python bench.py --no-cpu-baseline --no-render > gpurun_out/r2_8box35_1gpu.json 2>/dev/null; echo "1gpu rc=$?"
python - <<'PY'
import json
for v in ("8gpu35_weak", "8box35_1gpu"):
    try:
        d = json.loads(open(f"gpurun_out/r2_{v}.json").read().strip().splitlines()[-1])
        r = d.get("render") or {}
        print(v, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3),
              "sync", d["replicas_in_sync"], "render ms/frame", r.get("ms_per_frame"), r.get("error"))
    except Exception as e:
        print(v, "no result:", e)
PY
tail -3 gpurun_out/r2_8gpu35_weak.err
