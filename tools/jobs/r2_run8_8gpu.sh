# round 2, GPU job 8 (8 GPUs): weak scaling with the one-graph step (NCCL inside), strong scaling of ONE global
# N_rand = 1024 batch, frames sharded for rendering; 1-GPU reference on the same box
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
unset LUSH_LIB_PATH
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_8box_1gpu.json 2>/dev/null; echo "1gpu rc=$?"
timeout 300 $TR bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r2_8gpu_weak.json 2> gpurun_out/r2_8gpu_weak.err; echo "weak rc=$?"
timeout 300 $TR bench.py --gpus 8 --no-cpu-baseline --no-render --scaling strong > gpurun_out/r2_8gpu_strong.json 2> gpurun_out/r2_8gpu_strong.err; echo "strong rc=$?"
timeout 300 $TR bench.py --gpus 8 --no-cpu-baseline --no-render --graph split > gpurun_out/r2_8gpu_weak_split.json 2> gpurun_out/r2_8gpu_weak_split.err; echo "split rc=$?"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552"
timeout 300 $TR4 bench.py --gpus 4 --no-cpu-baseline --no-render > gpurun_out/r2_4gpu_weak.json 2> gpurun_out/r2_4gpu_weak.err; echo "weak4 rc=$?"
timeout 300 $TR4 bench.py --gpus 4 --no-cpu-baseline --no-render --scaling strong > gpurun_out/r2_4gpu_strong.json 2> gpurun_out/r2_4gpu_strong.err; echo "strong4 rc=$?"
python - <<'PY'
import json
for v in ("8box_1gpu", "8gpu_weak", "8gpu_weak_split", "8gpu_strong", "4gpu_weak", "4gpu_strong"):
    try:
        d = json.loads(open(f"gpurun_out/r2_{v}.json").read().strip().splitlines()[-1])
        r = d.get("render") or {}
        print(v, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_eager"], 3), "graph", d["config"].get("cuda_graph"),
              "sync", d["replicas_in_sync"], "scaling", d["scaling"], "render ms/frame", r.get("ms_per_frame"), r.get("error"))
    except Exception as e:
        print(v, "no result:", e)
PY
tail -3 gpurun_out/r2_8gpu_weak.err
