# round 2, GPU job 26: inference kernel under no_grad (test, render launch list, render leg of the bench), 32-row noise-MLP
# tiles, ncu --set full of the side kernels of one training step summarised ON THE BOX (reports exceed the 64 MiB merge)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp_tc_gpu.py tests/test_model_gpu.py tests/test_api_surface_gpu.py tests/test_parity_configs_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_26.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_26.log | tail -10
timeout 300 python tools/ncu_render_target.py > gpurun_out/r2_render_plain.log 2>&1; cat gpurun_out/r2_render_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_render.csv python tools/ncu_render_target.py > gpurun_out/r2_render_ncu.log 2>&1; echo "ncu render rc=$?"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_26.json 2> gpurun_out/r2_bench_26.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_26.json").read().strip().splitlines()[-1])
print(round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), [round(x["ms_per_step"], 3) for x in d["roofline"]["per_kernel"].values()])
print(d["render"])
PY
K='absmax|adam|blend|blur_loss|composite|enc_bwd|fv_post|gather_fwd|head_grads|motion|pack_|ray_|small_mlp|sample_|stratified|rbk'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip 120 -c 44 -o /tmp/r2_ncu_side python bench.py --steps 1 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_side.log 2>&1; echo "ncu side rc=$?"
python tools/ncu_source_top.py /tmp/r2_ncu_side.ncu-rep --top 8 > gpurun_out/r2_ncu_side_source.txt 2>&1
ncu -i /tmp/r2_ncu_side.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread > gpurun_out/r2_ncu_side_raw.csv 2>/dev/null
wc -l gpurun_out/r2_ncu_side_source.txt gpurun_out/r2_ncu_side_raw.csv
