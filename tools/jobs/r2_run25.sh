# round 2, GPU job 25: 32-row tiles of the fused noise-MLP forward on render chunks (test + render launch list), ncu
# --set full of every kernel of one training step (source-level sector / stall analysis, tools/ncu_source_top.py)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_25.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_25.log | tail -10
K='absmax|adam|blend|blur_loss|composite|enc_bwd|fv_post|gather_fwd|head_grads|mlp_tc|motion|pack_|ray_|small_mlp|sample_|stratified|rbk'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip 144 -c 56 -o gpurun_out/r2_ncu_allstep python bench.py --steps 1 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_allstep.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/ncu_render_target.py > gpurun_out/r2_render_plain.log 2>&1; cat gpurun_out/r2_render_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_render.csv python tools/ncu_render_target.py > gpurun_out/r2_render_ncu.log 2>&1; echo "ncu render rc=$?"
grep -c small_mlp gpurun_out/r2_launches_render.csv
ls -la gpurun_out/r2_ncu_allstep.ncu-rep
