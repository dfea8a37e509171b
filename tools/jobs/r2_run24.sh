# round 2, GPU job 24: ncu --set full of EVERY kernel of one training step (source-level sector / stall analysis of the
# side kernels too), launch list of two rendered frames
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_24.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke_24.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lush --launch-skip 144 -c 48 -o gpurun_out/r2_ncu_allstep python bench.py --steps 1 --warmup 3 --no-render --no-cpu-baseline --graph none > gpurun_out/r2_ncu_allstep.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/ncu_render_target.py > gpurun_out/r2_render_plain.log 2>&1; cat gpurun_out/r2_render_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_render.csv python tools/ncu_render_target.py > gpurun_out/r2_render_ncu.log 2>&1; echo "ncu render rc=$?"
ls -la gpurun_out/r2_ncu_allstep.ncu-rep gpurun_out/r2_launches_render.csv
