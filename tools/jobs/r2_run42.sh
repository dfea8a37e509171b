# round 2, GPU job 42: the tree as committed (main library): smoke + full GPU suite
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_42.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke_42.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests_42.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_42.log | tail -6
