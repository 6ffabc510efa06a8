cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/c1_tests.log
tail -6 gpurun_out/c1_tests.log
bash tools/call_bench.sh n1c
