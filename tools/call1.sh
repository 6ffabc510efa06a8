cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/c1_tests.log
tail -25 gpurun_out/c1_tests.log
