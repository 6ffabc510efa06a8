cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_baseline_parity.py -m gpu -q -s -k shading 2>&1 | grep -E "shading\[|Error|passed|failed|^E  " > gpurun_out/c5_tests.log
cat gpurun_out/c5_tests.log
