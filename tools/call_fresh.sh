cd $GRAFT_REPO_ROOT
for i in 1 2 3; do python tools/prep_bench.py 1000000 1 --no-host 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_prepare.py tests/test_abi.py -q -x 2>&1 | tail -3
