cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_prepare.py -x -q 2>&1 | tail -4
python tools/prep_bench.py 1000000 4 --no-host 2>&1 | tail -4
python tools/prep_bench.py 10000000 2 --no-host 2>&1 | tail -2
