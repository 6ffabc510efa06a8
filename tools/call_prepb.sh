cd $GRAFT_REPO_ROOT
RTC_B200_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_prepare.py -x -q 2>&1 | tail -8 > gpurun_out/prep_tests.log
tail -8 gpurun_out/prep_tests.log
python tools/prep_bench.py 1000000 4 > gpurun_out/prepb.log 2>&1
cat gpurun_out/prepb.log
python tools/prep_bench.py 10000000 2 > gpurun_out/prepb10.log 2>&1
cat gpurun_out/prepb10.log
