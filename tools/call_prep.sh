cd $GRAFT_REPO_ROOT
RTC_B200_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_prepare.py -x -q -s 2>&1 | tail -60 > gpurun_out/prep_tests.log
tail -40 gpurun_out/prep_tests.log
