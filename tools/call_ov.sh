cd $GRAFT_REPO_ROOT
rm -f gpurun_out/ov.log
for lib in librtcore_b200.so _variants/librtcore_b200_sh128.so; do
for b in 7 6 5; do
  echo "== $lib trace blocks $b" >> gpurun_out/ov.log
  RTC_TRACE_BLOCKS_PER_SM=$b RTC_B200_LIB=$PWD/raytracercore_b200/$lib python tools/prof_step.py --waves 2 --passes 4 --spp 8 2>&1 | grep "wall" >> gpurun_out/ov.log
done
done
cat gpurun_out/ov.log
