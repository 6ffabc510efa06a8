# Collapse-width experiment: build the variant libraries first, here in the container --
#   python -c "from raytracercore_b200 import build as B; [B.build_variant('w%d' % w, ['-DRTC_COLLAPSE_WIDTH=%d' % w]) for w in (6, 4)]"
# (they travel to the GPU box with the snapshot), then gpurun this script.
cd $GRAFT_REPO_ROOT
for lib in librtcore_b200.so librtcore_b200_w6.so librtcore_b200_w4.so; do
  echo "== $lib" >> gpurun_out/tq.log
  RTC_B200_VERBOSE=1 RTC_B200_LIB=$PWD/raytracercore_b200/$lib python tools/prof_step.py --passes 2 --spp 2 --counters 2>&1 | grep -v "^  [rca]" | grep -v "flatten:" | tail -5 >> gpurun_out/tq.log
  RTC_B200_LIB=$PWD/raytracercore_b200/$lib python tools/prof_step.py --passes 3 --spp 4 2>&1 | grep -v "^  [rca]" | tail -3 >> gpurun_out/tq.log
done
cat gpurun_out/tq.log
