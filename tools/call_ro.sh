cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_render.py -q -x -k "reordering" 2>&1 | tail -5
for wl in soup10m soup1m spheres100k; do
for m in 0 1 2; do
  echo "== $wl reorder $m" >> gpurun_out/ro.log
  python tools/prof_step.py --workload $wl --passes 3 --spp 2 --reorder $m 2>&1 | grep -v "^  [ra]" | tail -5 >> gpurun_out/ro.log
done
done
for m in 0 1 2; do
  echo "== soup10m waves 2 reorder $m" >> gpurun_out/ro.log
  python tools/prof_step.py --workload soup10m --passes 3 --spp 4 --waves 2 --reorder $m 2>&1 | grep "wall" >> gpurun_out/ro.log
done
cat gpurun_out/ro.log
