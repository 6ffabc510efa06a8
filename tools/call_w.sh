cd $GRAFT_REPO_ROOT
for wl in spheres100k bounce die soup10m; do
 for w in 2 1; do
  echo "== $wl waves $w"
  python tools/prof_step.py --workload $wl --passes 6 --spp 16 --waves $w 2>&1 | grep -E "wall"
 done
done
