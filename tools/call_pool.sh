cd $GRAFT_REPO_ROOT
rm -f gpurun_out/pool.log
for mp in 4194304 8388608 16777216 33554432; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-other-scenes --no-c5 --max-paths $mp > gpurun_out/pool_$mp.log 2>&1
  python - <<PY >> gpurun_out/pool.log
import json
for l in open("gpurun_out/pool_$mp.log"):
    if l.startswith("{"):
        d=json.loads(l); print($mp, round(d["value"],1), round(d["ms_per_step"],2), d["gpu_launches"])
PY
done
cat gpurun_out/pool.log
