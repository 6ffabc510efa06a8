cd $GRAFT_REPO_ROOT
N=$1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/n${N}_bench.log 2> gpurun_out/n${N}_bench.err
echo "rc=$?"
tail -c 1500 gpurun_out/n${N}_bench.err
python - <<PY
import json
for l in open("gpurun_out/n${N}_bench.log"):
    if l.startswith("{"):
        d=json.loads(l)
        print({k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, "e2e", d.get("e2e",{}).get("value"), d.get("e2e",{}).get("ms_per_step"))
        print("measured", d.get("measured")); print("parity", d.get("multi_gpu_parity"))
        print("c5", json.dumps(d.get("c5"))[:1800])
PY
