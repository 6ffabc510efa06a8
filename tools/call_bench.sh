cd $GRAFT_REPO_ROOT
T=${1:-b}
shift
python bench.py "$@" > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
echo "rc=$?"
tail -c 600 gpurun_out/${T}_bench.err
python - <<PY
import json
for l in open("gpurun_out/${T}_bench.log"):
    if l.startswith("{"):
        d=json.loads(l)
        print({k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, "e2e", d.get("e2e",{}).get("value"), "cpu", d.get("cpu_baseline",{}).get("value"))
        print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["traffic"])
        print("measured", d.get("measured")); print("limiters", d.get("limiters")); print("parity", d.get("multi_gpu_parity"))
        print("c5", json.dumps(d.get("c5"))[:1500]); print("per_scene", {k:(round(v["value"]),round(v["ms_per_step"],2)) for k,v in d.get("per_scene",{}).items()})
PY
