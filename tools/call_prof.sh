cd $GRAFT_REPO_ROOT
P=${1:-r2a}
python tools/prof_step.py --passes 2 --counters > gpurun_out/${P}_plain.log 2>&1 || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-scenes > gpurun_out/${P}_bench_short.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-scenes > gpurun_out/${P}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 5 -c 3 -o gpurun_out/${P}_trace -f python tools/prof_step.py --passes 1 > gpurun_out/${P}_ncu_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 5 -c 3 -o gpurun_out/${P}_shade -f python tools/prof_step.py --passes 1 > gpurun_out/${P}_ncu_shade.log 2>&1
cat gpurun_out/${P}_plain.log
ls -la gpurun_out/
