cd $GRAFT_REPO_ROOT
P=r2c
bash tools/call_bench.sh n1d
python tools/prof_step.py --passes 2 --counters > gpurun_out/${P}_plain.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-scenes > gpurun_out/${P}_bench_short.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-scenes > gpurun_out/${P}_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_prepare_launches.csv python tools/prep_bench.py 1000000 1 --no-host > gpurun_out/${P}_ncu_prepare.log 2>&1
for W in soup1m soup10m spheres100k; do bash tools/call_traffic.sh $P $W 5 > /dev/null 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 5 -c 3 -o gpurun_out/${P}_trace -f python tools/prof_step.py --passes 1 > gpurun_out/${P}_ncu_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 5 -c 3 -o gpurun_out/${P}_shade -f python tools/prof_step.py --passes 1 > gpurun_out/${P}_ncu_shade.log 2>&1
ls -la gpurun_out/ | grep r2c
cat gpurun_out/${P}_plain.log | tail -12
