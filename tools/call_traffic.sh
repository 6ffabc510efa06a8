# tools/call_traffic.sh <tag> <workload> <launches per pass = recursion + 1>: full ncu captures of every trace / shade launch of one
# pass; only the raw pages come back (the reports of three workloads exceed gpurun's 64 MiB return limit)
cd $GRAFT_REPO_ROOT
P=$1; W=$2; L=$3
python tools/prof_step.py --workload $W --passes 1 > gpurun_out/${P}_${W}_plain.log 2>&1 || exit 1
for K in trace shade; do
  ncu --set full --clock-control none -k regex:k_$K -s $L -c $L -o /tmp/${P}_${W}_$K -f python tools/prof_step.py --workload $W --passes 1 > gpurun_out/${P}_${W}_ncu_$K.log 2>&1
  ncu -i /tmp/${P}_${W}_$K.ncu-rep --page raw --csv > gpurun_out/${P}_${W}_${K}_raw.csv
done
cat gpurun_out/${P}_${W}_plain.log
