cd $GRAFT_REPO_ROOT
rm -f gpurun_out/ab2.log
bash tools/ab.sh gpurun_out/ab2.log --passes 3 --spp 4
bash tools/ab.sh gpurun_out/ab2.log --passes 3 --spp 4
grep -E "^==|trace-only" gpurun_out/ab2.log
