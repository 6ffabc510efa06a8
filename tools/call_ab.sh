cd $GRAFT_REPO_ROOT
L=gpurun_out/$1.log; shift
rm -f $L
bash tools/ab.sh $L "$@"
cat $L
