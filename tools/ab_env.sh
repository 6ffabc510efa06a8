#!/bin/bash
# tools/ab_env.sh <logfile> VAR "v1 v2 ..." <prof_step args...> : run the main library with VAR set to each value
log=$1; var=$2; vals=$3; shift 3
for v in $vals; do
  echo "== $var=$v" >> $log
  env $var=$v python tools/prof_step.py "$@" 2>&1 | grep -v "^  [rca]" >> $log
done
