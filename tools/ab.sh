#!/bin/bash
# A/B run of kernel variants on the GPU box: tools/ab.sh <logfile> <prof_step args...>  (variants = every .so in _variants/)
log=$1; shift
echo "== main" >> $log
python tools/prof_step.py "$@" 2>&1 | grep -v "^  [rca]" >> $log
for lib in raytracercore_b200/_variants/*.so; do
  echo "== $lib" >> $log
  RTC_B200_LIB=$PWD/$lib python tools/prof_step.py "$@" 2>&1 | grep -v "^  [rca]" >> $log
done
