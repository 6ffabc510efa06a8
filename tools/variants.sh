for v in "" _b5 _b4 _b3 _t64b10 _t256b2; do
  if [ -z "$v" ]; then unset RTC_B200_LIB; else export RTC_B200_LIB=$PWD/raytracercore_b200/librtcore_b200$v.so; fi
  echo "== variant '$v'"; python tools/prof_step.py --passes 3 2>&1 | grep -E "trace-only"
  python tools/prof_step.py --passes 3 2>&1 | grep -E "trace-only"
done
