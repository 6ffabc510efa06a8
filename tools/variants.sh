for v in "" _prmt; do
  if [ -z "$v" ]; then unset RTC_B200_LIB; else export RTC_B200_LIB=$PWD/raytracercore_b200/librtcore_b200$v.so; fi
  echo "== variant '$v'"; python tools/prof_step.py --passes 2 2>&1 | grep -E "trace-only|trace  "
done
