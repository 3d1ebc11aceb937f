# A/B of differently built libraries on the same box: bash tools/ab_libs.sh "<lib or -> ..." "<workload[:extra flags]> ..."
for lib in $1; do
  if [ "$lib" = "-" ]; then unset MDR_LIB_PATH; else export MDR_LIB_PATH=$PWD/$lib; fi
  for spec in $2; do
    w=${spec%%:*}; extra=""; [ "$spec" != "$w" ] && extra=${spec#*:}
    timeout 200 python bench.py --workload $w --steps 300 --warmup 30 --no-cpu-baseline $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', '$spec', round(d['ms_per_step']*1e3,2), 'us/step, frac', round(d['roofline']['frac'],3))"
  done
done
unset MDR_LIB_PATH
