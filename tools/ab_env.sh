# A/B of environment knobs on the same box and library: bash tools/ab_env.sh "<VAR=VALUE or -> ..." "<workload[:extra flags]> ..."
for kv in $1; do
  for spec in $2; do
    w=${spec%%:*}; extra=""; [ "$spec" != "$w" ] && extra=${spec#*:}
    ( [ "$kv" != "-" ] && export "$kv"
      timeout 200 python bench.py --workload $w --steps 300 --warmup 30 --no-cpu-baseline $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$kv', '$spec', round(d['ms_per_step']*1e3,2), 'us/step, frac', round(d['roofline']['frac'],3), 'e2e %.3g' % d['e2e']['value'])" )
  done
done
