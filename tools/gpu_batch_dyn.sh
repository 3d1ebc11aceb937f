for pw in 0 100; do for k in 1 2; do timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --prewarm-ms $pw 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('prewarm $pw', round(d['ms_per_step']*1e3,2), 'us/step, frac', round(d['roofline']['frac'],3), 'e2e %.3g' % d['e2e']['value'], d['clocks'])"; done; done
timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default', d['steps'], d['warmup'], round(d['ms_per_step']*1e3,2), 'us/step, frac', round(d['roofline']['frac'],3), 'e2e %.3g' % d['e2e']['value'])"
