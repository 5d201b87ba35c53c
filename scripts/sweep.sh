for b in 4 8 16 32; do python bench.py --steps 200 --no-cpu-baseline --e2e-loops 0 --global-batch $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('B=$b', round(d['ms_per_step'],4), 'ms/step', round(d['value'],0), {k:v['ms_per_launch'] for k,v in d['roofline']['kernels'].items()})
"; done
for wl in c3_exphander20_v8 c3_exphander40_v8 c3_exphander60_v4 c3_exphander60_v0 c3_dense; do python bench.py --steps 100 --no-cpu-baseline --e2e-loops 0 --workload $wl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl', round(d['ms_per_step'],4), 'ms/step', round(d['value'],0), {k:v['ms_per_launch'] for k,v in d['roofline']['kernels'].items() if 'attn' in k})
"; done
