for b in 4 32; do python bench.py --steps 200 --no-cpu-baseline --e2e-loops 0 --global-batch $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('B=$b', round(d['ms_per_step'],4), 'ms/step', round(d['value'],0), d['gpu_launches']/d['steps'], {k:v['ms_per_launch'] for k,v in d['roofline']['kernels'].items()})
"; done
