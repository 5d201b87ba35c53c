#!/bin/bash
# A/B of one environment switch on the default workload, same box: scripts/ab_env.sh VAR val1 val2 ...
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --no-cpu-baseline --e2e-loops 1 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - $var $v <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/ab_{sys.argv[2]}.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print(sys.argv[1], "=", sys.argv[2], "ms/step %.4f" % d["ms_per_step"], {n: round(v["ms_per_step"], 4) for n, v in k.items()})
PY
done
