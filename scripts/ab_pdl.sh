#!/bin/bash
# A/B of programmatic dependent launch (DA_NO_PDL=1 = plain stream order) on three sizes, same box
for args in "--global-batch 32" "--global-batch 4" "--workload c2_dense144_graphed --steps 120"; do
  for v in 1 0; do
    DA_NO_PDL=$v python bench.py $args --no-cpu-baseline --e2e-loops 1 > gpurun_out/abp.json 2> gpurun_out/abp.err
    python - "$args" $v <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/abp.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "DA_NO_PDL=" + sys.argv[2], "ms/step %.4f" % d["ms_per_step"], "value %.0f" % d["value"], "parity", d.get("parity_rel_err"))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "ERR", e, open("gpurun_out/abp.err").read()[-400:])
PY
  done
done
