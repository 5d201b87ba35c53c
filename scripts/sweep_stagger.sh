#!/bin/bash
# A/B of the persistent hidden-layer kernel's stream stagger (DA_HIDDEN_STAGGER_NS): ms per step and per-kernel times
for ns in ${@:-0 1000 4000 8000 12000}; do
  DA_HIDDEN_STAGGER_NS=$ns python bench.py --no-cpu-baseline --e2e-loops 1 > gpurun_out/sw_$ns.json 2> gpurun_out/sw_$ns.err
  python - $ns <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/sw_{sys.argv[1]}.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("stagger_ns", sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "dense_hidden %.4f" % k["attn_dense_hidden"]["ms_per_step"], "parity", d.get("parity_rel_err"))
PY
done
