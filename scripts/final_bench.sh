#!/bin/bash
# Round-2 bench lines of every workload of SURVEY.md section 8(d) (one B200): gpurun_out/r2c_final_<workload>.json
mkdir -p gpurun_out
python bench.py > gpurun_out/r2c_final_default.json 2> gpurun_out/r2c_final_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_final_reference.json 2>> gpurun_out/r2c_final_default.err
for wl in c3_exphander60_v4 c3_exphander60_v0 c3_exphander40_v8 c3_exphander20_v8 c3_dense c3_exphander60_v8_ddpm300 c2_dense144 c2_dense144_graphed c4_breakingbad64 c5_train144; do
  python bench.py --workload $wl --steps 120 > gpurun_out/r2c_final_$wl.json 2> gpurun_out/r2c_final_$wl.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c_final_*.json")):
    try:
        d = json.loads(open(f).read())
    except Exception as e:
        print(f, "ERR", open(f.replace(".json", ".err")).read()[-300:]); continue
    print(f.split("r2c_final_")[1][:-5].ljust(28), "value %.1f %s" % (d["value"], d["unit"]), "ms/step %.4f" % d["ms_per_step"], "e2e", d.get("e2e") and round(d["e2e"]["value"], 1),
          "parity", d.get("parity_rel_err"), "cpu", d.get("cpu_baseline", {}).get("value"), "roof", (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
PY
