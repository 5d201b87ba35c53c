"""profiles/r2_ncu_traffic.json from the ncu full-capture summary (scripts/ncu_round2.sh): per launch SITE of one denoising
step, dram__bytes_read.sum + dram__bytes_write.sum per launch.  bench.py reads the table into roofline.traffic.
    python scripts/ncu_traffic.py gpurun_out/r2b_full_summary.csv c3_exphander60_v8 profiles/r2b_ncu_full_step_summary.csv"""
import csv
import json
import re
import sys
from pathlib import Path

src, workload, committed = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units = rows[hdr_i], rows[hdr_i + 1]
data = rows[hdr_i + 2:]
ki = hdr.index("Kernel Name")
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v, u = r[col[name]].replace(",", ""), units[col[name]]
    x = float(v) if v not in ("", "n/a") else 0.0
    return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)


# one step = the launches between two consecutive prologue kernels
starts = [i for i, r in enumerate(data) if "prologue" in r[ki]]
assert len(starts) >= 2, "need two prologue launches in the capture"
step = data[starts[0]:starts[1]]
folded = any("attn_fold" in r[ki] or "head_fold" in r[ki] for r in step)   # weight-folded path (csrc/fold.cu)
n_gather = sum("gather_extra" in r[ki] for r in step)
site_of, seen = [], {}
for r in step:
    n = r[ki]
    base = re.sub(r"\(.*", "", n)
    k = seen[base] = seen.get(base, 0) + 1
    is_gemm = "linear_umma_kernel" in n
    bn = int(re.search(r"linear_umma_kernel<(?:\(int\))?(\d+)", n).group(1)) if is_gemm else 0
    mode = int(re.search(r"linear_umma_kernel<(?:\(int\))?\d+, (?:\(int\))?(\d+)", n).group(1)) if is_gemm else 0
    if "prologue" in n: site = "prologue"
    elif is_gemm and folded:   # GEMMs of a folded step, in launch order: first projection, two hidden layers, [Q | K | V'] of the last
        seen["_g"] = seen.get("_g", 0) + 1
        site = "qkvs_gemm_last" if mode == 3 else ("qkvs_gemm_first" if seen["_g"] == 1 else "qkvs_gemm_mid")
    elif is_gemm and bn == 128: site = "mlp2_gemm" if k == 1 else "qkvs_gemm_first"
    elif is_gemm and bn == 256: site = "qkvs_gemm_mid" if k <= 2 else "qkvs_gemm_last"
    elif is_gemm and bn == 32: site = "head_gemm"
    elif "attn_hidden_persist" in n or "attn_dense_kernel<32" in n or "attn_dense_kernel<(int)32" in n: site = "attn_dense_hidden"
    elif "attn_fold" in n or "attn_dense_fold" in n or "attn_dense_kernel<144" in n or "attn_dense_kernel<(int)144" in n: site = "attn_dense_last"
    elif "vrow32" in n: site = "attn_hidden"
    elif "gather_extra" in n: site = "pack_last" if k == n_gather else "pack_hidden"
    elif "head_f" in n: site = "head_final"
    else: site = "other"
    site_of.append(site)
acc = {}
for r, site in zip(step, site_of):
    e = acc.setdefault(site, {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
    e["launches"] += 1
    e["dram_bytes"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    e["us"] += val(r, "gpu__time_duration.sum") if units[col["gpu__time_duration.sum"]] in ("us", "usecond") else val(r, "gpu__time_duration.sum") / 1e3
out_p = Path("profiles/r2_ncu_traffic.json")
table = json.loads(out_p.read_text()) if out_p.exists() else {}
table["_source"] = committed
table["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch of each launch site of ONE denoising step, from "
                     "`ncu --set full --clock-control none` of bench.py (scripts/ncu_round2.sh); written by scripts/ncu_traffic.py")
table[workload] = {s: {"dram_bytes_per_launch": int(e["dram_bytes"] / e["launches"]), "launches_per_step": e["launches"],
                       "ncu_us_per_launch": round(e["us"] / e["launches"], 1)} for s, e in acc.items()}
out_p.write_text(json.dumps(table, indent=1))
tot = sum(e["dram_bytes"] for e in acc.values())
print(f"{len(step)} launches per step; DRAM traffic per step {tot / 1e9:.3f} GB; ncu time per step {sum(e['us'] for e in acc.values()):.0f} us")
for s, e in acc.items():
    print(f"  {s:20s} x{e['launches']}  {e['dram_bytes'] / e['launches'] / 1e6:8.1f} MB / launch  {e['us'] / e['launches']:8.1f} us / launch")
