"""profiles/r2b_stall_summary.txt: warp-stall samples of the `ncu --page source --csv` exports of scripts/ncu_round2b.sh, per
kernel: totals per stall reason (all instructions, and the arithmetic instructions alone, i.e. without the mbarrier
polling loops of the role warps), samples per opcode, and the hottest instructions.
    python scripts/ncu_stalls.py gpurun_out/r2b_source_*.csv > profiles/r2b_stall_summary.txt"""
import collections
import csv
import sys

csv.field_size_limit(10 ** 9)
POLL = {"BRA", "NOP", "SYNCS", "EXIT", "BSSY", "BSYNC", "WARPSYNC", "ELECT", "UTCHMMA", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG"}


def opcode(src):
    t = src.split()
    if not t:
        return "?"
    return (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]


for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr = rows[1]
    ci = {c: i for i, c in enumerate(hdr)}
    stall = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    seen, data = set(), []
    for r in rows[2:]:
        if len(r) < len(hdr) or r[ci["Address"]] in seen:
            continue
        seen.add(r[ci["Address"]])
        try:
            n = int(r[ci["# Samples"]] or 0)
        except ValueError:
            continue
        data.append((n, r))
    tot_all, tot_math, by_op = collections.Counter(), collections.Counter(), collections.Counter()
    for n, r in data:
        op = opcode(r[ci["Source"]])
        by_op[op] += n
        for c in stall:
            v = int(r[ci[c]] or 0)
            tot_all[c[6:]] += v
            if op not in POLL:
                tot_math[c[6:]] += v
    total = sum(n for n, _ in data)
    print(f"== {f.split('/')[-1]}: {len(data)} instructions, {total} warp samples")
    print("   stall reasons, all instructions:        ", ", ".join(f"{k} {v}" for k, v in tot_all.most_common(9)))
    print("   stall reasons, without polling / issue:  ", ", ".join(f"{k} {v}" for k, v in tot_math.most_common(9)))
    print("   samples per opcode:                      ", ", ".join(f"{k} {v}" for k, v in by_op.most_common(14)))
    data.sort(key=lambda x: -x[0])
    for n, r in data[:14]:
        st = sorted(((c[6:], int(r[ci[c]] or 0)) for c in stall), key=lambda x: -x[1])[:2]
        print(f"   {n:5d}  {r[ci['Source']][:60]:60s} {st}")
    print()
