"""profiles/r2_sass_summary.txt: per kernel of the shipped library, the SASS mnemonics that prove the Blackwell-native paths
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld / st -> LDTM / STTM, TMA -> UTMALDG / UTMASTG / UBLKCP; HMMA would be
the legacy mma.sync path).    python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import hashlib
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "diffassemble_b200" / "lib" / "libdiffassemble_b200.so"
PAT = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*)")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU", "HMMA", "HGMMA", "LDL", "STL"]
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
out, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("void ", "")
        cur = re.sub(r"\(.*", "", cur)
        k = 2
        base = cur
        while cur in out:
            cur = f"{base} #{k}"; k += 1
        out[cur] = collections.Counter()
        continue
    m = PAT.match(line)
    if m and cur:
        out[cur][m.group(1)] += 1
        out[cur]["_total"] += 1
print(f"library: {LIB.relative_to(ROOT)}  sha256 {hashlib.sha256(LIB.read_bytes()).hexdigest()[:16]}  (cuobjdump -sass, sm_100a)")
print(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{k:>7s}" for k in KEYS))
tot = collections.Counter()
for name, c in out.items():
    print(f"{name[:70]:70s} {c['_total']:7d} " + " ".join(f"{c[k]:7d}" for k in KEYS))
    tot.update(c)
print(f"{'TOTAL':70s} {tot['_total']:7d} " + " ".join(f"{tot[k]:7d}" for k in KEYS))
print("\nUTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,")
print("UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA / HGMMA (legacy tensor paths) must be 0; LDL / STL = spills.")
