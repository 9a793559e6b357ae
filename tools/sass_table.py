"""SASS op counts per kernel of libasr_b200.so (the Blackwell-native evidence: tcgen05 = UTC*MMA, TMEM = LDTM/STTM, TMA = UTMALDG/UTMASTG/UBLKCP):
   python tools/sass_table.py > profiles/r02_sass_opcounts.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "asr_streaming_b200", "libasr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
OPS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "HMMA", "LDGSTS", "LDSM", "SYNCS", "FFMA2", "MUFU"]
counts, order, cur = {}, [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"asr::\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*\)$", "", cur)
        cur = re.sub(r"^void ", "", cur)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?PT?\d*\s+)?([A-Z][A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["_total"] += 1
    for o in OPS:
        if op == o or op.startswith(o + "."):
            if o == "UTCHMMA" and ".2CTA" in op:
                continue
            if o == "UTCHMMA.2CTA" and ".2CTA" not in op:
                continue
            counts[cur][o] += 1
            break
    else:
        if op.startswith("UTCHMMA"):
            counts[cur]["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
print("# SASS op counts per kernel — `cuobjdump -sass asr_streaming_b200/libasr_b200.so` (sm_100a), round 2\n")
print("tcgen05.mma = `UTCHMMA` (`.2CTA` = cta_group::2), tcgen05.commit = `UTCBAR`, tcgen05.ld / st = `LDTM` / `STTM`, TMA loads / stores =")
print("`UTMALDG` / `UTMASTG`, bulk copies = `UBLKCP`, `mma.sync` = `HMMA`, `ldmatrix` = `LDSM`, mbarrier ops = `SYNCS`, packed fp32 = `FFMA2`.\n")
cols = [o for o in OPS if any(counts[k][o] for k in order)]
print("| kernel | instructions | " + " | ".join(f"`{c}`" for c in cols) + " |")
print("|---|---|" + "---|" * len(cols))
for k in order:
    c = counts[k]
    if not any(c[o] for o in ("UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "HMMA", "UBLKCP")) and c["_total"] < 400:
        continue
    name = k if len(k) < 110 else k[:107] + "..."
    print(f"| `{name}` | {c['_total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in cols) + " |")
