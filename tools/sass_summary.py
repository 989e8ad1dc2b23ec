#!/usr/bin/env python
"""profiles/r02_sass_summary.txt: per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAREDG/UBLKCP, legacy
mma.sync -> HMMA, cp.async -> LDGSTS), from `cuobjdump -sass` of the shipped library."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "facialmmt_b200/libfacialmmt_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "MUFU"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("fmmt::", "")
        cur = re.sub(r"^void ", "", cur)
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# SASS mnemonic counts per kernel, cuobjdump -sass {lib}")
print(f"{'kernel':70s} {'instrs':>7s} " + " ".join(f"{k:>8s}" for k in KEYS))
for k, c in counts.items():
    if total[k] == 0:
        continue
    print(f"{k[:70]:70s} {total[k]:7d} " + " ".join(f"{c.get(x, 0):8d}" for x in KEYS))
