#!/usr/bin/env python
"""cuobjdump -sass bow_b200/libbowgpu.so | python scripts/sass_tma.py > profiles/rN_sass_tma.txt
Per kernel: the TMA / mbarrier mnemonics and the main instructions of the hot loops."""
import collections
import re
import subprocess
import sys

cur = None
cnt = collections.OrderedDict()
pat = re.compile(r"\b(UTMALDG[.\w]*|UBLKCP[.\w]*|UTMAPF[.\w]*|SYNCS[.\w]*|UTMACCTL[.\w]*|LDS\.128|BAR\.SYNC[.\w]*|SHFL[.\w]*|MATCH[.\w]*"
                 r"|ATOM[.\w]*|RED[.\w]*|DADD|DMUL|DSETP[.\w]*|I2F\.F64[.\w]*)")
for line in sys.stdin:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur:
        m = pat.search(line)
        if m:
            cnt[cur][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt.keys()), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonics per kernel of bow_b200/libbowgpu.so (cuobjdump -sass, sm_100a): TMA (UTMALDG = cp.async.bulk.tensor, UBLKCP =")
print("# cp.async.bulk, UTMAPF = tensor prefetch), mbarrier (SYNCS.*), and the shared-memory / shuffle / FP64 instructions of the hot")
print("# loops.  No UTC*MMA / TMEM anywhere: nothing on this path is a contraction.")
tot = collections.Counter()
for (k, c), name in zip(cnt.items(), names):
    if not c:
        continue
    tma = {m: n for m, n in c.items() if m.startswith(("UTMA", "UBLKCP", "SYNCS"))}
    rest = {m: n for m, n in c.items() if m not in tma}
    for m, n in c.items():
        tot[m] += n
    print(f"{name[:170]}\n    TMA/mbarrier: {dict(sorted(tma.items()))}\n    other: {dict(sorted(rest.items()))}")
print("TOTAL", dict(sorted(tot.items())))
