#!/usr/bin/env python
"""SASS opcode histogram of the hot kernels of libfqss_sm100.so (proof that the Blackwell path is what ships):
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA loads), UTCBAR (tcgen05.commit), SYNCS (mbarrier), ...
usage: python profiles/sass_hist.py [lib.so] > profiles/sass_r03.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "fqss_b200/_lib/libfqss_sm100.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, per = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        per[fn][m.group(1)] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMAPF", "UTCBAR", "SYNCS", "LDGSTS", "REDUX", "ATOMG", "RED", "FFMA2", "MUFU", "F2I", "I2F"]
tot = collections.Counter()
print("# SASS opcode counts per kernel (cuobjdump -sass %s); key columns, then the five most frequent opcodes" % lib)
print("%-78s %6s %s" % ("kernel", "instr", " ".join("%7s" % k for k in KEY)))
for (fn, c), name in zip(per.items(), demangled):
    n = sum(c.values())
    if n == 0:
        continue
    tot.update(c)
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    top = ", ".join("%s %d" % kv for kv in c.most_common(5))
    print("%-78s %6d %s   | %s" % (short[:78], n, " ".join("%7d" % c.get(k, 0) for k in KEY), top))
print("%-78s %6d %s" % ("TOTAL", sum(tot.values()), " ".join("%7d" % tot.get(k, 0) for k in KEY)))
