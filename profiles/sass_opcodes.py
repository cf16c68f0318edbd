#!/usr/bin/env python
"""Opcode histogram of the SASS of every kernel in libb200reg.so (cuobjdump -sass; no GPU needed):
    python profiles/sass_opcodes.py [kernel-name-substring ...] > profiles/r02/sass_opcodes.txt
Shows, per kernel, the instruction count, the top opcodes and the Blackwell-specific ones the profiling guide names
(UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, UTMALDG/UTMASTG = tensor-map TMA, UTC*MMA / LDTM / STTM =
tcgen05, HMMA = legacy mma.sync, REDUX = warp reductions, MATCH / VOTE = warp votes)."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "fast-lio-sam-qn_b200", "csrc", "libb200reg.so")
want = sys.argv[1:]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, ops = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        ops[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        ops[kern][m.group(1)] += 1
special = ("UBLKCP", "SYNCS", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "HMMA", "REDUX", "MATCH", "VOTE", "SHFL", "ATOMS", "ATOMG", "RED",
           "DFMA", "DMUL", "DADD", "FFMA", "LDL", "STL")
for k, c in ops.items():
    if want and not any(w in k for w in want):
        continue
    n = sum(c.values())
    top = " ".join("%s:%d" % kv for kv in c.most_common(8))
    sp = " ".join("%s:%d" % (s, c[s]) for s in special if c[s])
    print("%-48s %6d instr | %s | %s" % (k[-48:], n, top, sp))
