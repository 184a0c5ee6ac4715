"""Per-kernel SASS evidence for the committed library: counts of the Blackwell tensor-core / TMA
mnemonics in `cuobjdump -sass libasyncrl_b200.so` (UTCHMMA / UTCIMMA = tcgen05.mma kind::f16 / i8,
LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTMALDG = tensor-map TMA, SYNCS = mbarrier ops).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "async-rl-tensorflow_b200", "libasyncrl_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "ELECT", "LDG", "STG", "HMMA", "FFMA"]
cur, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        rows[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        rows[cur]["_all"] += 1
        for p in pats:
            if op.startswith(p):
                rows[cur][p] += 1
print("# cuobjdump -sass %s (sm_100a): instruction counts per kernel" % os.path.relpath(so, ROOT))
print("%-78s %6s " % ("kernel", "SASS") + " ".join("%7s" % p for p in pats))
tot = collections.Counter()
for k, c in rows.items():
    name = re.sub(r"\(.*", "", k.replace("(anonymous namespace)::", "")).replace("arl::", "")[:78]
    print("%-78s %6d " % (name, c["_all"]) + " ".join("%7d" % c[p] for p in pats))
    tot.update(c)
print("%-78s %6d " % ("TOTAL", tot["_all"]) + " ".join("%7d" % tot[p] for p in pats))
