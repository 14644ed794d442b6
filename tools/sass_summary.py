#!/usr/bin/env python
"""Per-kernel SASS instruction census of libhrp_b200.so (cuobjdump -sass): the mnemonics that prove which hardware
paths a kernel uses on sm_100a -- UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA tiled load / store), UBLKCP (bulk
copy), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (legacy mma.sync: must be 0), plus the
register count from the ELF headers.  Runs on the build host (no GPU).  Usage: tools/sass_summary.py [lib] > out.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "holistic-robot-pose-estimation_b200/libhrp_b200.so"
MNEMS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2",
         "LDG", "STG", "LDS", "STS", "RED", "ATOM", "SHFL"]

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
        cur = None

demangle = {}
names = re.findall(r"Function : (\S+)", sass)
if names:
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, out))

counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["_total"] += 1
    if ".2CTA" in op:
        counts[cur][".2CTA"] += 1
    for mn in MNEMS:
        if op == mn or op.startswith(mn + ".") or (mn == "MUFU.EX2" and op.startswith("MUFU.EX2")):
            counts[cur][mn] += 1


def short(name):
    d = demangle.get(name, name)
    d = re.sub(r"\(.*\)$", "", d)
    d = d.replace("hrp::", "").replace("(anonymous namespace)::", "")
    d = re.sub(r"^void ", "", d)
    return d[:74]


cols = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "HMMA", ".2CTA", "MUFU.EX2", "LDG", "STG", "SHFL"]
print("SASS census of %s (sm_100a), cuobjdump -sass; columns are instruction counts in the kernel body" % lib)
print("%-74s %5s %6s " % ("kernel", "regs", "instr") + " ".join("%8s" % c for c in cols))
tot = collections.Counter()
for name, c in counts.items():
    r = regs.get(name, (0, 0))[0]
    print("%-74s %5d %6d " % (short(name), r, c["_total"]) + " ".join("%8d" % c[k] for k in cols))
    tot.update(c)
print("%-74s %5s %6d " % ("ALL KERNELS", "", tot["_total"]) + " ".join("%8d" % tot[k] for k in cols))
print()
print("tcgen05 kernels (UTCHMMA > 0): %d; kernels using TMA tiled loads: %d; TMA / bulk stores: %d; legacy HMMA anywhere: %d; "
      "cta_group::2 (.2CTA) anywhere: %d" % (
          sum(1 for c in counts.values() if c["UTCHMMA"]), sum(1 for c in counts.values() if c["UTMALDG"]),
          sum(1 for c in counts.values() if c["UTMASTG"] or c["UBLKCP"]), tot["HMMA"], tot[".2CTA"]))
