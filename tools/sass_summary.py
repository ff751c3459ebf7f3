"""cuobjdump -sass of libradnerf_b200.so -> per-kernel instruction totals and the mnemonics that identify the Blackwell paths
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, ...).  Runs without a GPU.

    python tools/sass_summary.py [profiles/rNN_sass_summary.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "rad-nerf_b200", "libradnerf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "HMMA", "SYNCS", "RED.", "ATOM", "LDG", "LDS", "STS", "STG", "BAR", "SHFL", "MATCH",
        "MUFU", "UTCATOMSWS", "ERRBAR", "NANOSLEEP", "MEMBAR", "FENCE", "CCTL")
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+([A-Z0-9_.]+)")
cur, stats = None, collections.OrderedDict()
for line in sass.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    m = pat.match(line) if cur else None
    if m:
        op = m.group(1)
        stats[cur]["_total"] += 1
        for key in KEYS:
            if op.startswith(key):
                stats[cur][key] += 1
out = ["SASS summary of rad-nerf_b200/libradnerf_b200.so (cuobjdump -sass, sm_100a; tools/sass_summary.py)",
       "per kernel / device function: total instructions and counts of the mnemonics that identify the Blackwell paths:",
       "  UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit -> mbarrier, LDTM = tcgen05.ld (TMEM -> registers), UBLKCP = cp.async.bulk (TMA bulk copy),",
       "  UTCATOMSWS = tcgen05.alloc/dealloc, HMMA = mma.sync, SYNCS = mbarrier ops, RED/ATOM = reductions / atomics, MATCH = match.any, NANOSLEEP = spin back-off", "",
       "%-100s %7s %s" % ("function (demangled prefix)", "insts", "mnemonics")]
for k, c in stats.items():
    if not c["_total"]:
        continue
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
    name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0][:100]
    out.append("%-100s %7d %s" % (name, c["_total"], " ".join("%s=%d" % (a, b) for a, b in c.items() if a != "_total" and b)))
text = "\n".join(out) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
else:
    print(text)
