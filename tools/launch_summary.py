"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) by kernel name.
Usage: python tools/launch_summary.py LAUNCHES.csv OUT.txt ["header note"]"""
import csv
import sys
from collections import defaultdict

src, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
lines = [l for l in open(src, errors="ignore") if not l.startswith("==")]
rows = list(csv.reader(lines))
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
tot, cnt, order = defaultdict(float), defaultdict(int), 0
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= vi:
        continue
    try:
        us = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    except ValueError:
        continue
    tot[r[ki]] += us
    cnt[r[ki]] += 1
    order += 1
total = sum(tot.values())
with open(out, "w") as f:
    if note:
        f.write("# %s\n" % note)
    f.write("# per-launch times are cold-cache and serialised: compare SHARES.  %d launches; total %.1f us\n" % (order, total))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write("%10.1f us %5.1f%%  n=%5d  %s\n" % (v, 100 * v / total, cnt[k], k[:140]))
print(open(out).read()[:3000])
