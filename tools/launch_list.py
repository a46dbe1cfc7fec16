"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the total time)."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ck, cv, cu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= cv:
        continue
    name = re.sub(r"\(.*", "", r[ck]).replace("void ", "").replace("ghb::<unnamed>::", "").replace("unnamed>::", "").strip()
    v = float(r[cv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(r[cu], 1e-6)
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v[1] for v in tot.values())
print("| kernel | launches | total ms | share |")
print("|---|---|---|---|")
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {ms:.3f} | {100 * ms / total:.1f}% |")
