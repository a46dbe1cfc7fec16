"""Executed warp-instructions and stall samples per source REGION of one kernel, from
`ncu -i rep --page source --csv --print-source cuda,sass`.  SASS rows follow the source line they belong to; lines of inlined
helpers (asm wrappers above `first`, CUDA headers) are attributed to the region of the closest preceding kernel line in SASS
ADDRESS order.  usage: python tools/ncu_regions3.py export.csv ncells first "name:lo-hi" ..."""
import csv, sys
path, ncells, first = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
regions = []
for a in sys.argv[4:]:
    name, r = a.split(":"); lo, hi = r.split("-"); regions.append((name, int(lo), int(hi)))
rows = list(csv.reader(open(path)))
inst = {}   # address -> (line or None, inst, samples)
main = True; line = None
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        main = r[1].endswith(".cu"); continue
    if r[0] == "Line No" or r[0] == "Function Name": continue
    if r[0] != "":
        try: line = int(r[0])
        except ValueError: line = None
        continue
    if len(r) > 5 and r[2].startswith("0x"):
        try:
            addr = int(r[2], 16); smp = int(r[4] or 0) if r[4] not in ("-",) else 0
        except ValueError:
            continue
        # columns: '', '', Address, Source, samples(all), samples(not issued), # Samples?, Instructions Executed ...
        inst[addr] = (line if (main and line is not None and line >= first) else None, r)
hdr = next(r for r in rows if r and r[0] == "Line No")
ci = [i for i, h in enumerate(hdr) if h == "Instructions Executed"][0]
cs = [i for i, h in enumerate(hdr) if h == "# Samples"][0]
agg = {n: [0, 0] for n, _, _ in regions}; agg["other"] = [0, 0]
cur = "other"
for addr in sorted(inst):
    ln, r = inst[addr]
    if ln is not None:
        cur = next((n for n, lo, hi in regions if lo <= ln <= hi), "other")
    try:
        agg[cur][0] += int(r[ci] or 0); agg[cur][1] += int(r[cs] or 0)
    except ValueError:
        pass
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total {ti / ncells:.0f} warp-inst/cell")
for n, v in agg.items():
    print(f"{n:14s} {v[0] / ncells:7.0f} inst/cell {100 * v[0] / ti:5.1f}%   samples {100 * v[1] / max(ts, 1):5.1f}%")
