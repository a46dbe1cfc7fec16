"""Executed warp-instructions per cell by opcode (and shared-memory wavefronts) from an ncu SASS source export:
ncu -i rep --page source --csv --print-source sass > x.csv ; python tools/ncu_opcodes.py x.csv ncells"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
ncells = float(sys.argv[2])
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci, cs = hdr.index("Instructions Executed"), hdr.index("Source")
cw = hdr.index("L1 Wavefronts Shared"); cwi = hdr.index("L1 Wavefronts Shared Ideal")
op = collections.Counter(); wf = collections.Counter(); tot = 0; wtot = 0; wid = 0
for r in rows[hi + 1:]:
    try:
        n = int(r[ci])
    except (ValueError, IndexError):
        continue
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+(?:\.[0-9]+)?)", r[cs])
    if m:
        k = m.group(1)
        op[k] += n; tot += n
        try:
            wf[k] += int(r[cw]); wtot += int(r[cw]); wid += int(r[cwi])
        except ValueError:
            pass
print(f"total {tot / ncells:.0f} warp-inst/cell; smem wavefronts {wtot / ncells:.0f}/cell (ideal {wid / ncells:.0f})")
for k, v in op.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 40):
    print(f"{v / ncells:8.1f} {k}" + (f"   wavefronts {wf[k] / ncells:.0f}" if wf[k] else ""))
