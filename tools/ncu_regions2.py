"""Aggregate an ncu `--page source --print-source cuda,sass --csv` export by REGION of the kernel source.
usage: python tools/ncu_regions2.py export.csv file.cu ncells  "name:lo-hi" ...  (regions by source line; helper lines
(inline asm wrappers, intrinsics headers) are attributed to the region of the previous non-helper line in SASS order)"""
import csv, sys, collections
path, cu, ncells = sys.argv[1], sys.argv[2], float(sys.argv[3])
regions = []
for a in sys.argv[4:]:
    name, r = a.split(":"); lo, hi = r.split("-"); regions.append((name, int(lo), int(hi)))
rows = list(csv.reader(open(path)))
# locate the SASS table: header row containing "Address" and "Source"
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
# find the second table (sass) -- ncu prints the cuda view first then sass rows attached; fall back to per-line view
hdr = rows[hi_]
c_inst = [i for i, h in enumerate(hdr) if h == "Instructions Executed"][0]
c_samp = [i for i, h in enumerate(hdr) if h == "# Samples"][0]
fname = ""
agg = collections.defaultdict(lambda: [0, 0])
helper = collections.defaultdict(lambda: [0, 0])
for r in rows[hi_ + 1:]:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) < len(hdr) or r[0] in ("Line No", ""):
        continue
    try:
        ln = int(r[0]); inst = int(r[c_inst] or 0); samp = int(r[c_samp] or 0)
    except ValueError:
        continue
    if fname == cu.split("/")[-1]:
        for name, lo, hi in regions:
            if lo <= ln <= hi:
                agg[name][0] += inst; agg[name][1] += samp; break
        else:
            agg[f"line {ln}"][0] += inst; agg[f"line {ln}"][1] += samp
    else:
        helper[f"{fname}:{ln}"][0] += inst; helper[f"{fname}:{ln}"][1] += samp
tot_i = sum(v[0] for v in agg.values()) + sum(v[0] for v in helper.values())
tot_s = sum(v[1] for v in agg.values()) + sum(v[1] for v in helper.values())
print(f"total {tot_i / ncells:.0f} warp-inst/cell, {tot_s} samples")
for k, v in sorted(list(agg.items()) + list(helper.items()), key=lambda kv: -kv[1][0]):
    if v[0] / tot_i > 0.003 or v[1] / tot_s > 0.003:
        print(f"{v[0] / ncells:8.0f} inst/cell {100 * v[0] / tot_i:5.1f}%   samples {100 * v[1] / tot_s:5.1f}%   {k}")
