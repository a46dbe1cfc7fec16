"""Per-region instruction / shared-memory-wavefront / stall-sample budget of condense_dmma_kernel from an ncu
source-page export:  ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv ;
python tools/ncu_regions.py src.csv <cells per launch>"""
import csv
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
ncell = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ci = hdr.index("Instructions Executed"); cw = hdr.index("L1 Wavefronts Shared"); cs = hdr.index("# Samples")
lines = {}
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) < len(hdr) or r[0] in ("Line No", ""):
        continue
    try:
        key = (fname, int(r[0]))
    except ValueError:
        continue
    d = lines.setdefault(key, dict(src=r[1], inst=0, wf=0, samp=0))

    def iv(x):
        try:
            return int(x or 0)
        except ValueError:
            return 0
    d["inst"] += iv(r[ci]); d["wf"] += iv(r[cw]); d["samp"] += iv(r[cs])
ti = sum(d["inst"] for d in lines.values()); tw = sum(d["wf"] for d in lines.values()); ts = sum(d["samp"] for d in lines.values())
print(f"per cell: {ti / ncell:.0f} warp-instructions, {tw / ncell:.0f} shared-memory wavefronts; {ts} stall samples")
src = open(os.path.join(os.path.dirname(__file__), "..", "gridaphybrid.jl_b200", "csrc", "condense_dmma.cu")).read().split("\n")


def find(sub, start=0):
    for i in range(start, len(src)):
        if sub in src[i]:
            return i + 1
    raise KeyError(sub)


k0 = find("condense_dmma_kernel(DmmaTables tb")
uw0 = find("================ update warps", k0)
marks = [("helpers (dmma, cp.async, barriers)", 1), ("panel_factor (W0)", find("void panel_factor(")),
         ("invert_unit_lower (owner)", find("void invert_unit_lower(")), ("invert_upper (W0)", find("void invert_upper(")),
         ("kernel prologue", k0), ("loader", find("load + re-layout", k0)),
         ("direct S loads / cp.async wait", find("S accumulators of the owned bottom row tiles (column tiles SJ0..CT-1) as C", k0)),
         ("W0 panel loop (barriers)", find("================ panel warp", k0)), ("UW: S accumulators from Bt", uw0),
         ("UW: stage head, fragments of the panel", find("for (int p = 0; p < NP; ++p) {", uw0)),
         ("UW: column tiles (gather, U12, trailing)", find("for (int J = Jfirst; J < CT; J += 3)", k0)),
         ("UW: bottom L21 = X inv(U)", find("---- bottom block, owned row tiles", k0)),
         ("UW: bottom, A21 tiles in smem", find("A21 part still in shared memory", k0)),
         ("UW: bottom, S tiles in registers", find("// S part in registers", k0)),
         ("UW: partial last panel", find("partial last panel (npiv < 8)", k0)),
         ("store S, g", find("---- store S, g", k0)), ("(backward kernel)", find("// ---- backward static condensation", k0))]
agg = {m[0]: dict(inst=0, wf=0, samp=0) for m in marks}
for (f, ln), d in lines.items():
    if not f.startswith("condense_dmma"):
        a = agg.setdefault("intrinsics: " + f, dict(inst=0, wf=0, samp=0))
    else:
        name = marks[0][0]
        for nm, st in marks:
            if ln >= st:
                name = nm
        a = agg[name]
    for k in ("inst", "wf", "samp"):
        a[k] += d[k]
print(f"| {'region':42s} | inst/cell | % | smem wavefronts/cell | % | stall samples % |")
print("|---|---|---|---|---|---|")
for nm, a in agg.items():
    if a["inst"] or a["samp"]:
        print(f"| {nm:42s} | {a['inst'] / ncell:.0f} | {100 * a['inst'] / ti:.1f} | {a['wf'] / ncell:.0f} | {100 * a['wf'] / max(tw, 1):.1f} | {100 * a['samp'] / ts:.1f} |")
