"""Aggregate an ncu `--page source --print-source cuda,sass --csv` export by CUDA source line."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
# first occurrence of each column name (cuda-line part), sass part repeats names
def col(name, nth=0):
    idxs = [i for i, h in enumerate(hdr) if h == name]
    return idxs[nth]
c_line, c_src = 0, 1
c_samp = col("# Samples"); c_inst = col("Instructions Executed")
stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
lines = {}
cur = None
fname = ""
for r in rows[hi + 1:]:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) < len(hdr) or r[0] == "Line No": continue
    if r[c_line] != "":
        cur = (fname, int(r[c_line]))
        d = lines.setdefault(cur, {"src": r[c_src], "samp": 0, "inst": 0, "st": {}})
        try:
            d["samp"] += int(r[c_samp] or 0); d["inst"] += int(r[c_inst] or 0)
            for h, i in stall_cols.items():
                d["st"][h] = d["st"].get(h, 0) + int(r[i] or 0)
        except ValueError:
            pass
tot_s = sum(d["samp"] for d in lines.values()) or 1
tot_i = sum(d["inst"] for d in lines.values()) or 1
print(f"total samples {tot_s}  total warp-inst {tot_i}")
for ln, d in sorted(lines.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = sorted(d["st"].items(), key=lambda kv: -kv[1])[:3]
    sts = " ".join(f"{k[6:]}={v}" for k, v in st if v)
    print(f"{100*d['samp']/tot_s:5.1f}%s {100*d['inst']/tot_i:5.1f}%i {ln[0][:18]}:{ln[1]:4d}: {d['src'].strip()[:90]}   [{sts}]")
