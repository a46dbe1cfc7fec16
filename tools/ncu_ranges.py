"""Per-range (line span) sample / instruction summary of an ncu cuda,sass source export."""
import csv, sys
path = sys.argv[1]; ncells = int(sys.argv[2])
ranges = [tuple(x.split(":")) for x in sys.argv[3:]]   # name:lo:hi
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
c_samp = [i for i,h in enumerate(hdr) if h=="# Samples"][0]; c_inst=[i for i,h in enumerate(hdr) if h=="Instructions Executed"][0]
stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
fname=""; agg={}
for r in rows[hi+1:]:
    if r and r[0]=="File Path": fname=r[1].split("/")[-1]; continue
    if len(r)<len(hdr) or r[0]=="Line No": continue
    if r[0]!="":
        key=(fname,int(r[0])); d=agg.setdefault(key,{"src":r[1],"s":0,"i":0,"st":{}})
        try:
            d["s"]+=int(r[c_samp] or 0); d["i"]+=int(r[c_inst] or 0)
            for h,i in stall_cols.items(): d["st"][h]=d["st"].get(h,0)+int(r[i] or 0)
        except ValueError: pass
tot=sum(d["s"] for d in agg.values()); toti=sum(d["i"] for d in agg.values())
print(f"total samples {tot}, warp-inst/cell {toti/ncells:.0f}")
for name,lo,hi_ in ranges:
    lo=int(lo); hi_=int(hi_)
    sel=[d for k,d in agg.items() if k[0]=="" and lo<=k[1]<=hi_]
    s=sum(d["s"] for d in sel); i=sum(d["i"] for d in sel)
    st={}
    for d in sel:
        for a,b in d["st"].items(): st[a]=st.get(a,0)+b
    top=sorted(st.items(), key=lambda kv:-kv[1])[:4]
    print(f"{name:14s} {lo:4d}-{hi_:4d}: samples {100*s/tot:5.1f}%  inst/cell {i/ncells:7.0f}  " + " ".join(f"{a[6:]}={100*b/max(s,1):.0f}%" for a,b in top))
oth=[d for k,d in agg.items() if k[0]!=""]
print("intrinsics hdrs: samples %.1f%% inst/cell %.0f" % (100*sum(d['s'] for d in oth)/tot, sum(d['i'] for d in oth)/ncells))
