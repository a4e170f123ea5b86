"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line."""
import csv, collections, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rows=list(csv.reader(open(path)))
cur=None; data=collections.defaultdict(list)
hdr=None
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1]; continue
    if len(r)>5 and r[0]=="Line No": hdr=r; continue
    if hdr and len(r)==len(hdr) and cur: data[cur].append(r)
iL=hdr.index("Line No"); iS=1; iI=hdr.index("Instructions Executed"); iSm=hdr.index("# Samples")
agg=[]
for f,v in data.items():
    for r in v:
        if not r[iL].strip(): continue
        try: n=int(r[iI]); s=int(r[iSm])
        except: continue
        if n or s: agg.append((f.split('/')[-1],int(r[iL]),n,s,r[iS].strip()[:110]))
tot_i=sum(a[2] for a in agg); tot_s=sum(a[3] for a in agg)
print('total instr',tot_i,'samples',tot_s)
pf=collections.Counter(); ps=collections.Counter()
for a in agg: pf[a[0]]+=a[2]; ps[a[0]]+=a[3]
for k in pf: print(k, f"{100*pf[k]/tot_i:.1f}% instr  {100*ps[k]/tot_s:.1f}% samples")
print("--- lines by source order (>=0.3% of either)")
for a in sorted(agg,key=lambda a:(a[0],a[1])):
    if a[2]/tot_i>=0.003 or a[3]/tot_s>=0.003:
        print(f"{a[0]}:{a[1]:4d} {100*a[2]/tot_i:5.2f}%i {100*a[3]/tot_s:5.2f}%s  {a[4]}")
