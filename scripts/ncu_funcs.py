"""Warp instructions per event and stall-sample share per DEVICE FUNCTION (source lines of `ncu --page source` grouped by the
function whose definition precedes them in the .cuh file): python scripts/ncu_funcs.py <rep> <events>."""
import csv, subprocess, sys, bisect
rep=sys.argv[1]; nev=float(sys.argv[2])
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None;hdr=None;agg={}
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Line No": hdr={n:i for i,n in enumerate(r)}; ai=r.index("Address"); continue
    if hdr is None or len(r)<10: continue
    if r[ai]=="-" and r[0].isdigit():
        try: agg[(cur,int(r[0]))]=(float(r[hdr["Instructions Executed"]]),float(r[hdr["# Samples"]]))
        except ValueError: pass
# function boundaries from source
import re
bounds={}
for f in set(k[0] for k in agg):
    try: src=open('/root/repo/bourse_b200/csrc/'+f).read().splitlines()
    except Exception: continue
    b=[]
    for i,l in enumerate(src,1):
        m=re.match(r'^(?:template.*>\s*)?(?:__device__|__global__|static).*?(\w+)\(',l)
        if m and not l.startswith(' '): b.append((i,m.group(1)))
    bounds[f]=b
tot=sum(v[0] for v in agg.values()); ts=sum(v[1] for v in agg.values())
fa={}
for (f,ln),(i,s) in agg.items():
    b=bounds.get(f)
    name=f
    if b:
        k=bisect.bisect_right([x[0] for x in b],ln)-1
        name=f+':'+(b[k][1] if k>=0 else '?')
    a=fa.setdefault(name,[0,0]); a[0]+=i; a[1]+=s
for n,(i,s) in sorted(fa.items(),key=lambda x:-x[1][1]):
    print(f"{i/nev:8.2f} inst/ev {i/tot*100:5.1f}%  samples {s/ts*100:5.1f}%  {n}")
