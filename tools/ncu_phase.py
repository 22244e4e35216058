import csv, sys, collections
# usage: ncu_phase.py file.csv kernel_substr  -> per (file,line) aggregated samples/instr, grouped by ranges given in argv
rows=csv.reader(open(sys.argv[1]))
sub=sys.argv[2]
fname=None; hdr=None; ok=False
agg=collections.defaultdict(lambda:[0.0,0.0,0.0])
for r in rows:
    if not r: continue
    if r[0]=="File Path": fname=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": ok = sub in r[1]; continue
    if r[0]=="Line No": hdr=r; iS=hdr.index("# Samples"); iI=hdr.index("Instructions Executed"); iT=hdr.index("Thread Instructions Executed"); continue
    if not ok or hdr is None or r[0]=="": continue
    try: ln=int(r[0])
    except: continue
    def f(x):
        try: return float(x or 0)
        except: return 0.0
    a=agg[(fname,ln)]; a[0]+=f(r[iS]); a[1]+=f(r[iI]); a[2]+=f(r[iT])
ts=sum(a[0] for a in agg.values()); ti=sum(a[1] for a in agg.values()); tt=sum(a[2] for a in agg.values())
print(f"total samples {ts:.0f} warp-instr {ti:.3e} thread-instr {tt:.3e} avg active {tt/ti:.1f}")
ranges=[]
for spec in sys.argv[3:]:
    name,fn,a,b=spec.split(':'); ranges.append((name,fn,int(a),int(b)))
out=collections.defaultdict(lambda:[0.0,0.0])
for (fn,ln),a in agg.items():
    nm="other:"+fn
    for name,f2,lo,hi in ranges:
        if f2 in fn and lo<=ln<=hi: nm=name; break
    out[nm][0]+=a[0]; out[nm][1]+=a[1]
for k,v in sorted(out.items(), key=lambda x:-x[1][0]): print(f"{k:40s} samples {100*v[0]/ts:5.1f}%  instr {100*v[1]/ti:5.1f}%")
