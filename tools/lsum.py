import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit'); mi=hdr.index('Metric Name')
d=collections.defaultdict(list)
for r in rows[1:]:
    if r[mi]!='gpu__time_duration.sum': continue
    v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1000
    elif r[ui]=='ms': v*=1000
    d[r[ki].replace('pprb200::','').replace('void ','')[:44]].append(v)
tot=sum(sum(v) for v in d.values())
for k,v in d.items(): print(f"  {k:46s} n={len(v):4d} total {sum(v)/1000:9.2f} ms {100*sum(v)/tot:5.1f}%  last6 {[round(x) for x in v[-6:]]}")
