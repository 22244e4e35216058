import csv, sys, collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
seq=[]
for r in rows[1:]:
    v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1000
    elif r[ui]=='ms': v*=1000
    seq.append((r[ki][:64],v))
idx=[i for i,(k,_) in enumerate(seq) if k.startswith('state_reset')]
run=seq[idx[0]:idx[1]] if len(idx)>1 else seq
# split by iter_end / phase_end
groups=[[]]
for k,v in run:
    groups[-1].append((k,v))
    if k.startswith('iter_end') or k.startswith('phase_end'): groups.append([])
for gi,g in enumerate(groups):
    if not g: continue
    print(f"group {gi}: total {sum(v for _,v in g):8.1f} us | " + " | ".join(f"{k.replace('void ','').replace('merge_','')[:28]} {v:.0f}" for k,v in g if v>8))
