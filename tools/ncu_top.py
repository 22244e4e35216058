import csv, sys, collections
# usage: ncu_top.py file.csv [kernel_index] [topN]
rows=list(csv.reader(open(sys.argv[1])))
kidx=int(sys.argv[2]) if len(sys.argv)>2 else 0
topn=int(sys.argv[3]) if len(sys.argv)>3 else 30
kern=-1; fname=None; hdr=None; out=[]; srccache={}
for r in rows:
    if not r: continue
    if r[0]=="File Path": fname=r[1].split('/')[-1]; continue
    if r[0]=="Function Name":
        k=r[1]
        if k not in srccache: srccache[k]=len(srccache)
        kern=srccache[k]; continue
    if r[0]=="Line No": hdr=r; continue
    if kern!=kidx or hdr is None: continue
    if r[0]=="" : continue   # sass rows
    try: ln=int(r[0])
    except: continue
    d=dict(zip(hdr,r))
    def num(k):
        try: return float(d.get(k,'0') or 0)
        except: return 0.0
    # hdr has duplicate 'Source' columns; r[1] is source text
    out.append((num("# Samples"), num("Instructions Executed"), fname, ln, r[1].strip()[:110], {k:num(k) for k in ("stall_barrier","stall_long_sb","stall_short_sb","stall_wait","stall_branch_resolving","stall_mio","stall_membar","stall_lg")}))
tot_s=sum(o[0] for o in out); tot_i=sum(o[1] for o in out)
print(f"kernel {kidx}: total samples {tot_s:.0f} total warp-instr {tot_i:.3e}")
for o in sorted(out,key=lambda x:-x[0])[:topn]:
    st=",".join(f"{k[6:]}={v:.0f}" for k,v in o[5].items() if v>0.08*max(o[0],1))
    print(f"{100*o[0]/tot_s:5.1f}%s {100*o[1]/tot_i:5.1f}%i {o[2]}:{o[3]:4d} {o[4]}  [{st}]")
