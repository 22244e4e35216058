import sys, time, os; sys.path.insert(0,'.'); sys.path.insert(0,'./tests')
import numpy as np
import oracle_bindings as ob
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
def cmp(g,K,L,it,d,tol,name,hub):
    col=ob.oracle_find_partitions(g)
    t=time.time(); o=ob.oracle_grank(g,K,L,it,d,tol,colour=col,hub_threshold=hub); to=time.time()-t
    t=time.time(); r=ppr.grank_csr(g,K,L,it,d,tol,colour=col,hub_threshold=hub); tg=time.time()-t
    ids_eq=(r.ids==o.ids).all(); sc_eq=(r.scores.view(np.uint64)==o.scores.view(np.uint64)).all(); cnt_eq=(r.cnt==o.cnt).all()
    maxd=np.abs(r.scores-o.scores).max() if g.n else 0
    st=r.stats
    keys=['iterations_run','nonsink_node_iterations','edge_reads','merged_entries','candidates','truncations','boundary_ties','algorithmic_bytes']
    seq=all(st[k]==o.stats[k] for k in keys)
    print(f"{name} hub>{hub}: ids {ids_eq} scores_bits {sc_eq} cnt {cnt_eq} max|d| {maxd:.2e} stats_eq {seq} it {st['iterations_run']} kernel_ms {st['kernel_ms']:.2f} oracle_s {to:.2f}", flush=True)
    if not seq: print('  gpu',{k:st[k] for k in keys}); print('  orc',{k:o.stats[k] for k in keys}, st['max_diff'], o.stats['max_diff'])
    if not (ids_eq and sc_eq):
        bad=np.nonzero((r.ids!=o.ids).any(1)|(r.scores!=o.scores).any(1))[0]; print('  bad nodes',bad[:10], 'deg', np.diff(g.row_ptr)[bad[:10]])
        v=bad[0]; print(r.ids[v][:8], o.ids[v][:8]); print(r.scores[v][:8], o.scores[v][:8])
which=sys.argv[1] if len(sys.argv)>1 else 'small'
if which=='small':
    cmp(G.ring(100),50,100,30,0.85,1e-3,'ring',1)
    rng=np.random.default_rng(1)
    n=100; g=G.from_edges(n,rng.integers(0,n,5000),rng.integers(0,n,5000)); cmp(g,n,n,100,0.85,-1,'random100 K=L=N',4)
    cmp(g,7,13,40,0.85,1e-5,'random100 K7 L13',4)
    g=G.rmat(10)
    for hub in (1,8,64): cmp(g,50,100,30,0.85,1e-3,'rmat10',hub)
    g=G.rmat(12)
    for hub in (1,8,64,512): cmp(g,50,100,30,0.85,1e-3,'rmat12',hub)
    cmp(G.rmat(9),20,40,8,0.0,-1.0,'rmat9 d=0',2)
    cmp(G.rmat(9),20,40,8,1.0,-1.0,'rmat9 d=1',2)
else:
    g=G.rmat(14); cmp(g,50,100,30,0.85,1e-3,'rmat14',8)
    g=G.rmat(16)
    for hub in (8,64,1024): cmp(g,50,100,30,0.85,-1,'rmat16',hub)
