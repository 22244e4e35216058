import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'./tests')
import numpy as np
import oracle_bindings as ob
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
def cmp(g,K,L,it,d,tol,name,hub=ppr.NEVER_HUB):
    col=ob.oracle_find_partitions(g)
    t=time.time(); o=ob.oracle_grank(g,K,L,it,d,tol,colour=col,hub_threshold=0 if hub==ppr.NEVER_HUB else hub); to=time.time()-t
    t=time.time(); r=ppr.grank_csr(g,K,L,it,d,tol,colour=col,hub_threshold=hub); tg=time.time()-t
    ids_eq=(r.ids==o.ids).all(); sc_eq=(r.scores.view(np.uint64)==o.scores.view(np.uint64)).all(); cnt_eq=(r.cnt==o.cnt).all()
    maxd=np.abs(r.scores-o.scores).max() if g.n else 0
    st=r.stats
    keys=['iterations_run','nonsink_node_iterations','edge_reads','merged_entries','candidates','truncations','boundary_ties','algorithmic_bytes']
    seq=all(st[k]==o.stats[k] for k in keys)
    print(f"{name}: ids {ids_eq} scores_bits {sc_eq} cnt {cnt_eq} max|d| {maxd:.2e} stats_eq {seq} it {st['iterations_run']} requeue {st['overflow_requeues']} kernel_ms {st['kernel_ms']:.2f} total_ms {st['total_ms']:.1f} oracle_s {to:.2f} gpu_s {tg:.2f}")
    if not seq: print('  gpu',{k:st[k] for k in keys}); print('  orc',{k:o.stats[k] for k in keys}, st['max_diff'], o.stats['max_diff'])
    return r
cmp(G.ring(100),50,100,30,0.85,1e-3,'ring config1')
cmp(G.ring(100),10,10,100,0.85,1e-4,'ring K=L=10')
cmp(G.ring(6),3,3,100,0.85,1e-4,'ring6 K=L=3')
rng=np.random.default_rng(1)
n=100; g=G.from_edges(n,rng.integers(0,n,5000),rng.integers(0,n,5000)); cmp(g,n,n,100,0.85,-1,'random100 K=L=N')
cmp(g,7,13,40,0.85,1e-5,'random100 K7 L13')
g=G.from_edges(10,[],[]); cmp(g,10,30,100,0.85,1e-4,'noedges')
g=G.rmat(10); cmp(g,2000,2000,30,0.85,-1,'rmat10 K=L=2000')
cmp(g,50,100,30,0.85,1e-3,'rmat10 K50 L100')
g=G.rmat(12); cmp(g,50,100,30,0.85,1e-3,'rmat12 K50 L100')
g=G.rmat(14); cmp(g,50,100,30,0.85,1e-3,'rmat14 K50 L100')
g=G.rmat(16); cmp(g,50,100,30,0.85,-1,'rmat16 K50 L100 tol-1')
