import sys, time, os; sys.path.insert(0,'.'); sys.path.insert(0,'./tests')
import numpy as np
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
scale=int(sys.argv[1]); hub=int(sys.argv[2]); iters=int(sys.argv[3]) if len(sys.argv)>3 else 30
g=G.rmat(scale); col=ppr.find_partitions_csr(g)
s=ppr.Session(g,100,colour=col,hub_threshold=hub)
for rep in range(2):
    s.grank(50,100,iters,0.85,-1.0)
    st=s.stats(); l,ms=s.kernel_time(0)
    print(f"rmat{scale} hub>{hub}: kernel_ms {st['kernel_ms']:.2f} merge_ms {ms:.2f} node_iters/s {st['node_iterations']/st['kernel_ms']*1e3:.3e} alg GB/s {st['algorithmic_bytes']/ms/1e6:.1f} requeues {st['overflow_requeues']}")
