import sys, time, os; sys.path.insert(0,'.')
import numpy as np
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
import torch
scale=int(sys.argv[1]); iters=int(sys.argv[2]) if len(sys.argv)>2 else 30; hub=int(sys.argv[3]) if len(sys.argv)>3 else 0
t0=time.perf_counter(); g=G.rmat(scale); t1=time.perf_counter(); col=ppr.find_partitions_csr(g); t2=time.perf_counter()
s=ppr.Session(g,100,colour=col,hub_threshold=hub); t3=time.perf_counter()
print(f"rmat{scale}: gen {t1-t0:.1f}s partitions {t2-t1:.1f}s session {t3-t2:.1f}s  mem {torch.cuda.mem_get_info()[0]/2**30:.1f} GiB free", flush=True)
for rep in range(2):
    s.grank(50,100,iters,0.85,-1.0)
    st=s.stats(); l,ms=s.kernel_time(0)
    print(f"rmat{scale} hub>{hub}: kernel_ms {st['kernel_ms']:.2f} merge_ms {ms:.2f} node_iters/s {st['node_iterations']/st['kernel_ms']*1e3:.3e} alg GB/s {st['algorithmic_bytes']/ms/1e6:.1f} requeues {st['overflow_requeues']} ties {st['boundary_ties']}/{st['truncations']}", flush=True)
