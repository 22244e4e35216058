import sys, time, os; sys.path.insert(0,'.')
import numpy as np
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
import torch
n=int(sys.argv[1]) if len(sys.argv)>1 else 8388608
t0=time.perf_counter(); g=G.barabasi_albert(n,4); t1=time.perf_counter(); col=ppr.find_partitions_csr(g); t2=time.perf_counter()
deg=g.out_degree(); print(f"BA n={g.n} e={g.n_edges} maxdeg {deg.max()} sinks {(deg==0).sum()} gen {t1-t0:.1f}s partitions {t2-t1:.1f}s colours {np.bincount(col)}", flush=True)
s=ppr.Session(g,100,colour=col); t3=time.perf_counter()
print(f"session {t3-t2:.1f}s mem free {torch.cuda.mem_get_info()[0]/2**30:.1f} GiB", flush=True)
for rep in range(2):
    s.grank(50,100,30,0.85,-1.0)
    st=s.stats(); l,ms=s.kernel_time(0)
    print(f"BA grank: kernel_ms {st['kernel_ms']:.1f} merge_ms {ms:.1f} node_iters/s {st['node_iterations']/st['kernel_ms']*1e3:.3e} alg GB/s {st['algorithmic_bytes']/ms/1e6:.1f} requeues {st['overflow_requeues']}", flush=True)
s.mc(50,100,1000,0.85)
st=s.stats(); l,ms=s.kernel_time(1); l2,ms2=s.kernel_time(0)
print(f"BA mc: kernel_ms {st['kernel_ms']:.1f} walk_ms {ms:.1f} steps/s {st['walk_steps']/ms*1e3:.3e} combine_ms {ms2:.1f}", flush=True)
