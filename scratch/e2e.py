import sys, time, os; sys.path.insert(0,'.')
import numpy as np
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
scale=int(sys.argv[1]) if len(sys.argv)>1 else 16
g=G.rmat(scale)
for rep in range(3):
    t0=time.perf_counter()
    r=ppr.grank_csr(g,50,100,30,0.85,1e-3)
    t=time.perf_counter()-t0
    st=r.stats
    print(f"grank e2e {t*1e3:.1f} ms: prep {st['prep_ms']:.1f} h2d+alloc {st['h2d_ms']:.1f} kernel {st['kernel_ms']:.1f} d2h {st['d2h_ms']:.1f} total {st['total_ms']:.1f}")
t0=time.perf_counter(); c=ppr.find_partitions_csr(g); print("find_partitions", (time.perf_counter()-t0)*1e3, "ms")
for rep in range(2):
    t0=time.perf_counter()
    r=ppr.mccompletepathv2_csr(g,50,100,1000,0.85)
    t=time.perf_counter()-t0
    st=r.stats
    print(f"mc e2e {t*1e3:.1f} ms: prep {st['prep_ms']:.1f} h2d+alloc {st['h2d_ms']:.1f} kernel {st['kernel_ms']:.1f} d2h {st['d2h_ms']:.1f} total {st['total_ms']:.1f}")
