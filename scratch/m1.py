import sys, time, os; sys.path.insert(0,'.'); sys.path.insert(0,'./tests')
import numpy as np
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
scale=int(sys.argv[1]); R=int(sys.argv[2]); rounds=int(sys.argv[3]) if len(sys.argv)>3 else 3
g=G.rmat(scale)
s=ppr.Session(g,100,hub_threshold=int(os.environ.get('HUB','0')))
for rep in range(2):
    s.mc(50,100,R,0.85,rounds=rounds)
    st=s.stats(); l,ms=s.kernel_time(1); l2,ms2=s.kernel_time(0)
    print(f"rmat{scale} R={R}: kernel_ms {st['kernel_ms']:.2f} walk_ms {ms:.2f} steps {st['walk_steps']:.3e} steps/s {st['walk_steps']/ms*1e3:.3e} walk alg GB/s {st['walk_algorithmic_bytes']/ms/1e6:.1f} requeues {st['overflow_requeues']} combine_ms {ms2:.2f} ({l2} rounds) combine GB/s {st['algorithmic_bytes']/max(ms2,1e-9)/1e6:.1f}")
