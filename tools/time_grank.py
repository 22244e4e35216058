"""tools/time_grank.py <scale> <iters> [reps] -- device time of one GRank job on R-MAT <scale> (K50 L100, tol -1)"""
import sys; sys.path.insert(0, '.')
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
scale = int(sys.argv[1]); iters = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
g = G.rmat(scale); col = ppr.find_partitions_csr(g)
s = ppr.Session(g, 100, colour=col)
for r in range(reps):
    s.grank(50, 100, iters, 0.85, -1.0)
    st = s.stats(); l, ms = s.kernel_time(0)
    import os
    print({k: v for k, v in os.environ.items() if k.startswith('PPRB200')}, end=' ')
    print(f"rmat{scale} it={iters}: kernel_ms {st['kernel_ms']:.2f} merge_ms {ms:.2f} frac {st['algorithmic_bytes']/ms/1e6/6543.1:.4f} requeues {st['overflow_requeues']}")
