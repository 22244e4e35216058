"""tools/time_grank.py <scale | ba:<nodes>> <iters> [reps] [hub] -- device time of one GRank job on R-MAT <scale> or a
Barabasi-Albert graph (K50 L100, tol -1); PPRB200_COUNTERS=1 also prints the merge_dense bookkeeping counters"""
import sys, os, ctypes as C; sys.path.insert(0, '.')
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G, _lib
iters = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2; hub = int(sys.argv[4]) if len(sys.argv) > 4 else 0
name = sys.argv[1]
g = G.barabasi_albert(int(name[3:]), 4) if name.startswith('ba:') else G.rmat(int(name))
col = ppr.find_partitions_csr(g)
s = ppr.Session(g, 100, colour=col, hub_threshold=hub)
for r in range(reps):
    s.grank(50, 100, iters, 0.85, -1.0)
    st = s.stats(); l, ms = s.kernel_time(0)
    print({k: v for k, v in os.environ.items() if k.startswith('PPRB200')}, end=' ')
    print(f"{name} it={iters}: kernel_ms {st['kernel_ms']:.2f} merge_ms {ms:.2f} frac {st['algorithmic_bytes']/ms/1e6/6543.1:.4f} requeues {st['overflow_requeues']}")
    if os.environ.get('PPRB200_COUNTERS'):
        d = (C.c_ulonglong * 8)()
        _lib.load().pprb200_debug_counters(s.handle, d)
        print("  dense: nodes %d pass2 %d tau0 %d successors-handed-over-after-reading %d wide %d tailfull %d finished-in-rounds %d oldnotfull %d" % tuple(d[i] for i in range(8)))
