"""tools/time_mc.py <scale> [R] [reps] -- device time of one MCCompletePathV2 job on R-MAT <scale> (K50 L100): walks / combine"""
import sys; sys.path.insert(0, '.')
import numpy as np
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
scale = int(sys.argv[1]); R = int(sys.argv[2]) if len(sys.argv) > 2 else 1000; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = G.rmat(scale)
s = ppr.Session(g, 100, colour=np.zeros(g.n, dtype=np.uint8))
for r in range(reps):
    s.mc(50, 100, R, 0.85)
    st = s.stats(); walk = s.kernel_time(1)[1]; comb = s.kernel_time(0)[1]
    print(f"rmat{scale} R={R}: kernel_ms {st['kernel_ms']:.2f} walk_ms {walk:.2f} combine_ms {comb:.2f} hops {st['walk_steps']} -> {st['walk_steps']/walk/1e6:.2f} G hops/s (walk kernel), "
          f"{st['walk_steps']/st['kernel_ms']/1e6:.2f} G hops/s (job) requeues {st['overflow_requeues']}")
