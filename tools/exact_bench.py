import sys, time; sys.path.insert(0,'.')
import numpy as np
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G, evaluate as EV
for scale, B, iters in ((16, 200, 100), (20, 200, 30)):
    g = G.rmat(scale)
    rng = np.random.default_rng(1)
    src = rng.choice(np.flatnonzero(g.out_degree() > 0), size=B, replace=False).astype(np.int32)
    EV.ppr_exact(g, src[:8], 2, 0.85, -1.0)
    t = time.time(); sc, its, ms = EV.ppr_exact(g, src, iters, 0.85, -1.0); wall = time.time() - t
    bytes_it = g.n_edges * (B * 8 + 4 + 8) + g.n * B * 16
    print(f"exact PPR rmat{scale} B={B} it={iters}: device {ms:.1f} ms ({ms/iters:.2f} ms/it), algorithmic {bytes_it*iters/ms/1e6:.0f} GB/s, wall {wall*1e3:.0f} ms, mass {sc.sum(1).mean():.4f}")
