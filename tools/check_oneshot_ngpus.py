"""tools/check_oneshot_ngpus.py <scale> <iters> <n_gpus> -- the one-shot call on 1 GPU and on n_gpus GPUs (one process) must return
the same bits, GRank and MC (R-MAT <scale>; above 2^20 edges the plan -- colouring, column words, need masks -- is made on the device)"""
import os, sys; sys.path.insert(0, '.')
import numpy as np
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
scale, iters, ng = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = G.rmat(scale)
res = {}
for n in (1, ng):
    os.environ['PPR_NUM_GPUS'] = str(n)
    a = ppr.grank_csr(g, 50, 100, iters, 0.85, -1.0)
    b = ppr.mccompletepathv2_csr(g, 50, 100, 200, 0.85)
    assert a.stats['n_gpus'] == n and b.stats['n_gpus'] == n, (a.stats['n_gpus'], n)
    res[n] = (a, b)
for i, name in enumerate(('grank', 'mc')):
    x, y = res[1][i], res[ng][i]
    same = (x.ids == y.ids).all() and (x.scores.view(np.uint64) == y.scores.view(np.uint64)).all() and (x.cnt == y.cnt).all()
    print(f"rmat{scale} {name}: 1 GPU vs {ng} GPUs bit-identical: {bool(same)}; kernel_ms {x.stats['kernel_ms']:.1f} vs {y.stats['kernel_ms']:.1f}")
    assert same
