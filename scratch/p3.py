import sys; sys.path.insert(0,'.')
import numpy as np
from approximated_personalized_pagerank_b200 import graphs as G, evaluate as EV
g = G.rmat(16); rng = np.random.default_rng(1)
src = rng.choice(np.flatnonzero(g.out_degree() > 0), size=200, replace=False).astype(np.int32)
print(EV.ppr_exact(g, src, 12, 0.85, -1.0)[2])
