import sys, os; sys.path.insert(0,'.')
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
scale=int(sys.argv[1]); iters=int(sys.argv[2])
g=G.rmat(scale); col=ppr.find_partitions_csr(g)
s=ppr.Session(g,100,colour=col)
s.grank(50,100,iters,0.85,-1.0)
print(s.stats()['kernel_ms'])
