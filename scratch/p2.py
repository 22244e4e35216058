import sys, os; sys.path.insert(0,'.')
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
scale=int(sys.argv[1]); R=int(sys.argv[2]); rounds=int(sys.argv[3])
g=G.rmat(scale)
s=ppr.Session(g,100)
s.mc(50,100,R,0.85,rounds=rounds)
st=s.stats(); l,ms=s.kernel_time(1)
print(st['kernel_ms'], 'walk_ms', ms, 'steps/s', st['walk_steps']/ms*1e3, 'requeues', st['overflow_requeues'])
