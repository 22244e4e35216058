import sys, time, os; sys.path.insert(0,'.'); sys.path.insert(0,'./tests')
import numpy as np
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
import oracle_bindings as ob
from helpers import assert_bit_identical
# parity with the sketch path forced on (small graphs) and at scale 18
for scale,K,L,it,hub in ((11,50,100,8,8),(12,50,100,30,8),(13,20,37,10,4),(12,50,130,6,8),(10,1,1,6,2)):
    g=G.rmat(scale); col=ppr.find_partitions_csr(g)
    got=ppr.grank_csr(g,K,L,it,0.85,1e-3,colour=col,hub_threshold=hub)
    want=ob.oracle_grank(g,K,L,it,0.85,1e-3,colour=col,hub_threshold=hub)
    assert_bit_identical(got,want,f"rmat{scale}")
    for k in ("merged_entries","truncations","boundary_ties","algorithmic_bytes","iterations_run"):
        assert got.stats[k]==want.stats[k],(k,got.stats[k],want.stats[k])
print("parity ok (PPRB200_SKETCH=%s)"%os.environ.get("PPRB200_SKETCH"))
