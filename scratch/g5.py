import sys, time, os; sys.path.insert(0,'.'); sys.path.insert(0,'./tests')
import numpy as np
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
import oracle_bindings as ob
from helpers import assert_bit_identical
hub=int(sys.argv[1]) if len(sys.argv)>1 else 8
for scale,K,L,it in ((11,50,100,8),(12,50,100,30),(9,300,500,6)):
    g=G.rmat(scale); col=ppr.find_partitions_csr(g)
    got=ppr.grank_csr(g,K,L,it,0.85,1e-3,colour=col,hub_threshold=hub)
    want=ob.oracle_grank(g,K,L,it,0.85,1e-3,colour=col,hub_threshold=hub)
    assert_bit_identical(got,want,f"rmat{scale}")
    for k in ("merged_entries","truncations","boundary_ties","algorithmic_bytes"):
        assert got.stats[k]==want.stats[k],(k,got.stats[k],want.stats[k])
print("parity ok hub",hub)
