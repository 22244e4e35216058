"""B200-native GRank / MCCompletePathV2 (approximate all-sources Personalized PageRank).

Drop-in for the hot paths of fruttasecca/approximated_personalized_pagerank: the C-ABI in include/pprb200.h
(libppr_b200.so, hand-written sm_100a CUDA), the reference's C++ template API in cpp/include/, and this
Python mirror of it. No CPU fallback exists: the shared library must be built and an sm_100 GPU present.
"""
from .api import (Baskets, Session, find_partitions_csr, grank, grank_csr, grankMulti, mccompletepathv2,
                  mccompletepathv2_csr, NEVER_HUB, DEFAULT_HUB_THRESHOLD)
from .graphs import CSRGraph, barabasi_albert, from_adjacency, from_edges, ring, rmat, rmat_numpy, to_adjacency

__all__ = ["Baskets", "Session", "find_partitions_csr", "grank", "grank_csr", "grankMulti", "mccompletepathv2",
           "mccompletepathv2_csr", "NEVER_HUB", "DEFAULT_HUB_THRESHOLD", "CSRGraph", "barabasi_albert", "from_adjacency", "from_edges", "ring",
           "rmat", "rmat_numpy", "to_adjacency"]
