"""Python mirror of the reference's public API, on top of the C-ABI (include/pprb200.h).

    grank(graph, K, L, iterations, damping, tolerance)                  include/grank.h:42-48
    grankMulti(graph, K, L, iterations, damping, tolerance, nThreads)   header-only/grankMulti.h:45-52
    mccompletepathv2(graph, K, L, iterations, damping)                  include/mccompletepathv2.h:182-187

``graph`` is a mapping key -> list of successor keys (every sink must be a key, README.md:68-74); the result
is ``{key: {key: score}}`` with at most K entries per node, as the reference returns it. Argument meaning,
check order, messages and the exit status on bad parameters follow the reference (it prints to stderr and
calls ``exit(EXIT_FAILURE)``; here ``SystemExit(1)`` after printing the same text). The ``*_csr`` variants
take and return flat arrays (the layout of the C-ABI) and are what the tests and bench.py use at scale.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import Dict, Hashable, Mapping, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .graphs import CSRGraph, from_adjacency

NEVER_HUB = 0xFFFFFFFF
DEFAULT_HUB_THRESHOLD = 12  # PPRB200_DEFAULT_HUB_THRESHOLD (hub_threshold = 0 selects it)
DEFAULT_MC_ROUNDS = 3
DEFAULT_MC_SEED = 0x5EED5EED5EED5EED


def _die(msg: str):
    print(msg, file=sys.stderr)
    raise SystemExit(1)


def _check(K, L, iterations, damping, nThreads=None):
    # same order and text as grank.h:51-55 / grankMulti.h:299-304 / mccompletepathv2.h:190-194
    if K == 0: _die("K must be positive")
    if L == 0: _die("L must be positive")
    if K > L: _die("K must be <= L")
    if iterations == 0: _die("iterations must be positive")
    if damping < 0 or damping > 1: _die("damping must be [0,1]")
    if nThreads is not None and nThreads == 0: _die("nThreads must be positive")


class Baskets:
    """Flat result: ids[n,K] (dense ids, -1 padded), scores[n,K], cnt[n]; rows sorted (score desc, id asc)."""

    def __init__(self, ids: np.ndarray, scores: np.ndarray, cnt: np.ndarray, stats: Optional[dict] = None):
        self.ids, self.scores, self.cnt, self.stats = ids, scores, cnt, stats

    def to_dict(self, g: CSRGraph) -> Dict[Hashable, Dict[Hashable, float]]:
        out = {}
        for v in range(g.n):
            c = int(self.cnt[v])
            out[g.key_of(v)] = {g.key_of(int(self.ids[v, i])): float(self.scores[v, i]) for i in range(c)}
        return out


def find_partitions_csr(g: CSRGraph, device: bool = False) -> np.ndarray:
    """colour[v] = 0 (partitions.first) / 1 (partitions.second); pprInternal.h:29-99. device=True: the large component is
    levelled on the GPU, as the session / one-shot entry points do for graphs of a million edges and more (same output)."""
    lib = _lib.load()
    colour = np.zeros(max(g.n, 1), dtype=np.uint8)
    fn = lib.pprb200_find_partitions_device if device else lib.pprb200_find_partitions
    _lib.check(fn(_lib.ptr(g.row_ptr), _lib.ptr(g.col), g.n, _lib.ptr(colour)))
    return colour[:g.n]


def grank_csr(g: CSRGraph, K: int, L: int, iterations: int, damping: float, tolerance: float,
              colour: Optional[np.ndarray] = None, hub_threshold: int = 0) -> Baskets:
    lib = _lib.load()
    n = g.n
    ids = np.full((n, K), -1, dtype=np.int32)
    scores = np.zeros((n, K), dtype=np.float64)
    cnt = np.zeros(max(n, 1), dtype=np.uint32)
    st = _lib.Stats()
    col_arr = None if colour is None else np.ascontiguousarray(colour, dtype=np.uint8)
    _lib.check(lib.pprb200_grank(_lib.ptr(g.row_ptr), _lib.ptr(g.col), n, _lib.ptr(col_arr), K, L, iterations,
                                 damping, tolerance, hub_threshold, _lib.ptr(ids), _lib.ptr(scores), _lib.ptr(cnt),
                                 C.byref(st)))
    return Baskets(ids, scores, cnt[:n], st.as_dict())


def mccompletepathv2_csr(g: CSRGraph, K: int, L: int, R: int, damping: float, seed: int = DEFAULT_MC_SEED,
                         rounds: int = DEFAULT_MC_ROUNDS, hub_threshold: int = 0) -> Baskets:
    lib = _lib.load()
    n = g.n
    ids = np.full((n, K), -1, dtype=np.int32)
    scores = np.zeros((n, K), dtype=np.float64)
    cnt = np.zeros(max(n, 1), dtype=np.uint32)
    st = _lib.Stats()
    _lib.check(lib.pprb200_mccompletepathv2(_lib.ptr(g.row_ptr), _lib.ptr(g.col), n, K, L, R, damping, seed, rounds,
                                            hub_threshold, _lib.ptr(ids), _lib.ptr(scores), _lib.ptr(cnt), C.byref(st)))
    return Baskets(ids, scores, cnt[:n], st.as_dict())


def grank(graph: Mapping[Hashable, Sequence[Hashable]], K: int, L: int, iterations: int, damping: float,
          tolerance: float) -> Dict[Hashable, Dict[Hashable, float]]:
    _check(K, L, iterations, damping)
    if len(graph) == 0:
        return {}
    g = from_adjacency(graph)
    return grank_csr(g, K, L, iterations, damping, tolerance).to_dict(g)


def grankMulti(graph, K: int, L: int, iterations: int, damping: float, tolerance: float, nThreads: int):
    """Same result as grank (test/grankMultiThreadTest.cc:384-576); nThreads is validated and otherwise unused:
    the data parallelism over source nodes happens on the GPU."""
    _check(K, L, iterations, damping, nThreads)
    if len(graph) == 0:
        return {}
    g = from_adjacency(graph)
    return grank_csr(g, K, L, iterations, damping, tolerance).to_dict(g)


def mccompletepathv2(graph, K: int, L: int, iterations: int, damping: float, seed: int = DEFAULT_MC_SEED,
                     rounds: int = DEFAULT_MC_ROUNDS):
    _check(K, L, iterations, damping)
    if len(graph) == 0:
        return {}
    g = from_adjacency(graph)
    return mccompletepathv2_csr(g, K, L, iterations, damping, seed, rounds).to_dict(g)


class Session:
    """Device-resident graph + baskets (pprb200_session_*): what bench.py times as `value`."""

    def __init__(self, g: CSRGraph, max_L: int, colour: Optional[np.ndarray] = None, hub_threshold: int = 0,
                 rank: int = 0, world: int = 1, stream: int = 0):
        self.lib = _lib.load()
        self.g = g
        self.handle = C.c_void_p()
        col_arr = None if colour is None else np.ascontiguousarray(colour, dtype=np.uint8)
        _lib.check(self.lib.pprb200_session_create(_lib.ptr(g.row_ptr), _lib.ptr(g.col), g.n, _lib.ptr(col_arr), max_L,
                                                   hub_threshold, rank, world, C.c_void_p(stream), C.byref(self.handle)))
        self.K = 0

    def grank(self, K, L, iterations, damping, tolerance):
        self.K = K
        _lib.check(self.lib.pprb200_session_grank(self.handle, K, L, iterations, damping, tolerance))

    def mc(self, K, L, R, damping, seed=DEFAULT_MC_SEED, rounds=DEFAULT_MC_ROUNDS):
        self.K = K
        _lib.check(self.lib.pprb200_session_mc(self.handle, K, L, R, damping, seed, rounds))

    def fetch(self, ids=None, scores=None, cnt=None) -> Baskets:
        n, K = self.g.n, self.K
        ids = np.full((n, K), -1, dtype=np.int32) if ids is None else ids
        scores = np.zeros((n, K), dtype=np.float64) if scores is None else scores
        cnt = np.zeros(max(n, 1), dtype=np.uint32) if cnt is None else cnt
        _lib.check(self.lib.pprb200_session_fetch(self.handle, _lib.ptr(ids), _lib.ptr(scores), _lib.ptr(cnt)))
        return Baskets(ids, scores, cnt[:n], None)

    def stats(self) -> dict:
        st = _lib.Stats()
        _lib.check(self.lib.pprb200_session_stats(self.handle, C.byref(st)))
        return st.as_dict()

    def kernel_time(self, which: int = 0) -> Tuple[int, float]:
        launches = C.c_uint32(0)
        ms = C.c_double(0)
        _lib.check(self.lib.pprb200_session_kernel_time(self.handle, which, C.byref(launches), C.byref(ms)))
        return launches.value, ms.value

    def launches(self) -> int:
        v = C.c_uint64(0)
        _lib.check(self.lib.pprb200_session_launches(self.handle, C.byref(v)))
        return v.value

    def close(self):
        if self.handle:
            self.lib.pprb200_session_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
