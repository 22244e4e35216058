"""ctypes bindings of the TEST-ONLY checkers: oracle/libppr_oracle.so (CPU restatement) and, when built,
oracle/_ref/libppr_ref.so (the unmodified reference compiled from /root/reference). Never imported by the
product package."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_PATH = ROOT / "oracle" / "libppr_oracle.so"
REF_PATH = ROOT / "oracle" / "_ref" / "libppr_ref.so"


class OracleStats(C.Structure):
    _fields_ = [("iterations_run", C.c_uint32), ("pad", C.c_uint32)] + [
        (k, C.c_uint64) for k in ("node_iterations", "nonsink_node_iterations", "edge_reads", "merged_entries", "candidates",
                                  "truncations", "boundary_ties", "algorithmic_bytes", "walk_steps", "walks")
    ] + [("max_diff", C.c_double * 2)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("pad", "max_diff")}
        d["max_diff"] = list(self.max_diff)
        return d


def P(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        if not ORACLE_PATH.exists():
            raise FileNotFoundError(f"{ORACLE_PATH} missing: run `make -C oracle` (or __graft_entry__.build())")
        _oracle = C.CDLL(str(ORACLE_PATH))
        _oracle.oracle_mc_coin_threshold.restype = C.c_uint32
        _oracle.oracle_mc_coin_threshold.argtypes = [C.c_double]
    return _oracle


def have_ref() -> bool:
    return REF_PATH.exists()


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(str(REF_PATH))
        _ref.ref_norm1.restype = C.c_double
    return _ref


class Result:
    def __init__(self, ids, scores, cnt, stats=None, seconds=None):
        self.ids, self.scores, self.cnt, self.stats, self.seconds = ids, scores, cnt, stats, seconds


def _outs(n, K):
    return (np.full((max(n, 1), K), -1, dtype=np.int32), np.zeros((max(n, 1), K), dtype=np.float64),
            np.zeros(max(n, 1), dtype=np.uint32))


def oracle_find_partitions(g):
    colour = np.zeros(max(g.n, 1), dtype=np.uint8)
    rc = oracle().oracle_find_partitions(P(g.row_ptr), P(g.col), C.c_int32(g.n), P(colour))
    assert rc == 0
    return colour[:g.n]


def oracle_grank(g, K, L, iterations, damping, tolerance, colour=None, hub_threshold=0, nthreads=0):
    if colour is None:
        colour = oracle_find_partitions(g)
    colour = np.ascontiguousarray(colour, dtype=np.uint8)
    ids, sc, cnt = _outs(g.n, K)
    st = OracleStats()
    rc = oracle().oracle_grank(P(g.row_ptr), P(g.col), C.c_int32(g.n), P(colour), C.c_uint32(K), C.c_uint32(L),
                               C.c_uint32(iterations), C.c_double(damping), C.c_double(tolerance), C.c_uint32(hub_threshold),
                               P(ids), P(sc), P(cnt), C.byref(st), C.c_int(nthreads))
    assert rc == 0, rc
    return Result(ids[:g.n], sc[:g.n], cnt[:g.n], st.as_dict())


def oracle_mc(g, K, L, R, damping, seed, rounds, hub_threshold=0, nthreads=0):
    ids, sc, cnt = _outs(g.n, K)
    st = OracleStats()
    rc = oracle().oracle_mccompletepathv2(P(g.row_ptr), P(g.col), C.c_int32(g.n), C.c_uint32(K), C.c_uint32(L), C.c_uint32(R),
                                          C.c_double(damping), C.c_uint64(seed), C.c_uint32(rounds), C.c_uint32(hub_threshold),
                                          P(ids), P(sc), P(cnt), C.byref(st), C.c_int(nthreads))
    assert rc == 0, rc
    return Result(ids[:g.n], sc[:g.n], cnt[:g.n], st.as_dict())


def oracle_ppr(g, source, iterations=100, damping=0.85, tolerance=-1.0):
    out = np.zeros(max(g.n, 1), dtype=np.float64)
    rc = oracle().oracle_ppr_single_source(P(g.row_ptr), P(g.col), C.c_int32(g.n), C.c_uint32(iterations), C.c_double(damping),
                                           C.c_double(tolerance), C.c_int32(source), P(out))
    assert rc == 0, rc
    return out[:g.n]


# ---- the unmodified reference (keys 0..n-1 inserted in ascending order) ----
def ref_iteration_order(g):
    order = np.zeros(max(g.n, 1), dtype=np.int32)
    ref().ref_iteration_order(P(g.row_ptr), P(g.col), C.c_int32(g.n), P(order))
    return order[:g.n]


def ref_find_partitions(g):
    colour = np.zeros(max(g.n, 1), dtype=np.uint8)
    ref().ref_find_partitions(P(g.row_ptr), P(g.col), C.c_int32(g.n), P(colour))
    return colour[:g.n]


def _sorted_rows(ids, sc, cnt):
    """reference baskets come back in map order: sort rows (score desc, id asc) like the canonical layout"""
    n, K = ids.shape
    for v in range(n):
        c = int(cnt[v]) if cnt[v] <= K else K
        if c > 1:
            o = np.lexsort((ids[v, :c], -sc[v, :c]))
            ids[v, :c] = ids[v, :c][o]
            sc[v, :c] = sc[v, :c][o]
    return ids, sc


def ref_grank(g, K, L, iterations, damping, tolerance, nthreads=None):
    ids, sc, cnt = _outs(g.n, K)
    sec = C.c_double(0)
    if nthreads is None:
        ref().ref_grank(P(g.row_ptr), P(g.col), C.c_int32(g.n), C.c_uint32(K), C.c_uint32(L), C.c_uint32(iterations),
                        C.c_double(damping), C.c_double(tolerance), P(ids), P(sc), P(cnt), C.byref(sec))
    else:
        ref().ref_grankMulti(P(g.row_ptr), P(g.col), C.c_int32(g.n), C.c_uint32(K), C.c_uint32(L), C.c_uint32(iterations),
                             C.c_double(damping), C.c_double(tolerance), C.c_uint32(nthreads), P(ids), P(sc), P(cnt), C.byref(sec))
    ids, sc = _sorted_rows(ids[:g.n], sc[:g.n], cnt[:g.n])
    return Result(ids, sc, cnt[:g.n], None, sec.value)


def ref_grank_althash(g, K, L, iterations, damping, tolerance):
    """the unmodified ppr::grank on the same graph under a different std::hash (the reference's self-noise probe)"""
    ids, sc, cnt = _outs(g.n, K)
    ref().ref_grank_althash(P(g.row_ptr), P(g.col), C.c_int32(g.n), C.c_uint32(K), C.c_uint32(L), C.c_uint32(iterations),
                            C.c_double(damping), C.c_double(tolerance), P(ids), P(sc), P(cnt))
    ids, sc = _sorted_rows(ids[:g.n], sc[:g.n], cnt[:g.n])
    return Result(ids, sc, cnt[:g.n])


def ref_mc(g, K, L, R, damping):
    ids, sc, cnt = _outs(g.n, K)
    sec = C.c_double(0)
    ref().ref_mccompletepathv2(P(g.row_ptr), P(g.col), C.c_int32(g.n), C.c_uint32(K), C.c_uint32(L), C.c_uint32(R),
                               C.c_double(damping), P(ids), P(sc), P(cnt), C.byref(sec))
    ids, sc = _sorted_rows(ids[:g.n], sc[:g.n], cnt[:g.n])
    return Result(ids, sc, cnt[:g.n], None, sec.value)


def ref_ppr(g, sources, iterations=100, damping=0.85, tolerance=-1.0):
    sources = np.ascontiguousarray(sources, dtype=np.int32)
    out = np.zeros((len(sources), max(g.n, 1)), dtype=np.float64)
    ref().ref_ppr_multi_source(P(g.row_ptr), P(g.col), C.c_int32(g.n), C.c_uint32(iterations), C.c_double(damping),
                               C.c_double(tolerance), P(sources), C.c_int32(len(sources)), P(out))
    return out[:, :g.n]


def to_reference_space(g):
    """The reference iterates its map in an implementation-defined order; the canonical dense id is the position
    in that order. Returns (g_dense, order) with g_dense = g relabelled so that dense id i = key order[i]."""
    order = ref_iteration_order(g)
    return g.relabel(order), order


def baskets_to_keyspace(res, order):
    """Map a dense-space result back to the reference's key space (rows indexed by key, ids = keys), rows re-sorted."""
    n, K = res.ids.shape
    ids = np.full((n, K), -1, dtype=np.int32)
    sc = np.zeros((n, K))
    cnt = np.zeros(n, dtype=np.uint32)
    for d in range(n):
        k = int(order[d])
        c = int(res.cnt[d])
        ids[k, :c] = order[res.ids[d, :c]]
        sc[k, :c] = res.scores[d, :c]
        cnt[k] = c
    ids, sc = _sorted_rows(ids, sc, cnt)
    return Result(ids, sc, cnt)


def ref_kendall(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    f = ref().ref_kendall
    f.restype = C.c_double
    return float(f(P(x), P(y), C.c_int32(len(x))))


def ref_jaccard(a, b):
    a = np.ascontiguousarray(a, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.int32)
    f = ref().ref_jaccard
    f.restype = C.c_double
    return float(f(P(a), C.c_int32(len(a)), P(b), C.c_int32(len(b))))
