"""Quality evaluator and graph ingest around the hot paths.

* ``ppr_exact``            batched exact PPR on the GPU (pprb200_ppr_exact) -- the reference's pprSingleSource
                           (include/internal/pprSingleSource.h:28-75) for many sources at once;
* ``kendall_correlation``  Kendall tau-b with the reference's tie accounting (include/internal/kendall.h:22-180);
* ``jaccard``              include/internal/pprInternal.h:174-186;
* ``benchmark_algorithm``  the five statistics of ppr::benchmarkAlgorithm (include/benchmarkAlgorithm.h:51-153) under
                           the same names, with the exact PPR of all sampled nodes computed in one device batch;
* ``import_graph_csv``     the demo's CSV edge-list loader (src/main.cc:78-112: repeated edges dropped, targets
                           materialised as keys) plus a binary CSR cache (``save_csr`` / ``load_csr``).

Host logic is numpy; the only device work is the power iteration, which fails loudly without a GPU.
"""
from __future__ import annotations

import ctypes as C
import math
from pathlib import Path
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .api import Baskets
from .graphs import CSRGraph

STAT_NAMES = ("jaccard average", "jaccard min", "kendall average", "kendall min", "average map size")


def ppr_exact(g: CSRGraph, sources: Sequence[int], iterations: int = 100, damping: float = 0.85,
              tolerance: float = 1e-4) -> Tuple[np.ndarray, np.ndarray, float]:
    """(scores[len(sources)][n], iterations_run[len(sources)], device ms). Parameter checks as pprSingleSource.h:37-39."""
    lib = _lib.load()
    src = np.ascontiguousarray(sources, dtype=np.int32)
    out = np.zeros((len(src), g.n), dtype=np.float64)
    its = np.zeros(max(len(src), 1), dtype=np.uint32)
    ms = C.c_double(0)
    _lib.check(lib.pprb200_ppr_exact(_lib.ptr(g.row_ptr), _lib.ptr(g.col), g.n, _lib.ptr(src), len(src), iterations, damping,
                                     tolerance, _lib.ptr(out), _lib.ptr(its), C.byref(ms)))
    return out, its[:len(src)], ms.value


def kendall_correlation(x: Sequence[float], y: Sequence[float]) -> float:
    """Kendall tau-b: (concordant - discordant) / sqrt((pairs - tiedX) * (pairs - tiedY)); when the denominator is 0 the
    reference returns 1 if as many pairs are tied in x as in y, else 0 (kendall.h:164-180). O(n^2) on purpose: n is a
    basket size, and counting pairs directly needs no argument for why a merge sort counts them."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.shape[0]
    total = n * (n - 1) // 2
    if n < 2:
        return 1.0  # total = 0: den = 0 and sameX == sameY == 0
    iu = np.triu_indices(n, k=1)
    dx = np.sign(x[iu[0]] - x[iu[1]])
    dy = np.sign(y[iu[0]] - y[iu[1]])
    same_x = int(np.count_nonzero(dx == 0))
    same_y = int(np.count_nonzero(dy == 0))
    num = int(np.sum(dx * dy))  # concordant - discordant; a pair tied in x or y contributes 0
    den = math.sqrt((total - same_x) * (total - same_y))
    if den == 0.0:
        return 1.0 if same_x == same_y else 0.0
    return num / den


def jaccard(a, b) -> float:
    a, b = set(a), set(b)
    if not a and not b:
        return 1.0
    inter = len(a & b)
    return inter / (len(a) + len(b) - inter)


def _top_ids(scores: np.ndarray, k: int) -> np.ndarray:
    """keepTop(k) of the exact vector restricted to reached nodes (score > 0), canonical ties (score desc, id asc)."""
    reached = np.flatnonzero(scores > 0)
    if reached.shape[0] <= k:
        return reached
    order = np.lexsort((reached, -scores[reached]))
    return reached[order[:k]]


def benchmark_algorithm(ppr: Baskets, g: CSRGraph, test_nodes: int, strict: bool, seed: Optional[int] = None,
                        iterations: int = 100, damping: float = 0.85, tolerance: float = 1e-4) -> Dict[str, float]:
    """ppr::benchmarkAlgorithm over dense ids. Sampling: a seeded shuffle of the candidate nodes (the reference seeds
    from random_device, benchmarkAlgorithm.h:59-60); ``strict`` skips nodes without out-edges (:72-76)."""
    if test_nodes == 0:
        raise SystemExit("testNodes must be positive")  # benchmarkAlgorithm.h:54
    nodes = np.arange(g.n, dtype=np.int64)
    if strict:
        nodes = nodes[g.out_degree() > 0]
    rng = np.random.default_rng(seed)
    rng.shuffle(nodes)
    sample = nodes[:min(len(nodes), test_nodes)]
    if len(sample) == 0:
        return {k: -1.0 for k in STAT_NAMES}
    exact, _, _ = ppr_exact(g, sample, iterations, damping, tolerance)
    jac, ken, size = [], [], []
    for i, v in enumerate(sample):
        c = int(ppr.cnt[v])
        ids = ppr.ids[v, :c]
        jac.append(jaccard(ids.tolist(), _top_ids(exact[i], c).tolist()))
        ken.append(kendall_correlation(ppr.scores[v, :c], exact[i][ids]))
        size.append(c)
    return {"jaccard average": float(np.mean(jac)), "jaccard min": float(min(min(jac), 1.0)),
            "kendall average": float(np.mean(ken)), "kendall min": float(min(min(ken), 1.0)),
            "average map size": float(np.mean(size))}


# ---- ingest --------------------------------------------------------------------------------------
def import_graph_csv(path) -> CSRGraph:
    """``node1,node2`` per line (src/main.cc:78-112). Repeated edges are dropped (:101-107), every target becomes a key
    even without out-edges (:99). Dense id = order of first appearance (target before source within a line, as
    ``graph[n2]`` runs first); ``keys[dense]`` is the integer of the file."""
    raw = np.loadtxt(path, delimiter=",", dtype=np.int64, ndmin=2)
    if raw.shape[0] == 0:
        return CSRGraph(np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), [])
    n1, n2 = raw[:, 0], raw[:, 1]
    seq = np.empty(2 * raw.shape[0], dtype=np.int64)
    seq[0::2] = n2
    seq[1::2] = n1
    keys, first = np.unique(seq, return_index=True)
    keys = keys[np.argsort(first, kind="stable")]
    dense = {int(k): i for i, k in enumerate(keys)}
    src = np.fromiter((dense[int(k)] for k in n1), dtype=np.int64, count=len(n1))
    dst = np.fromiter((dense[int(k)] for k in n2), dtype=np.int64, count=len(n2))
    n = len(keys)
    code = src * n + dst
    _, keep = np.unique(code, return_index=True)  # first occurrence of every (src, dst)
    keep.sort()
    src, dst = src[keep], dst[keep]
    order = np.argsort(src, kind="stable")  # successor lists keep file order
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=row_ptr[1:])
    return CSRGraph(row_ptr, dst[order].astype(np.int32), [int(k) for k in keys])


def save_csr(g: CSRGraph, path) -> None:
    keys = np.asarray(g.keys if g.keys is not None else [], dtype=np.int64)
    np.savez(path, row_ptr=g.row_ptr, col=g.col, keys=keys, has_keys=np.asarray(g.keys is not None))


def load_csr(path) -> CSRGraph:
    z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
    keys = [int(k) for k in z["keys"]] if bool(z["has_keys"]) else None
    return CSRGraph(np.ascontiguousarray(z["row_ptr"], dtype=np.int64), np.ascontiguousarray(z["col"], dtype=np.int32), keys)


def import_graph_cached(csv_path, cache_path=None) -> CSRGraph:
    """CSV once, binary CSR afterwards (the cache is rebuilt when the CSV is newer)."""
    csv_path = Path(csv_path)
    cache = Path(cache_path) if cache_path else csv_path.with_suffix(".csr.npz")
    if cache.exists() and cache.stat().st_mtime >= csv_path.stat().st_mtime:
        return load_csr(cache)
    g = import_graph_csv(csv_path)
    save_csr(g, cache)
    return g
