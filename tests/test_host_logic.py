"""CPU-only tests of the host half of libppr_b200.so and of the Python mirror of the reference API:
the C-ABI loads and exports every symbol of include/pprb200.h, parameter checks, findPartitions, generators."""
import ctypes as C
import io
import sys
from pathlib import Path
import contextlib

import numpy as np
import pytest

import oracle_bindings as ob
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import _lib, graphs as G
from conftest import golden_cases, load_golden, requires_ref

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.exported_symbols()
    assert {"pprb200_grank", "pprb200_mccompletepathv2", "pprb200_find_partitions", "pprb200_session_create",
            "pprb200_last_error"} <= set(names)
    for s in names:
        assert hasattr(lib, s), f"{s} declared in include/pprb200.h but not exported"
    assert b"sm_100a" in lib.pprb200_version()


def test_parameter_checks_match_reference_messages():
    """grank.h:51-55 order and text, reported through the C-ABI as PPRB200_ERR_PARAM + last_error"""
    lib = _lib.load()
    g = G.ring(4)
    cases = [((0, 3, 42, 0.5), "K must be positive"), ((2, 0, 32, 0.85), "L must be positive"),
             ((2, 1, 10, 0.5), "K must be <= L"), ((2, 2, 0, 0.5), "iterations must be positive"),
             ((2, 2, 10, 1.5), "damping must be [0,1]"), ((2, 2, 10, -1.5), "damping must be [0,1]")]
    for (K, L, it, d), msg in cases:
        rc = lib.pprb200_grank(_lib.ptr(g.row_ptr), _lib.ptr(g.col), g.n, None, K, L, it, d, 1e-4, 0, None, None, None, None)
        assert rc == -1 and lib.pprb200_last_error().decode() == msg
        rc = lib.pprb200_mccompletepathv2(_lib.ptr(g.row_ptr), _lib.ptr(g.col), g.n, K, L, it, d, 1, 3, 0, None, None, None, None)
        assert rc == -1 and lib.pprb200_last_error().decode() == msg


def test_python_mirror_exits_like_the_reference():
    """test/grankTest.cc:20-29: checks fire before the graph is touched (empty graph), message on stderr, exit 1"""
    for fn, args, msg in [(ppr.grank, ({}, 0, 3, 42, 0.5, 1e-4), "K must be positive"),
                          (ppr.grank, ({}, 2, 1, 10, 0.5, 1e-4), "K must be <= L"),
                          (ppr.grankMulti, ({}, 2, 2, 10, 0.5, 1e-4, 0), "nThreads must be positive"),
                          (ppr.mccompletepathv2, ({}, 2, 2, 0, 0.5), "iterations must be positive"),
                          (ppr.mccompletepathv2, ({}, 2, 2, 10, 1.5), "damping must be [0,1]")]:
        err = io.StringIO()
        with contextlib.redirect_stderr(err), pytest.raises(SystemExit) as e:
            fn(*args)
        assert e.value.code == 1 and err.getvalue().strip() == msg


def test_empty_graph_gives_empty_result():
    assert ppr.grank({}, 10, 30, 100, 0.85, 1e-4) == {}           # test/grankTest.cc:31-36
    assert ppr.grankMulti({}, 10, 30, 100, 0.85, 1e-4, 4) == {}
    assert ppr.mccompletepathv2({}, 10, 30, 100, 0.85) == {}


def test_malformed_graph_is_rejected():
    lib = _lib.load()
    rp = np.array([0, 1, 2], dtype=np.int64)
    col = np.array([1, 5], dtype=np.int32)  # successor 5 is not a node
    colour = np.zeros(2, dtype=np.uint8)
    assert lib.pprb200_find_partitions(_lib.ptr(rp), _lib.ptr(col), 2, _lib.ptr(colour)) == -3
    assert b"not a node" in lib.pprb200_last_error()
    with pytest.raises(KeyError):
        G.from_adjacency({1: [2]})


def test_no_cpu_fallback():
    """without a usable sm_100 device the compute entry points must fail loudly"""
    lib = _lib.load()
    if lib.pprb200_device_count() > 0:
        pytest.skip("a GPU is present")
    g = G.ring(8)
    with pytest.raises(_lib.PprB200Error) as e:
        ppr.grank_csr(g, 2, 4, 3, 0.85, 1e-3)
    assert e.value.code == -4 and "no CPU fallback" in e.value.message


@pytest.mark.parametrize("seed", range(6))
def test_find_partitions_matches_oracle(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    e = int(rng.integers(0, 3 * n))
    g = G.from_edges(n, rng.integers(0, n, e), rng.integers(0, n, e))
    assert (ppr.find_partitions_csr(g) == ob.oracle_find_partitions(g)).all()


def test_find_partitions_reference_cases():
    """test/internal/findPartitionsTest.cc:12-97 (sizes only, as the reference tests them)"""
    assert ppr.find_partitions_csr(G.from_edges(0, [], [])).size == 0
    c = ppr.find_partitions_csr(G.from_edges(100, [], []))
    assert (c == 0).all()                                                        # isolated nodes -> first
    c = ppr.find_partitions_csr(G.from_edges(100, [0] * 100, list(range(100))))  # star incl. self loop
    assert sorted([(c == 0).sum(), (c == 1).sum()]) == [1, 99]
    n = 100
    src = list(range(n)) + list(range(n, 2 * n))
    dst = list(range(n, 2 * n)) + list(range(n))
    c = ppr.find_partitions_csr(G.from_edges(2 * n, src, dst))
    assert (c == 0).sum() == n and (c == 1).sum() == n


@pytest.mark.parametrize("name", golden_cases("grank"))
def test_find_partitions_matches_reference_golden(name):
    g, z = load_golden(name)
    order = z["order"]
    assert (ppr.find_partitions_csr(g.relabel(order)) == z["colour"][order]).all()


@requires_ref
def test_find_partitions_matches_live_reference_on_rmat():
    g = G.rmat(11)
    gd, order = ob.to_reference_space(g)
    assert (ppr.find_partitions_csr(gd) == ob.ref_find_partitions(g)[order]).all()


def test_parallel_colouring_matches_the_fifo_oracle_on_large_frontiers():
    """The library colours level-synchronously on all host threads (host_graph.cc); the oracle keeps the reference's literal
    FIFO queue (pprInternal.h:54-99). Large frontiers (R-MAT hubs, BA), long chains (ring) and many components."""
    cases = [G.rmat(15), G.barabasi_albert(50000, 4), G.ring(30000)]
    rng = np.random.default_rng(11)
    n = 60000
    cases.append(G.from_edges(n, rng.integers(0, n, 45000), rng.integers(0, n, 45000)))  # thousands of small components
    m = 150000  # components that outlast the transpose-free phase's edge budget / start at a sink / follow isolated nodes
    cases.append(G.from_edges(m, np.arange(m - 1), np.arange(1, m)))          # a path: hundreds of thousands of levels
    cases.append(G.from_edges(m, np.arange(1, m), np.arange(m - 1)))          # reversed: node 0 is a sink root
    cases.append(G.from_edges(m, rng.integers(5, m, 400000), rng.integers(5, m, 400000)))  # nodes 0..4 isolated
    src = np.concatenate([[0, 1], rng.integers(2, m, 400000)])
    dst = np.concatenate([[1, 0], rng.integers(2, m, 400000)])
    cases.append(G.from_edges(m, src, dst))                                   # tiny first component, giant one later
    for g in cases:
        assert (ppr.find_partitions_csr(g) == ob.oracle_find_partitions(g)).all()


def test_host_thread_count_does_not_change_the_colouring():
    code = ("import sys; sys.path.insert(0, %r); import numpy as np, approximated_personalized_pagerank_b200 as ppr;"
            "from approximated_personalized_pagerank_b200 import graphs as G;"
            "c = ppr.find_partitions_csr(G.rmat(14)); sys.stdout.write(str(int(np.dot(c.astype(np.int64), np.arange(c.size) %% 1009))))" % str(ROOT))
    import os, subprocess
    outs = set()
    for t in ("1", "3", "16"):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PPRB200_HOST_THREADS=t))
        assert r.returncode == 0, r.stderr
        outs.add(r.stdout.strip())
    assert len(outs) == 1, outs


def test_host_pool_serialises_concurrent_callers():
    import threading
    g = G.rmat(13)
    want = ob.oracle_find_partitions(g)
    bad = []

    def work():
        for _ in range(5):
            if not (ppr.find_partitions_csr(g) == want).all():
                bad.append(1)
    ts = [threading.Thread(target=work) for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not bad


def test_malformed_successor_is_reported_with_its_edge():
    g = G.from_edges(5, [0, 1, 2], [1, 2, 3])
    col = g.col.copy()
    col[1] = 7
    with pytest.raises(_lib.PprB200Error, match="successor 7 at edge 1 is not a node"):
        ppr.find_partitions_csr(G.CSRGraph(g.row_ptr, col))


def test_rmat_generator_matches_numpy_restatement():
    a = G.rmat(9, 8, seed=5)
    b = G.rmat_numpy(9, 8, seed=5)
    assert (a.row_ptr == b.row_ptr).all() and (a.col == b.col).all()
    g = G.rmat(12)
    deg = g.out_degree()
    assert g.n == 4096 and g.n_edges == 65536 and 0.2 < (deg == 0).mean() < 0.6 and deg.max() > 500


def test_ba_generator_is_symmetric_power_law():
    g = G.barabasi_albert(2000, 4, seed=3)
    assert g.n_edges == 2 * (10 + (2000 - 5) * 4)
    src = np.repeat(np.arange(g.n), g.out_degree())
    fwd = set(zip(src.tolist(), g.col.tolist()))
    assert all((b, a) in fwd for a, b in fwd)
    assert g.out_degree().min() >= 4 and g.out_degree().max() > 40


def test_from_adjacency_keeps_order_and_multiplicity():
    g = G.from_adjacency({"b": ["a", "a", "b"], "a": []})
    assert g.keys == ["b", "a"] and g.row_ptr.tolist() == [0, 3, 3] and g.col.tolist() == [1, 1, 0]
    assert G.to_adjacency(g) == {"b": ["a", "a", "b"], "a": []}
