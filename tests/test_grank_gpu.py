"""GPU parity tests of the GRank path (run with -m gpu on a B200). Everything goes through the C-ABI
(pprb200_grank / session API) and is checked against the oracle bit for bit, against the reference's golden
vectors (<= 1e-9, identical membership) and against the reference's own unit-test expectations."""
import numpy as np
import pytest

import oracle_bindings as ob
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
from conftest import golden_cases, load_golden
from helpers import assert_bit_identical, compare_membership, rows_as_dicts

pytestmark = pytest.mark.gpu

TOL = 1e-9  # north-star tolerance for GRank scores (fp64)


def run_pair(g, K, L, it, d, tol, hub=None):
    """GPU and oracle on the same input, same partition, same hub threshold."""
    colour = ppr.find_partitions_csr(g)
    gpu_hub = ppr.NEVER_HUB if hub is None else hub
    got = ppr.grank_csr(g, K, L, it, d, tol, colour=colour, hub_threshold=gpu_hub)
    want = ob.oracle_grank(g, K, L, it, d, tol, colour=colour, hub_threshold=0 if hub is None else hub)
    return got, want


STAT_KEYS = ["iterations_run", "node_iterations", "nonsink_node_iterations", "edge_reads", "merged_entries", "candidates",
             "truncations", "boundary_ties", "algorithmic_bytes"]


def assert_same_stats(got, want):
    for k in STAT_KEYS:
        assert got.stats[k] == want.stats[k], (k, got.stats[k], want.stats[k])


@pytest.mark.parametrize("name", golden_cases("grank"))
def test_golden_reference_vectors_default_configuration(name):
    """the configuration bench.py times (hub_threshold = 0: order-free above out-degree 12) against the unmodified reference's
    output on the tie-free fixtures -- several have out-degrees far above 12 (random100_5000, complete100, rmat10_k50_l2000):
    identical membership, |d| <= 1e-9 (the fixed-point sums differ from the fma chain by <= outdeg * 2^-62)"""
    g, z = load_golden(name)
    order = z["order"]
    got = ppr.grank_csr(g.relabel(order), int(z["K"]), int(z["L"]), int(z["iterations"]), float(z["damping"]), float(z["tolerance"]),
                        colour=None, hub_threshold=0)
    gk = ob.baskets_to_keyspace(got, order)
    assert (gk.cnt == np.minimum(z["cnt"], int(z["K"]))).all()
    mism, maxd = compare_membership(gk, ob.Result(z["ids"], z["scores"], z["cnt"]))
    assert mism == 0
    assert maxd <= TOL


@pytest.mark.parametrize("name", golden_cases("grank"))
def test_golden_reference_vectors(name):
    """GPU vs the unmodified reference's output (tie-free fixtures): identical membership, |d| <= 1e-9"""
    g, z = load_golden(name)
    order = z["order"]
    gd = g.relabel(order)
    got = ppr.grank_csr(gd, int(z["K"]), int(z["L"]), int(z["iterations"]), float(z["damping"]), float(z["tolerance"]),
                        colour=None, hub_threshold=ppr.NEVER_HUB)   # colour=None: the library's own findPartitions
    gk = ob.baskets_to_keyspace(got, order)
    assert (gk.cnt == np.minimum(z["cnt"], int(z["K"]))).all()
    mism, maxd = compare_membership(gk, ob.Result(z["ids"], z["scores"], z["cnt"]))
    assert mism == 0
    assert maxd <= TOL
    assert maxd == 0.0  # in fact bit-identical: same fma chain in the same order


@pytest.mark.parametrize("scale,K,L,it,tol", [(8, 50, 100, 30, 1e-3), (10, 50, 100, 30, 1e-3), (10, 1, 1, 10, -1.0),
                                              (10, 7, 9, 12, 1e-4), (10, 50, 50, 8, -1.0), (11, 20, 101, 10, -1.0),
                                              (10, 50, 300, 10, -1.0), (9, 512, 512, 20, -1.0), (12, 50, 100, 30, 1e-3),
                                              (13, 50, 100, 6, -1.0)])
def test_bit_identical_to_oracle_on_rmat(scale, K, L, it, tol):
    """heavy-tailed graphs with thousands of boundary ties: the canonical tie-break makes the result unique"""
    got, want = run_pair(G.rmat(scale), K, L, it, 0.85, tol)
    assert_bit_identical(got, want, f"rmat{scale} K{K} L{L}")
    assert_same_stats(got, want)


HUB_STAT_KEYS = [k for k in STAT_KEYS if k != "candidates"]  # the order-free path filters hopeless candidates before counting


@pytest.mark.parametrize("scale,hub,K,L,it,tol", [(10, 4, 50, 100, 12, -1.0), (12, 8, 50, 100, 30, 1e-3), (12, 64, 50, 100, 30, 1e-3),
                                                  (13, 1, 20, 37, 8, -1.0), (14, 8, 50, 100, 10, -1.0), (11, 8, 300, 500, 6, -1.0),
                                                  (12, 16, 1, 1, 6, -1.0), (12, 0, 50, 100, 30, 1e-3)])
def test_order_free_path_bit_identical_to_oracle(scale, hub, K, L, it, tol):
    """nodes above the hub threshold accumulate in 2^-62 fixed point: same sums whatever the warp/CTA order.
    hub = 0 is the library default (PPRB200_DEFAULT_HUB_THRESHOLD), the configuration bench.py times."""
    g = G.rmat(scale)
    colour = ppr.find_partitions_csr(g)
    got = ppr.grank_csr(g, K, L, it, 0.85, tol, colour=colour, hub_threshold=hub)
    want = ob.oracle_grank(g, K, L, it, 0.85, tol, colour=colour, hub_threshold=hub if hub else ppr.DEFAULT_HUB_THRESHOLD)
    assert_bit_identical(got, want, f"rmat{scale} hub>{hub} K{K} L{L}")
    for k in HUB_STAT_KEYS:
        assert got.stats[k] == want.stats[k], (k, got.stats[k], want.stats[k])


def test_hub_split_into_chunks_across_ctas(monkeypatch):
    """a node whose successors are spread over several CTAs (global table, last chunk selects) gives the same bits"""
    monkeypatch.setenv("PPRB200_CHUNK", "64")
    g = G.rmat(12)
    got, want = run_pair(g, 50, 100, 10, 0.85, -1.0, hub=8)
    assert_bit_identical(got, want, "chunked hubs")
    for k in HUB_STAT_KEYS:
        assert got.stats[k] == want.stats[k], (k, got.stats[k], want.stats[k])


@pytest.mark.parametrize("scale,team_deg,team_chunk,it", [(12, 129, 64, 10), (13, 200, 32, 8), (14, 129, 1000, 6)])
def test_hub_teams_give_the_same_bits_as_single_owner_ctas(monkeypatch, scale, team_deg, team_chunk, it):
    """merge_dense_kernel's hub teams (a hub's successor list cut into chunks for several CTAs, sums met in a staging area,
    the last member finishes; pass 2 served by every member) forced onto ordinary nodes: same baskets, same statistics as the
    oracle -- and hence as the single-CTA path"""
    monkeypatch.setenv("PPRB200_TEAM_DEG", str(team_deg))
    monkeypatch.setenv("PPRB200_TEAM_CHUNK", str(team_chunk))
    g = G.rmat(scale)
    got, want = run_pair(g, 50, 100, it, 0.85, -1.0, hub=8)
    assert_bit_identical(got, want, f"hub teams rmat{scale} deg>{team_deg} chunk {team_chunk}")
    for k in HUB_STAT_KEYS:
        assert got.stats[k] == want.stats[k], (k, got.stats[k], want.stats[k])
    monkeypatch.setenv("PPRB200_NO_TEAMS", "1")
    solo = ppr.grank_csr(g, 50, 100, it, 0.85, -1.0, colour=ppr.find_partitions_csr(g), hub_threshold=8)
    assert_bit_identical(got, solo, "teams vs single-owner CTAs")


@pytest.mark.parametrize("limit", [8, 40])
def test_tail_table_overflow_is_finished_in_rounds_not_handed_over(monkeypatch, limit):
    """More surviving tail labels than the tail table admits (here: the table artificially closed after a few labels; on
    R-MAT-22 some 600 hubs per iteration do it by themselves): merge_dense_kernel repeats pass 2 slice by slice over the hash
    range, its labels moving to the compact candidate arrays round by round, instead of sending the node to
    merge_par_kernel. Same bits as the oracle, and the bookkeeping counters show that the path was taken."""
    import ctypes as C
    from approximated_personalized_pagerank_b200 import _lib
    monkeypatch.setenv("PPRB200_TAIL_LIMIT", str(limit))
    g = G.rmat(14)
    colour = ppr.find_partitions_csr(g)
    s = ppr.Session(g, 100, colour=colour, hub_threshold=0)
    try:
        s.grank(50, 100, 8, 0.85, -1.0)
        got = s.fetch()
        got.stats = s.stats()
        d = (C.c_ulonglong * 8)()
        _lib.load().pprb200_debug_counters(s.handle, d)
    finally:
        s.close()
    want = ob.oracle_grank(g, 50, 100, 8, 0.85, -1.0, colour=colour, hub_threshold=ppr.DEFAULT_HUB_THRESHOLD)
    assert_bit_identical(got, want, f"pass-2 rounds, tail limit {limit}")
    for k in HUB_STAT_KEYS:
        assert got.stats[k] == want.stats[k], (k, got.stats[k], want.stats[k])
    assert d[5] > 0 and d[6] > 0, list(d)   # tail table full, and finished in rounds


def test_not_full_baskets_stay_on_the_dense_kernel_or_are_handed_over_same_bits(monkeypatch):
    """first sweeps: the old basket is not full and bounds nothing; the L-th largest exact candidate does (default), or the
    node goes to merge_par_kernel unread (PPRB200_MIN_OLD=0, round 2's first scheme)"""
    g = G.rmat(13)
    res = []
    for mo in ("1", "0", "30"):
        monkeypatch.setenv("PPRB200_MIN_OLD", mo)
        got, want = run_pair(g, 50, 100, 5, 0.85, -1.0, hub=12)
        assert_bit_identical(got, want, f"min_old {mo}")
        res.append(got.stats["overflow_requeues"])
    assert res[0] <= res[2] <= res[1] and res[0] < res[1], res


def test_two_pass_sketch_path_on_a_large_graph():
    """graphs with more than 16 x 8192 nodes take the sketch-filtered two-pass merge for single-item hubs"""
    g = G.rmat(18)
    got, want = run_pair(g, 50, 100, 6, 0.85, -1.0, hub=8)
    assert_bit_identical(got, want, "rmat18 two-pass")
    for k in HUB_STAT_KEYS:
        assert got.stats[k] == want.stats[k], (k, got.stats[k], want.stats[k])


def test_order_free_path_vs_exact_order_stays_inside_the_tie_noise_band():
    """Per contribution the fixed-point path differs from the reference's fma chain by <= 2^-63. What that changes is
    WHICH of several mathematically tied candidates survives a cut: sums that are equal as real numbers are exactly
    equal in fixed point (-> canonical id order) but differ by an ulp in the fma chain, order-dependently. The reference
    itself moves by up to 1.5e-3 under a different hash order on these graphs; the two paths must
    stay inside that band, with the same stopping iteration and near-identical membership."""
    g = G.rmat(13)
    a, _ = run_pair(g, 50, 100, 30, 0.85, 1e-3, hub=None)
    b = ppr.grank_csr(g, 50, 100, 30, 0.85, 1e-3, hub_threshold=8)
    assert a.stats["iterations_run"] == b.stats["iterations_run"]
    da, db = rows_as_dicts(a.ids, a.scores, a.cnt), rows_as_dicts(b.ids, b.scores, b.cnt)
    overlap = np.mean([len(set(x) & set(y)) / max(1, len(set(x) | set(y))) for x, y in zip(da, db)])
    _, maxd = compare_membership(a, b)
    print(f"order-free vs exact-order: mean Jaccard {overlap:.4f}, max |d| on common keys {maxd:.2e}")
    assert overlap >= 0.97 and maxd <= 2e-3


@pytest.mark.parametrize("damping", [0.0, 0.5, 1.0])
def test_damping_edge_values(damping):
    got, want = run_pair(G.rmat(9), 20, 40, 8, damping, -1.0)
    assert_bit_identical(got, want, f"damping {damping}")
    got, want = run_pair(G.rmat(9), 20, 40, 8, damping, -1.0, hub=4)
    assert_bit_identical(got, want, f"damping {damping} order-free")


@pytest.mark.parametrize("tol", [0.01, 0.0005, 1e-5, 0.001, 0.0, -1.0])
def test_convergence_stops_at_the_same_iteration(tol):
    """grankMultiThreadTest.cc:384-479 uses these tolerances; the device-side flag must stop where the oracle does"""
    rng = np.random.default_rng(5)
    g = G.from_edges(300, rng.integers(0, 300, 2000), rng.integers(0, 300, 2000))
    got, want = run_pair(g, 300, 300, 60, 0.85, tol)
    assert got.stats["iterations_run"] == want.stats["iterations_run"]
    assert_bit_identical(got, want, f"tol {tol}")


def test_huge_iteration_bound_with_a_tolerance_costs_only_what_ran():
    """A caller's safety bound of a million iterations must not enqueue a million no-op sweeps: the loop is enqueued in
    windows and stops once the device-side convergence flag is clear (the reference stops at convergence, grank.h:92)"""
    import time
    rng = np.random.default_rng(5)
    g = G.from_edges(300, rng.integers(0, 300, 2000), rng.integers(0, 300, 2000))
    t0 = time.perf_counter()
    got, want = run_pair(g, 300, 300, 1_000_000, 0.85, 1e-5)
    sec = time.perf_counter() - t0
    assert 64 < got.stats["iterations_run"] == want.stats["iterations_run"] < 1000   # (converges in the second window)
    assert_bit_identical(got, want, "1e6 iterations bound")
    assert sec < 30, sec     # (a million enqueued sweeps of ~20 launches each take minutes)
    # a bound that is actually needed runs through several windows
    got, want = run_pair(g, 300, 300, 200, 0.85, 0.0)
    assert got.stats["iterations_run"] == want.stats["iterations_run"]
    assert_bit_identical(got, want, "200 iterations, tolerance 0")


def test_single_iteration_and_two_iterations():
    g = G.rmat(9)
    for it in (1, 2, 3):
        got, want = run_pair(g, 30, 60, it, 0.85, 10.0)  # huge tolerance: still at least 2 iterations when allowed
        assert got.stats["iterations_run"] == want.stats["iterations_run"] == min(it, 2)
        assert_bit_identical(got, want, f"iterations {it}")


def test_multi_edges_and_self_loops_count():
    """test/grankTest.cc:60,79: duplicates and self loops are weighted by multiplicity"""
    g = G.from_edges(5, [0, 0, 0, 1, 1, 2, 3, 3, 3, 3], [1, 1, 0, 2, 2, 2, 4, 4, 4, 0])
    got, want = run_pair(g, 5, 5, 50, 0.85, -1.0)
    assert_bit_identical(got, want, "multigraph")


def test_ragged_and_degenerate_graphs():
    for g in (G.from_edges(1, [], []), G.from_edges(1, [0], [0]), G.from_edges(2, [0, 1], [1, 0]), G.from_edges(10, [], []),
              G.from_edges(40, [0] * 39, list(range(1, 40))), G.from_edges(40, list(range(1, 40)), [0] * 39)):
        got, want = run_pair(g, 10, 30, 20, 0.85, 1e-4)
        assert_bit_identical(got, want, f"n={g.n} e={g.n_edges}")


def test_big_L_uses_the_global_table_stage():
    """L > N and L large enough that no shared-memory table class is usable"""
    got, want = run_pair(G.rmat(8), 300, 5000, 6, 0.85, -1.0)
    assert_bit_identical(got, want, "L=5000")


def test_deterministic_run_to_run():
    g = G.rmat(11)
    a = ppr.grank_csr(g, 50, 100, 10, 0.85, -1.0)
    b = ppr.grank_csr(g, 50, 100, 10, 0.85, -1.0)
    assert_bit_identical(a, b, "repeat")


# ---- the reference's own unit tests, re-expressed through the dict API (test/grankTest.cc) ----
def test_ref_no_edges_single_node_two_nodes():
    res = ppr.grank({i: [] for i in range(10)}, 10, 30, 100, 0.85, 1e-4)            # :38-50
    assert len(res) == 10 and all(len(res[i]) == 1 and abs(res[i][i] - 0.15) < 10e-5 for i in range(10))
    res = ppr.grank({0: [0]}, 10, 30, 100, 0.85, 1e-4)                               # :70-84
    assert abs(res[0][0] - 1.0) < 10e-5
    res = ppr.grank({0: [1], 1: [0]}, 10, 30, 100, 0.85, 1e-4)                       # :86-105
    assert len(res[0]) == 2 and res[0][0] > res[0][1] and res[1][1] > res[1][0]


def test_ref_top_l_sizes():
    rng = np.random.default_rng(3)                                                   # :52-68
    graph = {i: [] for i in range(30)}
    for _ in range(30):
        graph[int(rng.integers(30))].append(int(rng.integers(30)))
    for i in range(1, 30, 4):
        res = ppr.grank(graph, i, i, 100, 0.85, 1e-4)
        assert len(res) == 30 and all(len(b) <= i for b in res.values())


def test_ref_line_and_ring_monotone():
    graph = {i: [(i + 1) % 6] for i in range(6)}                                     # :107-152
    for K, L in ((10, 30), (3, 4), (3, 3)):
        res = ppr.grankMulti(graph, K, L, 100, 0.85, 1e-4, 4)
        for i in range(6):
            size = min(K, 6)
            assert len(res[i]) == size
            for u in range(min(size, 3) - 1 if K == 3 else size - 1):
                # the reference reads with operator[]: a missing key reads as 0 (L=4 wraps the truncated tail around)
                assert res[i].get((i + u) % 6, 0.0) > res[i].get((i + u + 1) % 6, 0.0)
    graph = {i: [(i + 1) % 100] for i in range(100)}                                 # :184-283
    for K, L, tol in ((10, 10, 1e-4), (10, 20, 1e-4), (10, 100, 1e-4), (100, 100, -1), (200, 200, -1)):
        res = ppr.grank(graph, K, L, 100, 0.85, tol)
        for i in range(100):
            size = min(K, 100)
            assert len(res[i]) == size
            for u in range(size - 1):
                assert res[i].get((i + u + 1) % 100, 0.0) > 0 and res[i].get((i + u) % 100, 0.0) > res[i].get((i + u + 1) % 100, 0.0)


def test_ref_star():
    graph = {i: ([0] if i else []) for i in range(6)}                                # :154-182
    res = ppr.grank(graph, 10, 30, 100, 0.85, 1e-4)
    assert len(res[0]) == 1 and abs(res[0][0] - 0.15) < 10e-5
    assert all(len(res[i]) == 2 and abs(res[i][0] - 0.15 * 0.85) < 10e-5 for i in range(1, 6))
    graph[0].append(0)
    res = ppr.grank(graph, 10, 30, 100, 0.85, 1e-4)
    assert all(len(res[i]) == 2 and abs(res[i][0] - 0.85) < 10e-5 for i in range(1, 6))


@pytest.mark.parametrize("name", ["ppr_ring100", "ppr_rmat10"])
def test_ref_same_as_pagerank(name):
    """:285-379 -- K=L=N, tol -1, 100 iterations equals pprSingleSource (golden, from the reference) within 10e-5"""
    g, z = load_golden(name)
    n = g.n
    got = ppr.grank_csr(g, n, n, 100, 0.85, -1.0, hub_threshold=ppr.NEVER_HUB)
    for i, s in enumerate(z["sources"]):
        dense = np.zeros(n)
        dense[got.ids[s, :got.cnt[s]]] = got.scores[s, :got.cnt[s]]
        assert np.abs(dense - z["ppr"][i]).max() < 10e-5


# ---- BASELINE full size: bit parity (the oracle finishes R-MAT-16 in seconds) + size-independent properties ----
@pytest.mark.parametrize("hub", [None, 8])
def test_rmat16_full_config_bit_identical_and_properties(hub):
    g = G.rmat(16)
    got, want = run_pair(g, 50, 100, 30, 0.85, 1e-3, hub=hub)
    assert_bit_identical(got, want, "rmat16")
    for k in (STAT_KEYS if hub is None else HUB_STAT_KEYS):
        assert got.stats[k] == want.stats[k], (k, got.stats[k], want.stats[k])
    sc, ids, cnt = got.scores, got.ids, got.cnt
    valid = np.arange(50)[None, :] < cnt[:, None]
    assert (np.diff(sc, axis=1)[valid[:, 1:]] <= 0).all()                       # sorted descending
    assert ((sc >= 0) & (sc <= 1)).all() and (sc.sum(1) <= 1 + 1e-12).all()     # PPR mass
    assert (ids[~valid] == -1).all() and (ids[valid] >= 0).all() and (ids[valid] < g.n).all()
    sinks = g.out_degree() == 0
    assert (cnt[sinks] == 1).all() and (ids[sinks, 0] == np.nonzero(sinks)[0]).all() and np.allclose(sc[sinks, 0], 0.15, atol=1e-15)
    srt = np.sort(np.where(valid, ids, np.arange(-50, 0)[None, :]), axis=1)
    assert (np.diff(srt, axis=1) != 0).all()                                    # keys unique inside a basket


def test_device_colouring_equals_the_host_colouring():
    """The session / one-shot entry points level the first non-trivial component with a BFS on the device (plan_device.cuh);
    the colouring must be the host's (= the reference's FIFO BFS, tests/test_host_logic.py) on every kind of graph."""
    rng = np.random.default_rng(11)
    graphs = [G.rmat(12), G.rmat(16), G.rmat(18), G.barabasi_albert(100_000, 4), G.ring(100),
              G.ring(1000),                                                     # deeper than the byte-sized level counter: host
              G.from_edges(1, [], []), G.from_edges(5, [], []),                 # isolated nodes only
              G.from_edges(6, [3, 4], [4, 5]),                                  # nodes 0-2 isolated, the root is node 3
              G.from_edges(6, [3, 1], [0, 0]),                                  # node 0 is a sink root (in-edges only)
              G.from_edges(400, rng.integers(0, 200, 900), rng.integers(0, 200, 900)),  # a component + isolated nodes
              G.from_edges(400, np.r_[rng.integers(0, 200, 900), rng.integers(200, 400, 900)],
                           np.r_[rng.integers(0, 200, 900), rng.integers(200, 400, 900)])]   # two large components
    for g in graphs:
        want = ppr.find_partitions_csr(g)
        got = ppr.find_partitions_csr(g, device=True)
        assert (got == want).all(), (g.n, g.n_edges, int((got != want).sum()))


def test_one_shot_call_with_the_plan_made_on_the_device_matches_the_oracle():
    """R-MAT-16 has 2^20 edges: pprb200_grank (colour = NULL) uploads the CSR, colours and encodes on the device"""
    g = G.rmat(16)
    got = ppr.grank_csr(g, 50, 100, 6, 0.85, -1.0, hub_threshold=0)
    want = ob.oracle_grank(g, 50, 100, 6, 0.85, -1.0, hub_threshold=ppr.DEFAULT_HUB_THRESHOLD)
    assert_bit_identical(got, want, "rmat16 one-shot, device plan")
