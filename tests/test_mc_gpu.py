"""GPU parity tests of the MCCompletePathV2 path (run with -m gpu on a B200), all through the C-ABI.

Bit parity is against oracle/ppr_oracle.c:oracle_mccompletepathv2 (same Philox streams, integer visit counts,
canonical ties). Against the reference itself only what the north-star asks can hold (the reference's MC output is
seeded from random_device and order-dependent): its unit-test expectations, and a mean L1 error vs
exact power-iteration PPR that is no worse than the reference's at the same walk budget."""
import numpy as np
import pytest

import oracle_bindings as ob
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
from conftest import requires_ref
from helpers import assert_bit_identical

pytestmark = pytest.mark.gpu

SEED = ppr.api.DEFAULT_MC_SEED


def run_pair(g, K, L, R, d, rounds, hub=None, seed=SEED):
    got = ppr.mccompletepathv2_csr(g, K, L, R, d, seed=seed, rounds=rounds, hub_threshold=ppr.NEVER_HUB if hub is None else hub)
    want = ob.oracle_mc(g, K, L, R, d, seed, rounds, hub_threshold=0 if hub is None else hub)
    return got, want


@pytest.mark.parametrize("scale,K,L,R,rounds", [(8, 50, 100, 100, 0), (8, 50, 100, 100, 1), (10, 50, 100, 200, 3), (10, 10, 10, 50, 2),
                                                (10, 1, 1, 20, 1), (11, 50, 200, 1000, 0), (12, 50, 100, 1000, 3), (9, 400, 600, 300, 2),
                                                (10, 50, 100, 1, 1), (10, 50, 100, 7, 0)])
def test_bit_identical_to_oracle_on_rmat(scale, K, L, R, rounds):
    got, want = run_pair(G.rmat(scale), K, L, R, 0.85, rounds)
    assert_bit_identical(got, want, f"mc rmat{scale} K{K} L{L} R{R} rounds{rounds}")
    for k in ("walk_steps", "walks", "merged_entries", "truncations", "boundary_ties", "iterations_run"):
        assert got.stats[k] == want.stats[k], (k, got.stats[k], want.stats[k])


@pytest.mark.parametrize("hub", [4, 64])
def test_bit_identical_with_the_order_free_combine(hub):
    got, want = run_pair(G.rmat(11), 50, 100, 200, 0.85, 3, hub=hub)
    assert_bit_identical(got, want, f"mc hub>{hub}")


@pytest.mark.parametrize("damping", [0.05, 0.3, 0.99])
def test_damping_values(damping):
    """(d = 0 is degenerate in the reference itself: map[node] = 1/factor = inf, times factor = NaN, mccompletepathv2.h:214-247)"""
    got, want = run_pair(G.rmat(9), 20, 40, 100, damping, 2)
    assert_bit_identical(got, want, f"mc damping {damping}")
    assert got.stats["walk_steps"] == want.stats["walk_steps"]


def test_damping_one_hits_the_step_cap():
    """d = 1: the reference would loop forever on a cycle (mccompletepathv2.h:155); both sides cap a walk at 4096 hops"""
    g = G.from_edges(3, [0, 1, 2], [1, 2, 0])
    got, want = run_pair(g, 3, 3, 4, 1.0, 1)
    assert got.stats["walk_steps"] == want.stats["walk_steps"] == 3 * 4 * 4096
    assert_bit_identical(got, want, "d=1 ring")


@pytest.mark.parametrize("damping", [0.99, 1.0])
def test_high_damping_on_a_clique_with_self_loops_stays_in_range(damping):
    """Visits per walk reach 1/(1-d) (d = 1: the 4096-hop cap) -- far above the 16 that the order-free 2^-59 accumulator of
    the combine rounds can hold. With the DEFAULT hub threshold (out-degree 20 > 12 would be order-free) the one-shot call
    must route every node to the exact-order fp64 path: bit-identical to the oracle's exact-order result, no wrap-around."""
    n = 20
    src = np.repeat(np.arange(n), n)
    dst = np.tile(np.arange(n), n)          # complete digraph with self loops, out-degree 20
    g = G.from_edges(n, src, dst)
    got = ppr.mccompletepathv2_csr(g, 10, 20, 50, damping, seed=SEED, rounds=2, hub_threshold=0)
    want = ob.oracle_mc(g, 10, 20, 50, damping, SEED, 2, hub_threshold=0)
    assert_bit_identical(got, want, f"mc clique d={damping}")
    if damping == 1.0:
        assert got.scores.max() > 16.0       # the range the fixed-point words would have wrapped at
    assert np.isfinite(got.scores).all() and (got.scores >= 0).all()


def test_session_refuses_high_damping_when_it_holds_order_free_nodes():
    from approximated_personalized_pagerank_b200._lib import PprB200Error
    n = 20
    g = G.from_edges(n, np.repeat(np.arange(n), n), np.tile(np.arange(n), n))
    s = ppr.Session(g, max_L=20, hub_threshold=0)
    try:
        with pytest.raises(PprB200Error, match="fixed-point accumulator"):
            s.mc(10, 20, 50, 0.99, seed=SEED, rounds=2)
        s.mc(10, 20, 50, 0.85, seed=SEED, rounds=2)   # fine below the limit
        s.fetch()
    finally:
        s.close()
    s = ppr.Session(g, max_L=20, hub_threshold=ppr.NEVER_HUB)
    try:
        s.mc(10, 20, 50, 0.99, seed=SEED, rounds=2)
        s.fetch()
    finally:
        s.close()


def test_seed_changes_the_walks_and_runs_repeat():
    g = G.rmat(10)
    a = ppr.mccompletepathv2_csr(g, 50, 100, 100, 0.85, seed=1, rounds=0)
    b = ppr.mccompletepathv2_csr(g, 50, 100, 100, 0.85, seed=1, rounds=0)
    c = ppr.mccompletepathv2_csr(g, 50, 100, 100, 0.85, seed=2, rounds=0)
    assert_bit_identical(a, b, "same seed")
    assert (a.scores != c.scores).any()


def test_many_distinct_visits_use_the_fallback_table(monkeypatch):
    """a source whose visited set outgrows the shared-memory table is redone with a table in global memory"""
    monkeypatch.setenv("PPRB200_WALK_TCAP", "1024")
    g = G.rmat(12)
    got, want = run_pair(g, 50, 100, 1000, 0.85, 1)
    assert got.stats["walk_steps"] == want.stats["walk_steps"]
    assert_bit_identical(got, want, "mc fallback table")


def test_ragged_and_degenerate_graphs():
    for g in (G.from_edges(1, [], []), G.from_edges(1, [0], [0]), G.from_edges(2, [0, 1], [1, 0]), G.from_edges(10, [], []),
              G.from_edges(40, [0] * 39, list(range(1, 40))), G.from_edges(40, list(range(1, 40)), [0] * 39)):
        for rounds in (0, 1, 3):
            got, want = run_pair(g, 10, 30, 100, 0.85, rounds)
            assert_bit_identical(got, want, f"mc n={g.n} e={g.n_edges} rounds={rounds}")


# ---- the reference's own unit tests (test/mccompletepathv2Test.cc), through the dict API ----
def test_ref_no_edges_and_stars():
    res = ppr.mccompletepathv2({i: [] for i in range(10)}, 10, 30, 100, 0.85)                 # :38-50
    assert len(res) == 10 and all(len(res[i]) == 1 and res[i][i] == 1.0 for i in range(10))
    graph = {i: ([0] if i else []) for i in range(6)}                                        # :154-182 star
    res = ppr.mccompletepathv2(graph, 10, 30, 100, 0.85)
    assert len(res[0]) == 1 and res[0][0] == 1.0
    assert all(len(res[i]) == 2 and abs(res[i][0] - 0.85) < 10e-5 for i in range(1, 6))
    graph = {0: [1, 2, 3, 4, 5], **{i: [] for i in range(1, 6)}}                             # :184-219 reversed star
    res = ppr.mccompletepathv2(graph, 10, 30, 1000, 0.85)
    assert all(len(res[i]) == 1 for i in range(1, 6))
    assert all(abs(res[0][i] - 0.85 / 5) < 0.05 for i in range(1, 6))


def test_ref_ring_weakly_monotone():
    graph = {i: [(i + 1) % 6] for i in range(6)}                                             # :107-152
    res = ppr.mccompletepathv2(graph, 10, 30, 1000, 0.85)
    for i in range(6):
        assert len(res[i]) == 6
        for u in range(5):
            assert res[i][(i + u) % 6] >= res[i][(i + u + 1) % 6]


def _l1_restricted(res, g, sources, exact, K, d):
    """mean over sources of the L1 distance restricted to (exact top-K) U basket, MC scores scaled by (1-d)"""
    tot = 0.0
    for i, s in enumerate(sources):
        est = np.zeros(g.n)
        est[res.ids[s, :res.cnt[s]]] = res.scores[s, :res.cnt[s]] * (1.0 - d)
        top = np.argsort(-exact[i], kind="stable")[:K]
        keys = np.union1d(top, res.ids[s, :res.cnt[s]])
        tot += np.abs(est[keys] - exact[i][keys]).sum()
    return tot / len(sources)


@requires_ref
@pytest.mark.parametrize("R,L", [(1000, 100), (100, 100)])
def test_l1_error_no_worse_than_the_reference(R, L):
    """north-star criterion: per-source L1 error against exact power-iteration PPR <= the reference's at the same R"""
    g = G.rmat(12)
    rng = np.random.default_rng(7)
    nonsink = np.nonzero(g.out_degree() > 0)[0]
    sources = rng.choice(nonsink, 200, replace=False)
    exact = np.stack([ob.oracle_ppr(g, int(s), 100, 0.85, -1.0) for s in sources])
    got = ppr.mccompletepathv2_csr(g, 50, L, R, 0.85)
    ref = ob.ref_mc(g, 50, L, R, 0.85)
    e_gpu = _l1_restricted(got, g, sources, exact, 50, 0.85)
    e_ref = _l1_restricted(ref, g, sources, exact, 50, 0.85)
    print(f"mean restricted L1 (R={R}, L={L}): B200 {e_gpu:.4f}  reference {e_ref:.4f}")
    assert e_gpu <= e_ref * 1.02


def test_combine_rounds_with_hub_teams(monkeypatch):
    """the combine rounds run on the same merge kernels as GRank: hub teams (forced onto ordinary nodes) change nothing"""
    monkeypatch.setenv("PPRB200_TEAM_DEG", "129")
    monkeypatch.setenv("PPRB200_TEAM_CHUNK", "48")
    got, want = run_pair(G.rmat(12), 50, 100, 200, 0.85, 3, hub=8)
    assert_bit_identical(got, want, "mc with hub teams")
    assert got.stats["walk_steps"] == want.stats["walk_steps"]
