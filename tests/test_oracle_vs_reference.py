"""Pins the oracle (oracle/ppr_oracle.c): against the committed golden vectors produced by the unmodified
reference (tests/golden/make_golden.py), against the reference's closed forms / known answers, and -- where
oracle/_ref is built -- against the reference itself on fresh inputs. CPU only."""
import numpy as np
import pytest

import oracle_bindings as ob
from approximated_personalized_pagerank_b200 import graphs as G
from conftest import golden_cases, load_golden, requires_ref
from helpers import compare_membership


@pytest.mark.parametrize("name", golden_cases("grank"))
def test_oracle_matches_reference_golden_bitwise(name):
    """tie-free fixtures: canonical restatement == reference, max |d| = 0"""
    g, z = load_golden(name)
    order = z["order"]
    gd = g.relabel(order)
    colour_dense = ob.oracle_find_partitions(gd)
    assert (colour_dense == z["colour"][order]).all(), "findPartitions differs from the reference's"
    o = ob.oracle_grank(gd, int(z["K"]), int(z["L"]), int(z["iterations"]), float(z["damping"]), float(z["tolerance"]),
                        colour=colour_dense)
    ok = ob.baskets_to_keyspace(o, order)
    assert (ok.cnt == np.minimum(z["cnt"], int(z["K"]))).all()
    mism, maxd = compare_membership(ok, ob.Result(z["ids"], z["scores"], z["cnt"]))
    assert mism == 0
    assert maxd == 0.0


def test_ring_closed_form():
    """BASELINE config 1 (README.md:105-115): basket of s holds (s+j)%100 -> 0.15*0.85^j, last entry 0.85^(size-1);
    sizes 32 (partitions.second) / 31 (first). SURVEY.md 8c."""
    g, z = load_golden("grank_ring100_config1")
    order = z["order"]
    gd = g.relabel(order)
    colour = ob.oracle_find_partitions(gd)
    o = ob.baskets_to_keyspace(ob.oracle_grank(gd, 50, 100, 30, 0.85, 1e-3, colour=colour), order)
    for s in range(100):
        size = int(o.cnt[s])
        assert size == (31 if z["colour"][s] == 0 else 32)
        got = {int(k): float(v) for k, v in zip(o.ids[s, :size], o.scores[s, :size])}
        for j in range(size):
            want = 0.15 * 0.85 ** j if j < size - 1 else 0.85 ** (size - 1)
            assert abs(got[(s + j) % 100] - want) <= 4e-16 * max(want, 1e-3)


def test_grank_equals_ppr_golden():
    """test/grankTest.cc:285-302: K=L=N, tol -1, 100 iterations == pprSingleSource within 10e-5"""
    g, z = load_golden("ppr_ring100")
    colour = ob.oracle_find_partitions(g)
    o = ob.oracle_grank(g, 100, 100, 100, 0.85, -1.0, colour=colour)
    for i, s in enumerate(z["sources"]):
        dense = np.zeros(100)
        dense[o.ids[s, :o.cnt[s]]] = o.scores[s, :o.cnt[s]]
        assert np.abs(dense - z["ppr"][i]).max() < 10e-5


def test_oracle_ppr_matches_reference_ppr_golden():
    g, z = load_golden("ppr_rmat10")
    for i, s in enumerate(z["sources"][:16]):
        mine = ob.oracle_ppr(g, int(s))
        assert np.abs(mine - z["ppr"][i]).max() < 1e-12


def test_known_answers():
    """test/grankTest.cc:38-50 (no edges -> {i:0.15}), :70-84 (self loop -> 1.0), :154-182 (star)"""
    g = G.from_edges(10, [], [])
    o = ob.oracle_grank(g, 10, 30, 100, 0.85, 1e-4)
    assert (o.cnt == 1).all() and (o.ids[:, 0] == np.arange(10)).all() and np.allclose(o.scores[:, 0], 0.15, atol=1e-15)
    g = G.from_edges(1, [0], [0])
    o = ob.oracle_grank(g, 10, 30, 100, 0.85, 1e-4)
    assert o.cnt[0] == 1 and abs(o.scores[0, 0] - 1.0) < 10e-5
    g = G.from_edges(6, [1, 2, 3, 4, 5], [0] * 5)
    o = ob.oracle_grank(g, 10, 30, 100, 0.85, 1e-4)
    assert o.cnt[0] == 1 and abs(o.scores[0, 0] - 0.15) < 1e-15
    for i in range(1, 6):
        row = dict(zip(o.ids[i, :2], o.scores[i, :2]))
        assert o.cnt[i] == 2 and abs(row[0] - 0.15 * 0.85) < 10e-5


def test_fixed_point_hub_path_is_close_to_exact_order():
    """nodes above hub_threshold accumulate order-free in 2^-62 fixed point: same membership away from ties,
    |d| <= outdeg * 2^-62"""
    g = G.rmat(9)
    a = ob.oracle_grank(g, 512, 512, 20, 0.85, -1.0, hub_threshold=0)
    b = ob.oracle_grank(g, 512, 512, 20, 0.85, -1.0, hub_threshold=8)
    mism, maxd = compare_membership(a, b)
    assert mism == 0 and maxd < 1e-13


@requires_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_matches_live_reference_no_truncation(seed):
    """fresh random multigraphs, K=L=N (keepTop never cuts) -> bit-identical to the reference, same partitions"""
    rng = np.random.default_rng(seed)
    n = 150
    g = G.from_edges(n, rng.integers(0, n, 1200), rng.integers(0, n, 1200))
    gd, order = ob.to_reference_space(g)
    colour = ob.oracle_find_partitions(gd)
    assert (colour == ob.ref_find_partitions(g)[order]).all()
    r = ob.ref_grank(g, n, n, 40, 0.85, 1e-6)
    o = ob.baskets_to_keyspace(ob.oracle_grank(gd, n, n, 40, 0.85, 1e-6, colour=colour), order)
    mism, maxd = compare_membership(o, r)
    assert mism == 0 and maxd == 0.0


@requires_ref
def test_reference_primitives_keep_top_and_norm1():
    """keepTop (test/internal/keepTopTest.cc:42-69) and norm1 (norm1Test.cc) of the reference vs the conventions
    the oracle uses"""
    import ctypes as C
    ids = np.arange(501, dtype=np.int32)
    for L in (0, 1, 7, 250, 500, 501, 600):
        i2 = ids.copy()
        s2 = ids.astype(np.float64).copy()
        kept = ob.ref().ref_keep_top(C.c_uint32(L), ob.P(i2), ob.P(s2), C.c_int32(501))
        assert kept == min(L, 501)
        assert sorted(i2[:kept]) == list(range(501 - kept, 501))  # the L largest scores survive
    a_i = np.array([1, 2, 3], dtype=np.int32); a_s = np.array([0.5, 0.25, 0.125])
    b_i = np.array([2, 3, 4], dtype=np.int32); b_s = np.array([0.25, 0.5, 1.0])
    d = ob.ref().ref_norm1(ob.P(a_i), ob.P(a_s), C.c_int32(3), ob.P(b_i), ob.P(b_s), C.c_int32(3))
    assert d == 0.5 + 0.0 + 0.375 + 1.0
