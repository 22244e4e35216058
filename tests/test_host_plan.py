"""The host front half of a session (pprb200_debug_host_plan: colouring, storage order, rank labels, CSR encode, work
items) checked on the CPU against its specification (device_common.cuh "HBM layout"): no device involved."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import approximated_personalized_pagerank_b200 as ppr  # noqa: E402
from approximated_personalized_pagerank_b200 import _lib, graphs as G  # noqa: E402
import oracle_bindings as ob  # noqa: E402

COL_SINK = 0x80000000
COL_POS_MASK = 0x3FFFFFFF


def host_plan(g, colour=None, hub=0, rank=0, world=1):
    lib = _lib.load()
    n, E = g.n, g.n_edges
    pos_of = np.zeros(max(n, 1), np.int32)
    rank_of = np.zeros(max(n, 1), np.int32)
    row_off = np.zeros(n + 1, np.int64)
    enc = np.zeros(max(E, 1), np.uint32)
    cap = n + E // 32 + 16
    item_pos, item_off, item_len = np.zeros(cap, np.int32), np.zeros(cap, np.int64), np.zeros(cap, np.int32)
    summary = np.zeros(16, np.int32)
    col_arr = None if colour is None else np.ascontiguousarray(colour, np.uint8)
    _lib.check(lib.pprb200_debug_host_plan(_lib.ptr(g.row_ptr), _lib.ptr(g.col), n, _lib.ptr(col_arr), hub, rank, world,
                                           _lib.ptr(pos_of), _lib.ptr(rank_of), _lib.ptr(row_off), _lib.ptr(enc), _lib.ptr(item_pos),
                                           _lib.ptr(item_off), _lib.ptr(item_len), cap, _lib.ptr(summary)))
    M, n_items = int(summary[0]), int(summary[1])
    assert n_items <= cap
    return dict(M=M, n_items=n_items, chunk=int(summary[2]), mid_deg=int(summary[3]), range_begin=summary[4:6].copy(),
                range_end=summary[6:8].copy(), item_begin=summary[8:12].reshape(2, 2).copy(), item_end=summary[12:16].reshape(2, 2).copy(),
                pos_of=pos_of[:n], rank_of=rank_of[:n], row_off=row_off[:M + 1], enc=enc[:E], item_pos=item_pos[:n_items],
                item_off=item_off[:n_items], item_len=item_len[:n_items])


def check_plan(g, hub):
    colour = ob.oracle_find_partitions(g)
    P = host_plan(g, None, hub)
    n = g.n
    deg = g.out_degree()
    hub_eff = hub if hub else ppr.DEFAULT_HUB_THRESHOLD
    # --- storage positions: exactly the non-sink nodes, ordered by (colour, class, out-degree desc, dense id asc)
    nonsink = np.flatnonzero(deg > 0)
    assert P["M"] == len(nonsink)
    assert (P["pos_of"][deg == 0] == -1).all()
    order = np.empty(P["M"], np.int64)
    order[P["pos_of"][nonsink]] = nonsink  # also proves the positions are a permutation of 0..M-1
    assert sorted(P["pos_of"][nonsink].tolist()) == list(range(P["M"]))
    cls = np.where(deg <= hub_eff, 0, np.where(deg <= P["mid_deg"], 1, 2))
    key = [(int(colour[v]), int(cls[v]), -int(deg[v]), int(v)) for v in order]
    assert key == sorted(key)
    # --- rank labels: in-degree descending, ties by dense id
    indeg = np.bincount(g.col, minlength=n)
    by_rank = np.empty(n, np.int64)
    by_rank[P["rank_of"]] = np.arange(n)
    k2 = [(-int(indeg[v]), int(v)) for v in by_rank]
    assert k2 == sorted(k2)
    # --- CSR in storage order, column words decode to the caller's successors in order
    assert (np.diff(P["row_off"]) == deg[order]).all()
    for p in np.random.default_rng(0).choice(P["M"], size=min(P["M"], 300), replace=False) if P["M"] else []:
        v = order[p]
        succ = g.col[g.row_ptr[v]:g.row_ptr[v + 1]]
        words = P["enc"][P["row_off"][p]:P["row_off"][p + 1]]
        for s_, w in zip(succ, words):
            if deg[s_] == 0:
                assert int(w) == (COL_SINK | int(P["rank_of"][s_]))
            else:
                assert int(w) & COL_POS_MASK == P["pos_of"][s_] and (int(w) >> 30) == int(colour[s_])
    # --- exact-order ranges and order-free work items tile their classes
    for c in (0, 1):
        in_range = order[P["range_begin"][c]:P["range_end"][c]]
        assert ((colour[in_range] == c) & (cls[in_range] == 0)).all()
        assert len(in_range) == int(((colour[nonsink] == c) & (cls[nonsink] == 0)).sum())
        for k in (0, 1):  # 0 = mid, 1 = big
            b, e = P["item_begin"][c][k], P["item_end"][c][k]
            covered = {}
            for i in range(b, e):
                p, off, ln = int(P["item_pos"][i]), int(P["item_off"][i]), int(P["item_len"][i])
                v = order[p]
                assert colour[v] == c and cls[v] == k + 1 and 0 < ln <= P["chunk"]
                assert off == P["row_off"][p] + covered.get(p, 0)  # chunks of a node follow each other
                covered[p] = covered.get(p, 0) + ln
            want = {int(P["pos_of"][v]): int(deg[v]) for v in nonsink if colour[v] == c and cls[v] == k + 1}
            assert covered == want
    return P


@pytest.mark.parametrize("scale,hub", [(8, 0), (10, 4), (12, 0), (13, 30)])
def test_host_plan_on_rmat(scale, hub):
    check_plan(G.rmat(scale), hub)


def test_host_plan_on_degenerate_graphs():
    check_plan(G.from_edges(7, [], []), 0)                               # only sinks
    check_plan(G.ring(50), 0)
    check_plan(G.from_edges(300, [0] * 5000, np.arange(5000) % 300), 2)  # one hub above the default chunk: several items
    rng = np.random.default_rng(2)
    check_plan(G.from_edges(500, rng.integers(0, 500, 6000), rng.integers(0, 20, 6000)), 3)  # few targets: in-degree ties


def test_host_plan_with_degrees_around_the_counting_sort_cap():
    """out- and in-degrees at and above 4095 share one counting-sort key and are ordered by a comparison sort afterwards:
    4094 / 4095 / 4096 / 6000 successors, two hubs with the same degree (tie by id), and targets with in-degree >= 4095"""
    n = 7000
    src, dst = [], []
    for v, d in ((5, 4094), (9, 4095), (2, 4096), (11, 6000), (3, 6000), (40, 4095)):
        src += [v] * d
        dst += list(range(100, 100 + d))
    rng = np.random.default_rng(4)
    for t, d in ((60, 4095), (61, 4094), (62, 5000), (63, 5000)):   # in-degrees around the cap
        src += rng.integers(1000, n, d).tolist()
        dst += [t] * d
    g = G.from_edges(n, src, dst)
    for hub in (0, 5000, ppr.NEVER_HUB):
        check_plan(g, hub)


def test_host_plan_shards_partition_the_single_gpu_plan():
    g = G.rmat(12)
    colour = ppr.find_partitions_csr(g)
    P1 = host_plan(g, colour)
    world = 4
    seen_items, n_seq = [], 0
    for r in range(world):
        P = host_plan(g, colour, 0, r, world)
        assert (P["pos_of"] == P1["pos_of"]).all() and (P["enc"] == P1["enc"]).all()  # every rank holds the whole graph
        seen_items += list(zip(P["item_pos"].tolist(), P["item_off"].tolist(), P["item_len"].tolist()))
        n_seq += int((P["range_end"] - P["range_begin"]).sum())
    # the union of the ranks' items covers every order-free node exactly once (chunk sizes may differ with world)
    cover = {}
    for p, off, ln in seen_items:
        cover[p] = cover.get(p, 0) + ln
    cover1 = {}
    for p, ln in zip(P1["item_pos"].tolist(), P1["item_len"].tolist()):
        cover1[p] = cover1.get(p, 0) + ln
    assert cover == cover1
    assert n_seq == int((P1["range_end"] - P1["range_begin"]).sum())
    owner = np.zeros(g.n, np.int32)
    _lib.check(_lib.load().pprb200_shard_owner(_lib.ptr(g.row_ptr), _lib.ptr(g.col), g.n, _lib.ptr(colour), 0, world, _lib.ptr(owner)))
    for r in range(world):
        P = host_plan(g, colour, 0, r, world)
        mine = set(np.flatnonzero(owner == r).tolist())
        order = np.empty(P["M"], np.int64)
        nonsink = np.flatnonzero(g.out_degree() > 0)
        order[P["pos_of"][nonsink]] = nonsink
        assert {int(order[p]) for p in set(P["item_pos"].tolist())} <= mine


def test_host_plan_does_not_depend_on_the_thread_count():
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import hashlib, numpy as np;"
            "from approximated_personalized_pagerank_b200 import graphs as G; import test_host_plan as T;"
            "P = T.host_plan(G.rmat(15));"
            "h = hashlib.sha256(); [h.update(np.ascontiguousarray(P[k]).tobytes()) for k in ('pos_of','rank_of','row_off','enc','item_pos','item_off','item_len')];"
            "sys.stdout.write(h.hexdigest())" % (str(ROOT), str(ROOT / "tests")))
    outs = set()
    for t in ("1", "5", "16"):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PPRB200_HOST_THREADS=t))
        assert r.returncode == 0, r.stderr[-1500:]
        outs.add(r.stdout.strip())
    assert len(outs) == 1, outs
