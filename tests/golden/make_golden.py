"""Regenerates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libppr_ref.so, compiled from
/root/reference by oracle/Makefile) in the CPU container. The GPU box has no /root/reference: the parity tests
there read these committed fixtures.

    python tests/golden/make_golden.py

Each fixture stores the graph itself (CSR over the reference's int keys), the parameters, the reference's map
iteration order and partition colours (both implementation-defined) and the reference's
baskets in key space with rows sorted (score desc, key asc).
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import oracle_bindings as ob  # noqa: E402
from approximated_personalized_pagerank_b200 import graphs as G  # noqa: E402


def star(n=6, self_loop=False):
    src = list(range(1, n)) + ([0] if self_loop else [])
    dst = [0] * (n - 1) + ([0] if self_loop else [])
    return G.from_edges(n, src, dst)


def in_out_star(n=100, self_loop=True, out_star=True):
    src = list(range(n - 1)) + ([0] if self_loop else []) + ([0] * (n - 1) if out_star else [])
    dst = [0] * (n - 1) + ([0] if self_loop else []) + (list(range(n - 1)) if out_star else [])
    return G.from_edges(n, src, dst)


def complete(n=100):
    src = np.repeat(np.arange(n), n)
    dst = np.tile(np.arange(n), n)
    return G.from_edges(n, src, dst)


def random_multi(n, e, seed):
    rng = np.random.default_rng(seed)
    return G.from_edges(n, rng.integers(0, n, e), rng.integers(0, n, e))


CASES = {
    # name: (graph, K, L, iterations, damping, tolerance)
    "ring100_config1": (G.ring(100), 50, 100, 30, 0.85, 1e-3),           # BASELINE config 1, README.md:105-115
    "ring100_k10_l10": (G.ring(100), 10, 10, 100, 0.85, 1e-4),           # test/grankTest.cc:184-200
    "ring100_k10_l20": (G.ring(100), 10, 20, 100, 0.85, 1e-4),           # :202-216
    "ring100_full": (G.ring(100), 100, 100, 100, 0.85, -1.0),            # :237-259, :285-302
    "ring100_k200": (G.ring(100), 200, 200, 100, 0.85, -1.0),            # :261-283
    "ring6_k3_l4": (G.ring(6), 3, 4, 100, 0.85, 1e-4),                   # :107-152
    "ring6_k3_l3": (G.ring(6), 3, 3, 100, 0.85, 1e-4),
    "star6": (star(6, False), 10, 30, 100, 0.85, 1e-4),                  # :154-172
    "star6_selfloop": (star(6, True), 10, 30, 100, 0.85, 1e-4),          # :174-182
    "instar100": (in_out_star(100, False, False), 100, 100, 100, 0.85, -1.0),   # :304-318
    "instar100_selfloop": (in_out_star(100, True, False), 100, 100, 100, 0.85, -1.0),
    "inoutstar100": (in_out_star(100, True, True), 100, 100, 100, 0.85, -1.0),  # :330-341
    "random100_5000": (random_multi(100, 5000, 7), 100, 100, 100, 0.85, -1.0),  # :343-361
    "complete100": (complete(100), 100, 100, 100, 0.85, -1.0),           # :363-379
    "random200_tol": (random_multi(200, 1500, 11), 200, 200, 60, 0.85, 1e-3),   # grankMultiThreadTest.cc:384-479 (positive tolerance)
    "rmat8_full": (G.rmat(8), 256, 256, 30, 0.85, -1.0),
    "rmat10_k50_l2000": (G.rmat(10), 50, 2000, 30, 0.85, 1e-3),          # SURVEY.md 8c probe: tie-free, 21 iterations
    "noedges10": (G.from_edges(10, [], []), 10, 30, 100, 0.85, 1e-4),    # :38-50
    "selfloop1": (G.from_edges(1, [0], [0]), 10, 30, 100, 0.85, 1e-4),   # :70-84
}


def main():
    for name, (g, K, L, it, d, tol) in CASES.items():
        r = ob.ref_grank(g, K, L, it, d, tol)
        rm = ob.ref_grank(g, K, L, it, d, tol, nthreads=4)
        assert (r.ids == rm.ids).all() and (r.scores == rm.scores).all(), name  # grankMulti == grank
        order = ob.ref_iteration_order(g)
        colour = ob.ref_find_partitions(g)
        np.savez_compressed(HERE / f"grank_{name}.npz", row_ptr=g.row_ptr, col=g.col, K=K, L=L, iterations=it, damping=d,
                            tolerance=tol, order=order, colour=colour, ids=r.ids, scores=r.scores, cnt=r.cnt)
        print(name, "n", g.n, "max cnt", int(r.cnt.max()))
    # exact PPR (pprSingleSource.h, 100 it, tol -1) for the MC L1 criterion and the GRank == PPR tests
    g = G.rmat(10)
    srcs = np.array([v for v in range(g.n) if g.row_ptr[v + 1] > g.row_ptr[v]][:64], dtype=np.int32)
    ppr = ob.ref_ppr(g, srcs)
    np.savez_compressed(HERE / "ppr_rmat10.npz", row_ptr=g.row_ptr, col=g.col, sources=srcs, ppr=ppr)
    g = G.ring(100)
    np.savez_compressed(HERE / "ppr_ring100.npz", row_ptr=g.row_ptr, col=g.col, sources=np.arange(100, dtype=np.int32),
                        ppr=ob.ref_ppr(g, np.arange(100)))
    print("ppr fixtures written")


if __name__ == "__main__":
    main()
