"""Regenerates tests/golden/eval_kendall_jaccard.npz (known answers of the reference's kendallCorrelation and jaccard,
kendall.h:22-180 / pprInternal.h:174-186) and tests/golden/eval_example_head.npz (pprSingleSource on the first edges of
the reference's example.txt) by running the UNMODIFIED reference through oracle/_ref/libppr_ref.so in the CPU container.

    python tests/golden/make_golden_eval.py
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import oracle_bindings as ob  # noqa: E402
from approximated_personalized_pagerank_b200 import evaluate as EV  # noqa: E402


def main():
    rng = np.random.default_rng(7)
    xs, ys, taus = [], [], []
    for n, ties in ((1, 0), (2, 0), (2, 2), (5, 0), (17, 0), (50, 0), (50, 3), (50, 10), (100, 5), (200, 0), (64, 64)):
        for _ in range(3):
            x = rng.random(n)
            y = x + 0.3 * rng.standard_normal(n)
            if ties:
                x = np.round(x * ties) / ties
                y = np.round(y * ties) / ties
            if ties == n:
                y = np.full(n, 0.5)
            xs.append(x); ys.append(y); taus.append(ob.ref_kendall(x, y))
    sets_a, sets_b, jac = [], [], []
    for na, nb in ((0, 0), (0, 3), (5, 5), (50, 50), (50, 20), (100, 100)):
        a = rng.choice(200, size=na, replace=False).astype(np.int32)
        b = rng.choice(200, size=nb, replace=False).astype(np.int32)
        sets_a.append(a); sets_b.append(b); jac.append(ob.ref_jaccard(a, b))
    np.savez_compressed(HERE / "eval_kendall_jaccard.npz", n_k=len(xs), n_j=len(jac), tau=np.array(taus), jac=np.array(jac),
                        **{f"x{i}": x for i, x in enumerate(xs)}, **{f"y{i}": y for i, y in enumerate(ys)},
                        **{f"a{i}": a for i, a in enumerate(sets_a)}, **{f"b{i}": b for i, b in enumerate(sets_b)})
    # the head of the reference's own dataset: ingest + exact PPR
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
        f.write("".join(open("/root/reference/example.txt").readlines()[:4000]))
    g = EV.import_graph_csv(f.name)
    src = np.flatnonzero(g.out_degree() > 0)[:24].astype(np.int32)
    exact = ob.ref_ppr(g, src, 100, 0.85, 1e-4)
    np.savez_compressed(HERE / "eval_example_head.npz", row_ptr=g.row_ptr, col=g.col, keys=np.array(g.keys), sources=src, exact=exact)
    print("kendall cases", len(xs), "jaccard cases", len(jac), "example head: nodes", g.n, "edges", g.n_edges)


if __name__ == "__main__":
    main()
