"""SURVEY.md 8d parity check (iii) / 7-1: the GPU against the LIVE reference on R-MAT with truncation, next to the
boundary-tie count and the reference's own self-noise band.

The reference cuts ties arbitrarily (nth_element over hash order, /root/reference/include/internal/pprInternal.h:115-119)
and the choice feeds later iterations, so on heavy-tailed graphs "identical to the reference" is only defined up to the
band inside which the reference moves against ITSELF when nothing but std::hash changes (oracle/ref_shim.cc:
ref_grank_althash). This test measures both and writes the numbers to gpurun_out/parity_report.json (committed as
profiles/r2/parity_report.json, quoted by bench.py); it asserts that the GPU -- default configuration (order-free above
out-degree 12) and exact-order configuration -- stays inside that band."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle_bindings as ob
import approximated_personalized_pagerank_b200 as ppr
from approximated_personalized_pagerank_b200 import graphs as G
from helpers import rows_as_dicts

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ob.have_ref(), reason="needs oracle/_ref (the reference build)")]
ROOT = Path(__file__).resolve().parent.parent


def diff(a, b):
    """rows with different key sets, differing members in total, max |score difference| over common keys, mean Jaccard"""
    da, db = rows_as_dicts(a.ids, a.scores, a.cnt), rows_as_dicts(b.ids, b.scores, b.cnt)
    rows = members = 0
    maxd, jac = 0.0, []
    for x, y in zip(da, db):
        sx, sy = set(x), set(y)
        if sx != sy:
            rows += 1
            members += len(sx ^ sy) // 2
        jac.append(len(sx & sy) / max(1, len(sx | sy)))
        for k in sx & sy:
            maxd = max(maxd, abs(x[k] - y[k]))
    return {"sources_with_different_membership": rows, "differing_members": members, "max_abs_score_delta": maxd,
            "mean_jaccard": float(np.mean(jac))}


@pytest.mark.parametrize("scale", [10, 12])
def test_gpu_vs_live_reference_next_to_the_reference_self_noise(scale):
    K, L, it, d, tol = 50, 100, 30, 0.85, 1e-3
    g = G.rmat(scale)
    gd, order = ob.to_reference_space(g)      # dense id = position in the reference's map iteration order
    ref = ob.ref_grank(g, K, L, it, d, tol)   # key space
    alt = ob.ref_grank_althash(g, K, L, it, d, tol)
    band = diff(ref, alt)
    out = {"graph": f"R-MAT scale {scale}", "K": K, "L": L, "iterations": it, "tolerance": tol, "sources": g.n,
           "reference_vs_reference_other_hash": band}
    for name, hub in (("gpu_default", 0), ("gpu_exact_order", ppr.NEVER_HUB)):
        got = ppr.grank_csr(gd, K, L, it, d, tol, colour=None, hub_threshold=hub)
        r = diff(ob.baskets_to_keyspace(got, order), ref)
        r["boundary_ties"] = got.stats["boundary_ties"]
        r["truncations"] = got.stats["truncations"]
        r["iterations_run"] = got.stats["iterations_run"]
        out[name + "_vs_reference"] = r
    want = ob.oracle_grank(gd, K, L, it, d, tol, hub_threshold=ppr.DEFAULT_HUB_THRESHOLD)
    out["oracle_iterations_run"] = want.stats["iterations_run"]
    p = ROOT / "gpurun_out" / "parity_report.json"
    p.parent.mkdir(exist_ok=True)
    allr = json.loads(p.read_text()) if p.exists() else {}
    allr[f"rmat{scale}"] = out
    p.write_text(json.dumps(allr, indent=1))
    print(json.dumps(out))
    # inside the band: no more disagreement with the reference than the reference has with itself (x1.5 + slack)
    for name in ("gpu_default_vs_reference", "gpu_exact_order_vs_reference"):
        assert out[name]["sources_with_different_membership"] <= 1.5 * band["sources_with_different_membership"] + 8, (name, out)
        assert out[name]["max_abs_score_delta"] <= 2.0 * band["max_abs_score_delta"] + 1e-9, (name, out)
        assert out[name]["mean_jaccard"] >= band["mean_jaccard"] - 0.01
