"""Comparison helpers shared by the parity tests."""
import numpy as np


def rows_as_dicts(ids, scores, cnt):
    out = []
    K = ids.shape[1]
    for v in range(ids.shape[0]):
        c = min(int(cnt[v]), K)
        out.append({int(ids[v, i]): float(scores[v, i]) for i in range(c)})
    return out


def assert_bit_identical(a, b, what=""):
    """a, b: results with ids/scores/cnt in the canonical layout (rows sorted (score desc, id asc))."""
    assert (a.cnt == b.cnt).all(), f"{what}: basket sizes differ at nodes {np.nonzero(a.cnt != b.cnt)[0][:10]}"
    assert (a.ids == b.ids).all(), f"{what}: basket membership/order differs at nodes {np.nonzero((a.ids != b.ids).any(1))[0][:10]}"
    sa = np.ascontiguousarray(a.scores).view(np.uint64)
    sb = np.ascontiguousarray(b.scores).view(np.uint64)
    assert (sa == sb).all(), f"{what}: score bits differ, max |d| = {np.abs(a.scores - b.scores).max():.3e}"


def compare_membership(a, b):
    """-> (#rows with different key sets, max |score difference| over common keys)"""
    da, db = rows_as_dicts(a.ids, a.scores, a.cnt), rows_as_dicts(b.ids, b.scores, b.cnt)
    mism, maxd = 0, 0.0
    for x, y in zip(da, db):
        if set(x) != set(y):
            mism += 1
        for k in set(x) & set(y):
            maxd = max(maxd, abs(x[k] - y[k]))
    return mism, maxd
