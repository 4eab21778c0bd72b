"""Shared helpers for the test-suite (test infrastructure)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
U32_MAX = 0xFFFFFFFF


def goldens():
    with open(os.path.join(HERE, "golden", "reference_goldens.json")) as f:
        return json.load(f)


def fixtures():
    return np.load(os.path.join(HERE, "golden", "ksparse_fixtures.npz"))


def dense_csr(costs):
    rows = [np.asarray(r, dtype=np.float64) for r in costs]
    row_ptr = np.zeros(len(rows) + 1, dtype=np.uint32)
    row_ptr[1:] = np.cumsum([len(r) for r in rows])
    cols = np.concatenate([np.arange(len(r), dtype=np.uint32) for r in rows])
    vals = np.concatenate(rows)
    return len(rows), max(len(r) for r in rows), row_ptr, cols, vals


def objective_of(row_ptr, cols, vals, p2o):
    """Sum of the chosen arcs' ORIGINAL values (independent of any solver's sign handling)."""
    total = 0.0
    for i, j in enumerate(p2o):
        if j == U32_MAX:
            continue
        a, b = int(row_ptr[i]), int(row_ptr[i + 1])
        hit = np.nonzero(cols[a:b] == j)[0]
        assert hit.size >= 1, f"person {i} assigned to object {j} it has no arc to"
        total += float(vals[a + hit[0]])
    return total


def check_matching(n_rows, n_cols, row_ptr, cols, p2o, o2p, num_unassigned):
    """A valid (partial) matching: mutually consistent vectors, arcs exist, count of unassigned persons agrees."""
    p2o = np.asarray(p2o, dtype=np.uint32)
    o2p = np.asarray(o2p, dtype=np.uint32)
    assert p2o.size == n_rows and o2p.size == n_cols
    assigned = np.nonzero(p2o != U32_MAX)[0]
    assert int(n_rows - assigned.size) == int(num_unassigned)
    assert np.all(p2o[assigned] < n_cols)
    assert np.array_equal(o2p[p2o[assigned]], assigned.astype(np.uint32))
    owned = np.nonzero(o2p != U32_MAX)[0]
    assert owned.size == assigned.size
    assert np.array_equal(p2o[o2p[owned]], owned.astype(np.uint32))
    # every chosen arc exists in the person's row
    starts = np.asarray(row_ptr[:-1], dtype=np.int64)
    counts = np.diff(np.asarray(row_ptr, dtype=np.int64))
    row_of = np.repeat(np.arange(n_rows), counts)
    hit = (np.asarray(cols, dtype=np.int64) == p2o.astype(np.int64)[row_of])
    has = np.zeros(n_rows, dtype=bool)
    has[row_of[hit]] = True
    assert np.all(has[assigned])
    del starts


def random_sparse_instance(rng, n, m, k, integer=True, lo=0, hi=100):
    """k distinct sorted columns per row + a planted diagonal-ish perfect matching when n <= m."""
    row_ptr = np.arange(0, n * k + 1, k, dtype=np.uint32)
    cols = np.empty(n * k, dtype=np.uint32)
    perm = rng.permutation(m)[:n]
    for i in range(n):
        others = rng.choice(m - 1, size=k - 1, replace=False)
        others = others + (others >= perm[i])
        cols[i * k:(i + 1) * k] = np.sort(np.concatenate([[perm[i]], others]))
    if integer:
        vals = rng.integers(lo, hi, size=n * k).astype(np.float64)
    else:
        vals = rng.uniform(lo, hi, size=n * k)
    return row_ptr, cols, vals


def scipy_optimum(n, m, row_ptr, cols, vals, maximize=False):
    """Independent optimum of a feasible sparse instance (scipy's LAPJVsp on the biadjacency matrix)."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import min_weight_full_bipartite_matching
    v = np.asarray(vals, dtype=np.float64)
    shift = 1.0 - v.min() if not maximize else 1.0 + v.max()
    w = (v + shift) if not maximize else (shift - v)     # strictly positive weights, explicit zeros are dropped by scipy
    mat = csr_matrix((w, np.asarray(cols, dtype=np.int64), np.asarray(row_ptr, dtype=np.int64)), shape=(n, m))
    r, c = min_weight_full_bipartite_matching(mat)
    total = 0.0
    for i, j in zip(r, c):
        a, b = int(row_ptr[i]), int(row_ptr[i + 1])
        total += float(v[a + np.nonzero(cols[a:b] == j)[0][0]])
    return total
