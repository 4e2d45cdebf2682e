"""The k-nearest specification (csrc/knn.cu header; the reference exports the names and defines nothing):
oracle-level properties -- exactly k neighbours, the k smallest distances, ties towards the smaller index,
selection from a covering r-ball table equals selection from all pairs, mutual = union with the transpose."""
import numpy as np

import fixtures as fx


def test_knn_brute_properties(orc):
    V = fx.uniform_samples(400, 2, 5)
    k = 12
    cp, rv, nz = orc.knn_brute(V, k)
    assert np.array_equal(np.diff(cp), np.full(400, k))
    for q in (0, 17, 399):
        rows = rv[cp[q] - 1:cp[q + 1] - 1] - 1
        d = np.sqrt(((V[q] - V) ** 2).sum(1))
        d[q] = np.inf
        assert np.all(np.diff(rows) > 0) and q not in rows
        assert np.max(d[rows]) <= np.partition(d, k - 1)[k - 1] + 1e-15       # the k smallest
        assert np.allclose(nz[cp[q] - 1:cp[q + 1] - 1], d[rows], rtol=1e-15)


def test_ties_go_to_the_smaller_index(orc):
    g = np.stack(np.meshgrid(np.arange(7.0), np.arange(7.0)), -1).reshape(-1, 2) / 8.0     # lattice: many equal distances
    cp, rv, nz = orc.knn_brute(g, 3)
    q = 24                                                     # centre point (3, 3): four neighbours at distance 1/8
    rows = rv[cp[q] - 1:cp[q + 1] - 1] - 1
    assert rows.tolist() == [17, 23, 25]                       # of {17, 23, 25, 31} the three smaller indices


def test_selection_from_a_covering_ball_table(orc):
    V = fx.uniform_samples(600, 3, 9)
    k = 10
    full = orc.knn_brute(V, k)
    ball = orc.rball_brute(V, 0.35)                            # every column holds > k entries at this radius
    assert np.diff(ball[0]).min() >= k
    sel = orc.knn_of_table(*ball, k)
    assert all(np.array_equal(a, b) for a, b in zip(full[:2], sel[:2])) and full[2].tobytes() == sel[2].tobytes()


def test_mutual_is_union_with_transpose(orc):
    V = fx.uniform_samples(300, 2, 3)
    K = orc.knn_brute(V, 6)
    M = orc.union_transpose(K, K, 300)
    sets = [set(K[1][K[0][v] - 1:K[0][v + 1] - 1]) for v in range(300)]
    for v in (0, 5, 150, 299):
        exp = set(sets[v]) | {w + 1 for w in range(300) if (v + 1) in sets[w]}
        got = M[1][M[0][v] - 1:M[0][v + 1] - 1]
        assert sorted(exp) == got.tolist()
    # the mutual relation is symmetric
    pairs = {(v, int(w) - 1) for v in range(300) for w in M[1][M[0][v] - 1:M[0][v + 1] - 1]}
    assert all((w, v) in pairs for v, w in pairs)
