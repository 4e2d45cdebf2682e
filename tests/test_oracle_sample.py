"""Oracle for batched free-state sampling (oracle/sample.c): the candidate stream is a pure function of
(seed, candidate number), the accepted set is exactly the free candidates in order."""
import numpy as np

import fixtures as fx


def test_candidate_stream_is_counter_based_and_uniform(orc):
    S = orc.StateSpace([-1.0, 2.0, 0.0], [1.0, 5.0, 0.5])
    a = np.array([orc.sample_candidate(S, 11, c) for c in range(4000)])
    b = np.array([orc.sample_candidate(S, 11, c) for c in range(3999, -1, -1)])[::-1]
    assert a.tobytes() == b.tobytes()                                  # order of evaluation does not matter
    assert np.all(a > S.lo) and np.all(a < S.hi)
    assert np.allclose(a.mean(0), (S.lo + S.hi) / 2, atol=0.05 * (S.hi - S.lo).max())
    assert not np.array_equal(a, np.array([orc.sample_candidate(S, 12, c) for c in range(4000)]))
    # coordinates 0 and 1 share one Philox block, coordinate 2 uses the next: known answer through the raw generator
    out = orc.philox4x32_10([5, 0, 0, 0x53414D50], [11, 0])
    u = ((out[0] >> 5) * 67108864.0 + (out[1] >> 6) + 0.5) / 9007199254740992.0
    assert orc.sample_candidate(S, 11, 5)[0] == S.lo[0] + u * (S.hi[0] - S.lo[0])


def test_sample_free_keeps_exactly_the_free_candidates_in_order(orc):
    O = orc.Obstacles2D(fx.ISRR_2H, fixed_point_test=True)
    S = orc.StateSpace([0, 0], [1, 1])
    V, used = orc.sample_free(O, S, 3000, 99)
    cand = np.array([orc.sample_candidate(S, 99, c) for c in range(used)])
    free = orc.states_free(O, S, cand)
    assert free[-1] and free.sum() == 3000                              # stops right at the N-th free candidate
    assert V.tobytes() == cand[free].tobytes()
    assert 0.7 < 3000 / used < 0.85                                     # ISRR_2H leaves ~78% of the square free
    # a budget of candidates that is too small returns what was found
    V2, used2 = orc.sample_free(O, S, 3000, 99, max_candidates=100)
    assert used2 == 100 and len(V2) == free[:100].sum()


def test_morton_order_is_a_stable_spatial_sort_of_the_same_set(orc):
    O = orc.Obstacles2D(fx.ISRR_2H, fixed_point_test=True)
    S = orc.StateSpace([0, 0], [1, 1])
    V, used = orc.sample_free(O, S, 5000, 5)
    M, used_m = orc.sample_free(O, S, 5000, 5, order=1)
    assert used_m == used
    keys = np.array([orc.morton_key(S, v) for v in V], dtype=np.uint64)
    assert M.tobytes() == V[np.argsort(keys, kind="stable")].tobytes()
    # known answers of the key: coordinate 0 in the least significant position, 20 bits per coordinate
    assert orc.morton_key(S, [0.0, 0.0]) == 0
    assert orc.morton_key(S, [2.0 ** -20, 0.0]) == 1 and orc.morton_key(S, [0.0, 2.0 ** -20]) == 2
    assert orc.morton_key(S, [3 * 2.0 ** -20, 0.0]) == 0b1001
    assert orc.morton_key(S, [1.0, 1.0]) == orc.morton_key(S, [1 - 2.0 ** -21, 1 - 2.0 ** -21])   # clamped to 2^20 - 1
    S3 = orc.StateSpace([0, 0, 0, -1], [2, 2, 2, 1])
    assert orc.morton_key(S3, [0.0, 0.0, 2.0 ** -19, 0.7]) == 4      # only the first three coordinates count
    # spatial coherence: consecutive samples are much closer than in draw order
    assert np.linalg.norm(np.diff(M, axis=0), axis=1).mean() < 0.1 * np.linalg.norm(np.diff(V, axis=0), axis=1).mean()


def test_candidate_stream_matches_the_pinned_golden_file(orc):
    """tests/golden/sample_stream.json (gen_sample_golden.py): the stream's definition must not drift"""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "sample_stream.json")) as f:
        G = json.load(f)
    spaces = {"unit2": ([0.0, 0.0], [1.0, 1.0]), "box3": ([-1.0, 2.0, 0.0], [1.0, 5.0, 0.5]),
              "di4": ([0.0, 0.0, -1.5, -1.5], [1.0, 1.0, 1.5, 1.5])}
    for key, hexes in G["candidates"].items():
        name, seed, c = key.split("/")
        S = orc.StateSpace(*spaces[name])
        x = orc.sample_candidate(S, int(seed), int(c))
        assert [float(v).hex() for v in x] == hexes, key
    for key, k in G["morton"].items():
        name, c = key.split("/")
        S = orc.StateSpace(*spaces[name])
        assert orc.morton_key(S, orc.sample_candidate(S, 1, int(c))) == k, key
