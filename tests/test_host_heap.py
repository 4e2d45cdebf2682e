"""The planner's queue restates Julia 0.5's Collections.PriorityQueue (fmt.jl:51,78,86): the order in which EQUAL costs
leave it is a property of that heap, not of (cost, index) ordering.  Product class vs the test oracle's independently
written heap on tie-heavy streams, the re-prioritisation path, and a hand-traced case."""
import random

from oracle_fmt import JuliaHeap


def test_priority_queue_matches_the_independent_heap_on_tie_heavy_streams(mp):
    from mpb200.planners import PriorityQueue
    rng = random.Random(7)
    differs_from_sorted = 0
    for _ in range(300):
        A, B = PriorityQueue(), JuliaHeap()
        key, outA, outB, live = 0, [], [], {}
        for _ in range(300):
            if rng.random() < 0.6 or len(A) == 0:
                key += 1
                pr = rng.choice([0.5, 1.0, 1.0, 1.0, 2.0, rng.random()])
                A[key] = pr
                B.push(key, pr)
                live[key] = pr
            else:
                a, b = A.dequeue(), B.pop()
                outA.append(a); outB.append(b)
                # always a minimum-priority element, but not always the smallest key among the ties
                assert live[a] == min(live.values())
                if a != min(k for k, v in live.items() if v == live[a]):
                    differs_from_sorted += 1
                del live[a]
        assert outA == outB
    assert differs_from_sorted > 0          # (cost, index) tuples in heapq would have produced a different order


def test_hand_traced_ties_and_reprioritisation(mp):
    from mpb200.planners import PriorityQueue
    q = PriorityQueue()
    for k in (1, 2, 3, 4, 5):
        q[k] = 1.0
    # array [1 2 3 4 5]; dequeue 1 -> 5 moves to the root and stays (children are not strictly smaller) -> 5 leaves next
    assert [q.dequeue() for _ in range(2)] == [1, 5]
    q[7] = 0.5
    q[4] = 0.25                              # existing key, smaller priority: percolates up past 7
    assert q.dequeue() == 4 and q.dequeue() == 7
    q[2] = 3.0                               # existing key, larger priority: percolates down
    assert [q.dequeue() for _ in range(len(q))][-1] == 2
