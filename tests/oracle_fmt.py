"""Test infrastructure: FMT* (fmt.jl:4-119) restated over the ORACLE's lazy predicates -- per-vertex
r-ball queries (cached on first touch, nearneighbors.jl:129-135), per-candidate lazy edge checks
(fmt.jl:75) -- written independently of the product's planner."""
import numpy as np


class JuliaHeap:
    """Julia 0.5 Collections.PriorityQueue semantics (binary heap, strict comparisons when percolating), written for the
    tests independently of the product's class: equal priorities leave in the order THIS heap shape gives them."""

    def __init__(self):
        self.a = [None]          # 1-based: a[i] = [priority, key]

    def empty(self):
        return len(self.a) == 1

    def push(self, key, pr):
        a = self.a
        a.append([pr, key])
        i = len(a) - 1
        item = a[i]
        while i > 1 and item[0] < a[i // 2][0]:
            a[i] = a[i // 2]
            i //= 2
        a[i] = item

    def pop(self):
        a = self.a
        top = a[1]
        last = a.pop()
        n = len(a) - 1
        if n >= 1:
            i = 1
            while True:
                l, r = 2 * i, 2 * i + 1
                if l > n:
                    break
                j = r if (r <= n and not (a[l][0] < a[r][0])) else l
                if a[j][0] < last[0]:
                    a[i] = a[j]
                    i = j
                else:
                    break
            a[i] = last
        return top[1]


def fmt_oracle(V, is_goal, neighborsF, neighborsB, point_free, edge_free, checkpts=True):
    """V: N x d samples (sample 1 = init).  neighborsF/B(v) -> (inds 1-based ascending, dists);
    point_free(i0) -> bool; edge_free(y0, x0) -> (bool, segment checks run).  Returns dict."""
    N = len(V)
    F = np.array([point_free(i) for i in range(N)]) if checkpts else np.ones(N, dtype=bool)
    A = np.zeros(N, dtype=np.int64)
    W = np.ones(N, dtype=bool)
    H = np.zeros(N, dtype=bool)
    C = np.zeros(N)
    W[0] = False
    H[0] = True
    heap = JuliaHeap()
    z = 1
    checks = 0
    while not is_goal[z - 1]:
        H_new = []
        inds, _ = neighborsF(z)
        for x in [int(i) for i in inds if W[i - 1]]:
            if checkpts and not F[x - 1]:
                continue
            bi, bd = neighborsB(x)
            keep = [k for k in range(len(bi)) if H[bi[k] - 1]]
            costs = [C[bi[k] - 1] + bd[k] for k in keep]
            if not costs:      # only possible with k-nearest connections (x in knnF(z) but no open node in knnB(x)):
                continue       # findmin of an empty set would throw in the reference; specified here as 'skip x'
            j = min(range(len(costs)), key=lambda t: (costs[t], t))      # findmin: first minimum
            c_min, y_min = costs[j], int(bi[keep[j]])
            ok, n = edge_free(y_min - 1, x - 1)
            checks += n
            if ok:
                A[x - 1] = y_min
                C[x - 1] = c_min
                heap.push(x, c_min)
                H_new.append(x)
                W[x - 1] = False
        for x in H_new:
            H[x - 1] = True
        H[z - 1] = False
        if not heap.empty():
            z = heap.pop()
        else:
            break
    sol = [z]
    while sol[0] != 1:
        sol.insert(0, int(A[sol[0] - 1]))
        if sol[0] == 0:
            break
    return dict(solved=bool(is_goal[z - 1]), cost=float(C[z - 1]), path=sol, tree=A, checks=checks)
