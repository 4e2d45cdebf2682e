"""Test infrastructure: FMT* (fmt.jl:4-119) restated over the ORACLE's lazy predicates -- per-vertex
r-ball queries (cached on first touch, nearneighbors.jl:129-135), per-candidate lazy edge checks
(fmt.jl:75) -- written independently of the product's planner."""
import heapq

import numpy as np


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
    heap = []
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
                heapq.heappush(heap, (c_min, x))
                H_new.append(x)
                W[x - 1] = False
        for x in H_new:
            H[x - 1] = True
        H[z - 1] = False
        if heap:
            _, z = heapq.heappop(heap)
        else:
            break
    sol = [z]
    while sol[0] != 1:
        sol.insert(0, int(A[sol[0] - 1]))
        if sol[0] == 0:
            break
    return dict(solved=bool(is_goal[z - 1]), cost=float(C[z - 1]), path=sol, tree=A, checks=checks)
