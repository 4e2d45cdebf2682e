#!/usr/bin/env python
"""Golden vectors for the car metrics, produced by re-running the REFERENCE's own formulas
(src/statespaces/simplecars.jl: dubins :102-213, reedsshepp :230-523 -- the sweep over the path families in the
reference's order, `c <= cnew && return`, the timeflip / reflect / backwards post-processing) in 40-digit mpmath
arithmetic at seeded random state pairs.  Written from the cited lines independently of oracle/cars.c (a second
restatement, in a different language and number system): it pins the oracle's float64 arithmetic -- elementary
routines included -- to ~1e-12, and WHICH family wins, hence the control.

Besides cost and control every vector records `gap`: the distance between the best and the second-best DISTINCT
candidate length; the test only compares the control when the gap is comfortably above float64 noise.

Run in the build container (needs mpmath; /root/reference is not read):   python tests/golden/gen_cars_golden.py
Output: tests/golden/cars.json
"""
import json
import os

import numpy as np
from mpmath import mp, mpf

mp.dps = 40
PI = mp.pi
TWO_PI = 2 * mp.pi


def mod2pi(x):                      # utils.jl:91  mod(x, 2pi): floored
    r = x - TWO_PI * mp.floor(x / TWO_PI)
    return r


def seg(turn, d):                   # carsegment2stepcontrol(t, d) = (abs(d), (sign(d), t))
    return (abs(d), mp.sign(d), mpf(turn))


def dubins(s1, s2, r):
    vx, vy = (s2[0] - s1[0]) / r, (s2[1] - s1[1]) / r
    d = mp.sqrt(vx * vx + vy * vy)
    th = mp.atan2(vy, vx)
    a, b = mod2pi(s1[2] - th), mod2pi(s2[2] - th)
    ca, sa, cb, sb = mp.cos(a), mp.sin(a), mp.cos(b), mp.sin(b)
    cands = []

    def lsl():
        tmp = 2 + d * d - 2 * (ca * cb + sa * sb - d * (sa - sb))
        if tmp < 0:
            return None
        t_ = mp.atan2(cb - ca, d + sa - sb)
        t, p, q = mod2pi(-a + t_), mp.sqrt(max(tmp, 0)), mod2pi(b - t_)
        return t + p + q, [seg(1, t), seg(0, p), seg(1, q)]

    def rsr():
        tmp = 2 + d * d - 2 * (ca * cb + sa * sb - d * (sb - sa))
        if tmp < 0:
            return None
        t_ = mp.atan2(ca - cb, d - sa + sb)
        t, p, q = mod2pi(a - t_), mp.sqrt(max(tmp, 0)), mod2pi(-b + t_)
        return t + p + q, [seg(-1, t), seg(0, p), seg(-1, q)]

    def rsl():
        tmp = d * d - 2 + 2 * (ca * cb + sa * sb - d * (sa + sb))
        if tmp < 0:
            return None
        p = mp.sqrt(max(tmp, 0))
        t_ = mp.atan2(ca + cb, d - sa - sb) - mp.atan2(mpf(2), p)
        t, q = mod2pi(a - t_), mod2pi(b - t_)
        return t + p + q, [seg(-1, t), seg(0, p), seg(1, q)]

    def lsr():
        tmp = -2 + d * d + 2 * (ca * cb + sa * sb + d * (sa + sb))
        if tmp < 0:
            return None
        p = mp.sqrt(max(tmp, 0))
        t_ = mp.atan2(-ca - cb, d + sa + sb) - mp.atan2(mpf(-2), p)
        t, q = mod2pi(-a + t_), mod2pi(-b + t_)
        return t + p + q, [seg(1, t), seg(0, p), seg(-1, q)]

    def rlr():
        tmp = (6 - d * d + 2 * (ca * cb + sa * sb + d * (sa - sb))) / 8
        if abs(tmp) >= 1:
            return None
        p = TWO_PI - mp.acos(tmp)
        t_ = mp.atan2(ca - cb, d - sa + sb)
        t = mod2pi(a - t_ + p / 2)
        q = mod2pi(a - b - t + p)
        return t + p + q, [seg(-1, t), seg(1, p), seg(-1, q)]

    def lrl():
        tmp = (6 - d * d + 2 * (ca * cb + sa * sb - d * (sa - sb))) / 8
        if abs(tmp) >= 1:
            return None
        p = TWO_PI - mp.acos(tmp)
        t_ = mp.atan2(-ca + cb, d + sa - sb)
        t = mod2pi(-a + t_ + p / 2)
        q = mod2pi(b - a - t + p)
        return t + p + q, [seg(1, t), seg(-1, p), seg(1, q)]

    best, path = mp.inf, None
    for f in (lsl, rsr, rsl, lsr, rlr, lrl):
        out = f()
        if out is None:
            continue
        cands.append(out[0])
        if not best <= out[0]:
            best, path = out
    return best * r, [(t * r, u1, u2 / r) for t, u1, u2 in path], cands


def R(x, y):
    return mp.sqrt(x * x + y * y), mp.atan2(y, x)


def M(t):
    m = mod2pi(t)
    return m - TWO_PI if m > PI else m


def Tau(u, v, E, N):
    delta = M(u - v)
    A = mp.sin(u) - mp.sin(delta)
    B = mp.cos(u) - mp.cos(delta) - 1
    _, th = R(E * A + N * B, N * A - E * B)
    t = 2 * mp.cos(delta) - 2 * mp.cos(v) - 2 * mp.cos(u) + 3
    return M(th + PI) if t < 0 else M(th)


def Omega(u, v, E, N, t):
    return M(Tau(u, v, E, N) - u + v - t)


def LpSpLp(tx, ty, tt):
    r, th = R(tx - mp.sin(tt), ty - 1 + mp.cos(tt))
    u, t = r, mod2pi(th)
    v = mod2pi(tt - t)
    return t + u + v, [seg(1, t), seg(0, u), seg(1, v)]


def LpSpRp(tx, ty, tt):
    r, th = R(tx + mp.sin(tt), ty - 1 - mp.cos(tt))
    if r * r < 4:
        return None
    u = mp.sqrt(r * r - 4)
    _, th1 = R(u, mpf(2))
    t = mod2pi(th + th1)
    v = mod2pi(t - tt)
    return t + u + v, [seg(1, t), seg(0, u), seg(-1, v)]


def LpRmLp(tx, ty, tt):
    E, N = tx - mp.sin(tt), ty + mp.cos(tt) - 1
    if E * E + N * N > 16:
        return None
    r, th = R(E, N)
    u = mp.acos(1 - r * r / 8)
    t = mod2pi(th - u / 2 + PI)
    v = mod2pi(PI - u / 2 - th + tt)
    u = -u
    return t - u + v, [seg(1, t), seg(-1, u), seg(1, v)]


def LpRmLm(tx, ty, tt):
    E, N = tx - mp.sin(tt), ty + mp.cos(tt) - 1
    if E * E + N * N > 16:
        return None
    r, th = R(E, N)
    u = mp.acos(1 - r * r / 8)
    t = mod2pi(th - u / 2 + PI)
    v = mod2pi(PI - u / 2 - th + tt) - TWO_PI
    u = -u
    return t - u - v, [seg(1, t), seg(-1, u), seg(1, v)]


def LpRpuLmuRm(tx, ty, tt):
    E, N = tx + mp.sin(tt), ty - mp.cos(tt) - 1
    p = (2 + mp.sqrt(E * E + N * N)) / 4
    if p < 0 or p > 1:
        return None
    u = mp.acos(p)
    t = mod2pi(Tau(u, -u, E, N))
    v = mod2pi(Omega(u, -u, E, N, tt)) - TWO_PI
    return t + 2 * u - v, [seg(1, t), seg(-1, u), seg(1, -u), seg(-1, v)]


def LpRmuLmuRp(tx, ty, tt):
    E, N = tx + mp.sin(tt), ty - mp.cos(tt) - 1
    p = (20 - E * E - N * N) / 16
    if p < 0 or p > 1:
        return None
    u = -mp.acos(p)
    t = mod2pi(Tau(u, u, E, N))
    v = mod2pi(Omega(u, u, E, N, tt))
    return t - 2 * u + v, [seg(1, t), seg(-1, u), seg(1, u), seg(-1, v)]


def LpRmSmLm(tx, ty, tt):
    E, N = tx - mp.sin(tt), ty + mp.cos(tt) - 1
    D, beta = R(E, N)
    if D < 2:
        return None
    gamma = mp.acos(2 / D)
    F = mp.sqrt(D * D / 4 - 1)
    t = mod2pi(PI + beta - gamma)
    u = 2 - 2 * F
    if u > 0:
        return None
    v = mod2pi(-3 * PI / 2 + gamma + tt - beta) - TWO_PI
    return t + PI / 2 - u - v, [seg(1, t), seg(-1, -PI / 2), seg(0, u), seg(1, v)]


def LpRmSmRm(tx, ty, tt):
    E, N = tx + mp.sin(tt), ty - mp.cos(tt) - 1
    D, beta = R(E, N)
    if D < 2:
        return None
    t = mod2pi(beta + PI / 2)
    u = 2 - D
    if u > 0:
        return None
    v = mod2pi(-PI - tt + beta) - TWO_PI
    return t + PI / 2 - u - v, [seg(1, t), seg(-1, -PI / 2), seg(0, u), seg(-1, v)]


def LpRmSmLmRp(tx, ty, tt):
    E, N = tx + mp.sin(tt), ty - mp.cos(tt) - 1
    D, beta = R(E, N)
    if D < 2:
        return None
    gamma = mp.acos(2 / D)
    F = mp.sqrt(D * D / 4 - 1)
    t = mod2pi(PI + beta - gamma)
    u = 4 - 2 * F
    if u > 0:
        return None
    v = mod2pi(PI + beta - tt - gamma)
    return t + PI - u + v, [seg(1, t), seg(-1, -PI / 2), seg(0, u), seg(1, -PI / 2), seg(-1, v)]


def reedsshepp(s1, s2, r):
    dx, dy = (s2[0] - s1[0]) / r, (s2[1] - s1[1]) / r
    ct, st = mp.cos(s1[2]), mp.sin(s1[2])
    target = (dx * ct + dy * st, -dx * st + dy * ct, mod2pi(s2[2] - s1[2]))
    timeflip = lambda s: (-s[0], s[1], -s[2])
    reflect = lambda s: (s[0], -s[1], -s[2])
    backwards = lambda s: (s[0] * mp.cos(s[2]) + s[1] * mp.sin(s[2]), s[0] * mp.sin(s[2]) - s[1] * mp.cos(s[2]), s[2])
    T = target
    tT, rT = timeflip(T), reflect(T)
    trT = reflect(tT)
    bT = backwards(T)
    btT, brT = timeflip(bT), reflect(bT)
    btrT = reflect(btT)
    four = [(T, ""), (tT, "t"), (rT, "r"), (trT, "rt")]
    eight = four + [(bT, "b"), (btT, "bt"), (brT, "br"), (btrT, "brt")]
    plan = [(LpSpLp, four), (LpSpRp, four), (LpRmLp, [(T, ""), (rT, "r")]), (LpRmLm, eight), (LpRpuLmuRm, four),
            (LpRmuLmuRp, four), (LpRmSmLm, eight), (LpRmSmRm, eight), (LpRmSmLmRp, four)]
    best, path, post, cands = mp.inf, None, "", []
    for f, targets in plan:
        for tg, name in targets:
            out = f(*tg)
            if out is None:
                continue
            cands.append(out[0])
            if not best <= out[0]:
                best, path, post = out[0], out[1], name
    u = [(t * r, u1, u2 / r) for t, u1, u2 in path]          # scaleradius!, scalespeed!(s = 1)
    if "t" in post:
        u = [(t, -u1, u2) for t, u1, u2 in u]
    if "r" in post:
        u = [(t, u1, -u2) for t, u1, u2 in u]
    if "b" in post:
        u = u[::-1]
    return best * r, u, cands


def gap_of(cands):
    c = sorted(cands)
    best = c[0]
    others = [x for x in c if x - best > mpf(10) ** -25]
    return float(others[0] - best) if others else float("inf")


def main():
    rng = np.random.Generator(np.random.PCG64(20240606))
    out = {"dps": mp.dps, "turning_radius": 0.7, "cases": []}
    r = mpf(0.7)
    pairs = [(rng.uniform(0, 3, 2).tolist() + [rng.uniform(0, 2 * np.pi)],
              rng.uniform(0, 3, 2).tolist() + [rng.uniform(0, 2 * np.pi)]) for _ in range(150)]
    # close pairs (within a couple of turning radii: the C|C|C and C Cu|Cu C families win there)
    for _ in range(100):
        v = rng.uniform(0, 3, 2).tolist() + [rng.uniform(0, 2 * np.pi)]
        w = [v[0] + rng.normal(0, 0.5), v[1] + rng.normal(0, 0.5), rng.uniform(0, 2 * np.pi)]
        pairs.append((v, w))
    for v, w in pairs:
        s1, s2 = [mpf(x) for x in v], [mpf(x) for x in w]
        case = {"v": [float(x).hex() for x in v], "w": [float(x).hex() for x in w]}
        for name, fn in (("dubins", dubins), ("reedsshepp", reedsshepp)):
            c, u, cands = fn(s1, s2, r)
            case[name] = {"cost": mp.nstr(c, 25), "gap": gap_of(cands),
                          "control": [[mp.nstr(t, 25), int(u1), mp.nstr(u2, 25)] for t, u1, u2 in u]}
        out["cases"].append(case)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cars.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
