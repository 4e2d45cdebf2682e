#!/usr/bin/env python
"""Golden vectors for the linear-quadratic steering cost, produced by re-running the REFERENCE's
own SymPy construction (src/statespaces/linearquadratic.jl:94-157: expAt for nilpotent A, G, Ginv,
xbar, cost, dcost, ddcost, x(s)) with Python sympy -- the same library the reference calls through
SymPy.jl/PyCall -- for DoubleIntegrator(2) (linearquadratic.jl:46-53), then evaluating the
symbolic expressions in 50-digit arithmetic at seeded random points, and running topt_newton
(linearquadratic.jl:175-190) on float64 lambdified closures.

Run in the build container (needs sympy; /root/reference is not read -- the construction is
restated from the cited lines):   python tests/golden/gen_lq_golden.py
Output: tests/golden/lq_di2.json
"""
import json
import os

import numpy as np
import sympy as sp


def build(d=2, rho=1):
    n = 2 * d
    A = sp.Matrix(n, n, lambda i, j: 1 if j == i + d else 0)          # [0 I; 0 0]
    B = sp.Matrix(n, d, lambda i, j: 1 if i == j + d else 0)          # [0; I]
    c = sp.zeros(n, 1)
    R = rho * sp.eye(d)
    t, s = sp.symbols("t s", real=True)
    xS = sp.Matrix(sp.symbols("x1:%d" % (n + 1), real=True))
    yS = sp.Matrix(sp.symbols("y1:%d" % (n + 1), real=True))

    def expAt(tt):                                                    # linearquadratic.jl:94-98
        assert (A ** n).is_zero_matrix
        return sum((A ** i * (tt ** i / sp.factorial(i)) for i in range(n)), sp.zeros(n, n))

    expAtS, expAsS = expAt(t), expAt(s)
    GS = sp.integrate(expAtS * B * R.inv() * B.T * expAtS.T, t)       # :138
    GinvS = GS.inv()
    cdriftS = sp.integrate(expAtS, t) * c
    xbarS = expAtS * xS + cdriftS
    costS = t + ((yS - xbarS).T * GinvS * (yS - xbarS))[0]            # :142
    dcostS = sp.diff(costS, t)
    ddcostS = sp.diff(costS, t, 2)
    xofsS = expAsS * xS + sp.integrate(expAsS, s) * c + \
        sp.integrate(expAsS * B * R.inv() * B.T * expAsS.T, s) * expAt(t - s).T * GinvS * (yS - xbarS)   # :145-146
    simp = lambda e: sp.simplify(sp.expand(e))                        # Sym2Function, :101
    return dict(t=t, s=s, x=xS, y=yS, cost=simp(costS), dcost=simp(dcostS), ddcost=simp(ddcostS),
                xofs=[simp(e) for e in xofsS])


def topt_newton(dc, ddc, x0, x1, tm, tol=1e-6):                       # linearquadratic.jl:175-190
    b = tm
    if dc(x0, x1, b) < 0:
        return tm
    a = tm / 100
    while dc(x0, x1, a) > 0:
        a /= 2
    t = tm / 2
    cdval = dc(x0, x1, t)
    while abs(cdval) > tol and abs(a - b) > tol:
        t = t - cdval / ddc(x0, x1, t)
        if t < a or t > b:
            t = (a + b) / 2
        cdval = dc(x0, x1, t)
        if cdval > 0:
            b = t
        else:
            a = t
    return t


def main():
    S = build()
    args = list(S["x"]) + list(S["y"]) + [S["t"]]
    f64 = {k: sp.lambdify(args, S[k], "math") for k in ("cost", "dcost", "ddcost")}
    call = lambda f: (lambda x0, x1, t: f(*x0, *x1, t))
    rng = np.random.Generator(np.random.PCG64(20240604))
    lo = np.array([0, 0, -1.5, -1.5]); hi = np.array([1, 1, 1.5, 1.5])
    cases = []
    for k in range(160):
        x0 = lo + rng.random(4) * (hi - lo)
        if k % 4 == 0:
            x1 = lo + rng.random(4) * (hi - lo)                       # far pair: optimum beyond r
        else:                                                         # near pair moving along its velocity
            tau = 0.1 + 0.5 * rng.random()
            x1 = x0 + np.concatenate([x0[2:] * tau, np.zeros(2)]) + (rng.random(4) - 0.5) * np.array([0.06, 0.06, 0.3, 0.3])
        r = float(rng.choice([0.5, 0.69, 1.0]))
        tt = float(0.05 + rng.random() * 1.2)
        ss = float(rng.random() * tt)
        sub = {**{S["x"][i]: sp.Float(float(x0[i]), 60) for i in range(4)},
               **{S["y"][i]: sp.Float(float(x1[i]), 60) for i in range(4)},
               S["t"]: sp.Float(tt, 60), S["s"]: sp.Float(ss, 60)}
        hp = lambda e: float(sp.N(e.subs(sub), 50))
        topt = topt_newton(call(f64["dcost"]), call(f64["ddcost"]), x0, x1, r)
        sub_t = dict(sub); sub_t[S["t"]] = sp.Float(topt, 60)
        cases.append(dict(x0=x0.tolist(), x1=x1.tolist(), r=r, t=tt, s=ss,
                          cost=hp(S["cost"]), dcost=hp(S["dcost"]), ddcost=hp(S["ddcost"]),
                          xofs=[hp(e) for e in S["xofs"]],
                          topt=topt, cost_at_topt=float(sp.N(S["cost"].subs(sub_t), 50))))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lq_di2.json")
    meta = dict(generator="tests/golden/gen_lq_golden.py", sympy=sp.__version__, system="DoubleIntegrator(2), R = I",
                printed_cost=str(S["cost"]), printed_dcost=str(S["dcost"]))
    json.dump(dict(meta=meta, cases=cases), open(out, "w"), indent=0)
    print("wrote", out, len(cases), "cases;", "cost =", meta["printed_cost"][:120], "...")


if __name__ == "__main__":
    main()
