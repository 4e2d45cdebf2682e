#!/usr/bin/env python
"""Golden vectors for the GENERAL linear-affine steering cost (nilpotent A, drift c), produced by re-running the
REFERENCE's own SymPy construction (src/statespaces/linearquadratic.jl:94-157: expAt, G, Ginv, cdrift, xbar, cost,
dcost, ddcost, x(s)) with Python sympy for two systems the closed double-integrator form does not cover:
  "di2_drift": DoubleIntegrator(2) with constant drift c = (0, 0, 0.3, -0.5) and R = diag(1, 2)
  "triple":    a triple integrator (n = 3, m = 1), c = (0, 0.1, 0.2), R = [2]
The symbolic expressions are evaluated in 50-digit arithmetic at seeded random points; topt_newton
(linearquadratic.jl:175-190) runs on float64 lambdified closures.

Run in the build container (needs sympy; /root/reference is not read):  python tests/golden/gen_lq_general_golden.py
Output: tests/golden/lq_general.json
"""
import json
import os

import numpy as np
import sympy as sp

from gen_lq_golden import topt_newton


def build(A, B, c, R):
    A, B, c, R = sp.Matrix(A), sp.Matrix(B), sp.Matrix(c), sp.Matrix(R)
    n = A.shape[0]
    t, s = sp.symbols("t s", real=True)
    xS = sp.Matrix(sp.symbols("x1:%d" % (n + 1), real=True))
    yS = sp.Matrix(sp.symbols("y1:%d" % (n + 1), real=True))

    def expAt(tt):                                                    # linearquadratic.jl:94-98
        assert (A ** n).is_zero_matrix
        return sum((A ** i * (tt ** i / sp.factorial(i)) for i in range(n)), sp.zeros(n, n))

    expAtS, expAsS = expAt(t), expAt(s)
    GS = sp.integrate(expAtS * B * R.inv() * B.T * expAtS.T, t)       # :138
    GinvS = GS.inv()
    cdriftS = sp.integrate(expAtS, t) * c                             # :140
    xbarS = expAtS * xS + cdriftS
    costS = t + ((yS - xbarS).T * GinvS * (yS - xbarS))[0]            # :142
    dcostS = sp.diff(costS, t)
    ddcostS = sp.diff(costS, t, 2)
    xofsS = expAsS * xS + sp.integrate(expAsS, s) * c + \
        sp.integrate(expAsS * B * R.inv() * B.T * expAsS.T, s) * expAt(t - s).T * GinvS * (yS - xbarS)   # :145-146
    simp = lambda e: sp.simplify(sp.expand(e))                        # Sym2Function, :101
    return dict(t=t, s=s, x=xS, y=yS, n=n, cost=simp(costS), dcost=simp(dcostS), ddcost=simp(ddcostS),
                xofs=[simp(e) for e in xofsS])


SYSTEMS = {
    "di2_drift": dict(
        A=[[0, 0, 1, 0], [0, 0, 0, 1], [0, 0, 0, 0], [0, 0, 0, 0]], B=[[0, 0], [0, 0], [1, 0], [0, 1]],
        c=[0, 0, sp.Rational(3, 10), sp.Rational(-1, 2)], R=[[1, 0], [0, 2]],
        lo=[0, 0, -1.5, -1.5], hi=[1, 1, 1.5, 1.5]),
    "triple": dict(
        A=[[0, 1, 0], [0, 0, 1], [0, 0, 0]], B=[[0], [0], [1]], c=[0, sp.Rational(1, 10), sp.Rational(1, 5)], R=[[2]],
        lo=[0, -1, -2], hi=[1, 1, 2]),
}


def main():
    out = {}
    for name, sysd in SYSTEMS.items():
        S = build(sysd["A"], sysd["B"], sysd["c"], sysd["R"])
        n = S["n"]
        args = list(S["x"]) + list(S["y"]) + [S["t"]]
        f64 = {k: sp.lambdify(args, S[k], "math") for k in ("cost", "dcost", "ddcost")}
        call = lambda f: (lambda x0, x1, t: f(*x0, *x1, t))
        rng = np.random.Generator(np.random.PCG64(20241017 + n))
        lo, hi = np.array(sysd["lo"], dtype=float), np.array(sysd["hi"], dtype=float)
        Af = np.array(sysd["A"], dtype=float)
        cf = np.array([float(v) for v in sysd["c"]])
        cases = []
        for k in range(60):
            x0 = lo + rng.random(n) * (hi - lo)
            if k % 4 == 0:
                x1 = lo + rng.random(n) * (hi - lo)                   # far pair: optimum beyond r
            else:                                                     # near pair: roughly where the drift-free flow goes
                tau = 0.1 + 0.5 * rng.random()
                x1 = x0 + tau * (Af @ x0 + cf) + (rng.random(n) - 0.5) * 0.1 * (hi - lo)
            r = float(rng.choice([0.5, 0.8, 1.2]))
            tt = float(0.08 + rng.random() * 1.2)
            ss = float(rng.random() * tt)
            sub = {**{S["x"][i]: sp.Float(float(x0[i]), 60) for i in range(n)},
                   **{S["y"][i]: sp.Float(float(x1[i]), 60) for i in range(n)},
                   S["t"]: sp.Float(tt, 60), S["s"]: sp.Float(ss, 60)}
            hp = lambda e: float(sp.N(e.subs(sub), 50))
            topt = topt_newton(call(f64["dcost"]), call(f64["ddcost"]), x0, x1, r)
            sub_t = dict(sub); sub_t[S["t"]] = sp.Float(topt, 60)
            cases.append(dict(x0=x0.tolist(), x1=x1.tolist(), r=r, t=tt, s=ss,
                              cost=hp(S["cost"]), dcost=hp(S["dcost"]), ddcost=hp(S["ddcost"]),
                              xofs=[hp(e) for e in S["xofs"]],
                              topt=topt, cost_at_topt=float(sp.N(S["cost"].subs(sub_t), 50))))
        out[name] = dict(A=[[float(v) for v in row] for row in sysd["A"]], B=[[float(v) for v in row] for row in sysd["B"]],
                         c=[float(v) for v in sysd["c"]], R=[[float(v) for v in row] for row in sysd["R"]],
                         printed_cost=str(S["cost"])[:400], cases=cases)
        print(name, len(cases), "cases; cost =", str(S["cost"])[:100], "...")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lq_general.json")
    json.dump(dict(meta=dict(generator="tests/golden/gen_lq_general_golden.py", sympy=sp.__version__), systems=out),
              open(path, "w"), indent=0)
    print("wrote", path)


if __name__ == "__main__":
    main()
