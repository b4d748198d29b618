#!/usr/bin/env python
"""Generates tests/golden/sympy_*.npz — known-answer vectors that are INDEPENDENT of this repo's code.

The reference's evaluator (ExaModels.jl) cannot run in this container, and none of the reference's
tests pins cons!/jac_coord!/hess_coord! values.  What can be pinned without it is
the MATHEMATICS of the transcribed models: this script restates, in sympy and from the reference's
model files only, the nonlinear programs that `ExaTranscriptionBackend` produces

    ode_5x5      test/madnlp.jl:4-11 (== test/ipopt.jl:4-11), nvar 51 / ncon 70 (test/ipopt.jl:183-186)
    quadrotor_fd examples/quadrotor.jl:6-77, backward finite difference (InfiniteOpt default), N = 5
    quadrotor_oc ESCAPE34/quadrotor.jl:4-76, OrthogonalCollocation(3) + piecewise-constant controls, N = 4
    pandemic     ESCAPE34/pandemic.jl:4-34, 4 public time supports (+10 extra), 2 scenarios
    farmer       examples/2stage_example.jl:6-37, 4 scenarios
    operators    every operator of src/operators.jl:8-43 (true mathematical definition) plus + - * / ^

with the x / θ / row layout of src/transform.jl (finite variables first, infinite variables as
column-major blocks, derivative variables last :134-158; rows in constraint order, then derivative
approximations :511-562, then collocation restrictions :565-601), differentiates them SYMBOLICALLY,
and evaluates everything in 40-digit arithmetic at a seeded point before rounding to fp64.  It imports
nothing from this repository: neither the engine, nor the oracle, nor the hand transcriptions in
`infiniteexamodels.jl_b200/models.py`.

Stored per model: x, y, sigma, obj, grad, cons, the Jacobian and the lower triangle of the Hessian of
the Lagrangian  sigma*f + y'c  as sorted (row, col, value) triplets of their structurally non-zero
entries.  tests/test_sympy_golden.py sums the duplicate COO entries of the oracle / the CUDA engine
into the same dense form and compares.

usage:  python tests/golden/make_sympy_golden.py        (about a minute)
"""
import os

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
DIGITS = 40


def R(v):
    """exact rational value of a double (model data are fp64 numbers on both sides)"""
    return sp.Rational(float(v))


def trapezoid(s):
    s = np.asarray(s, dtype=np.float64)
    c = np.zeros_like(s)
    d = np.diff(s)
    c[:-1] += d / 2
    c[1:] += d / 2
    return c


class NLP:
    def __init__(self, name):
        self.name = name
        self.vars = []     # sympy symbols, x order
        self.x0 = []
        self.cons = []     # sympy expressions, row order
        self.obj = sp.Integer(0)

    def var(self, label, n, start=0.0):
        syms = [sp.Symbol(f"{label}_{k}", real=True) for k in range(n)]
        self.vars += syms
        self.x0 += [start] * n
        return syms

    def dump(self, seed=0, scale=0.1, x=None):
        rng = np.random.default_rng(seed)
        n, m = len(self.vars), len(self.cons)
        if x is None:
            x = np.asarray(self.x0) + scale * rng.uniform(-1, 1, n)
        y = rng.uniform(-1, 1, m)
        sigma = 0.7
        sub = {s: R(v) for s, v in zip(self.vars, x)}
        pos = {s: i for i, s in enumerate(self.vars)}

        def val(e):
            return float(sp.N(e.xreplace(sub), DIGITS))

        obj = val(self.obj)
        cons = np.array([val(c) for c in self.cons])
        grad = np.zeros(n)
        for s in self.obj.free_symbols:
            grad[pos[s]] = val(sp.diff(self.obj, s))
        jr, jc, jv = [], [], []
        H = {}

        def add_h(expr, w):
            fs = sorted(expr.free_symbols, key=lambda s: pos[s])
            for a in fs:
                da = sp.diff(expr, a)
                if da.is_number:
                    continue
                for b in fs:
                    if pos[b] > pos[a]:
                        continue
                    d2 = sp.diff(da, b)
                    if d2 == 0:
                        continue
                    key = (pos[a], pos[b])
                    H[key] = H.get(key, sp.Integer(0)) + w * d2.xreplace(sub)

        for i, c in enumerate(self.cons):
            for s in sorted(c.free_symbols, key=lambda s: pos[s]):
                d = sp.diff(c, s)
                if d == 0:
                    continue
                jr.append(i); jc.append(pos[s]); jv.append(val(d))
            add_h(c, R(y[i]))
        add_h(self.obj, R(sigma))
        keys = sorted(H)
        hr = np.array([k[0] for k in keys], dtype=np.int64)
        hc = np.array([k[1] for k in keys], dtype=np.int64)
        hv = np.array([float(sp.N(H[k], DIGITS)) for k in keys])
        out = os.path.join(HERE, f"sympy_{self.name}.npz")
        np.savez_compressed(out, x=x, y=y, sigma=sigma, obj=obj, grad=grad, cons=cons,
                            jr=np.array(jr, dtype=np.int64), jc=np.array(jc, dtype=np.int64), jv=np.array(jv),
                            hr=hr, hc=hc, hv=hv, nvar=n, ncon=m)
        print(f"{self.name}: nvar={n} ncon={m} nnz(J)={len(jv)} nnz(tril H)={len(hv)} obj={obj!r} -> {out}")


# ------------------------------------------------------------------------------------------------
def ode_5x5():
    nt = nx = 5
    ts, xs = np.linspace(0, 1, nt), np.linspace(-1, 1, nx)
    P = NLP("ode_5x5")
    z = P.var("z", 1, 10.0)[0]
    y = np.array(P.var("y", nt * nx)).reshape(nx, nt).T        # y[i_t, i_x], t fastest in memory
    dy = np.array(P.var("dy", nt * nx)).reshape(nx, nt).T
    P.cons += [dy[i, j] - (sp.sin(y[i, j]) + z + R(1.2)) for j in range(nx) for i in range(nt)]
    P.cons += [y[i, j] + z - R(ts[i]) for j in range(nx) for i in range(nt)]
    dt = np.diff(ts)
    P.cons += [R(dt[i - 1]) * dy[i, j] - y[i, j] + y[i - 1, j] for j in range(nx) for i in range(1, nt)]
    wt, wx = trapezoid(ts), trapezoid(xs)
    P.obj = sum(R(wx[j] * wt[i]) * (y[i, j] ** 2 + 2 * z) for j in range(nx) for i in range(nt))
    P.dump()


def quadrotor(method, N):
    Tend = 60.0
    pub = np.linspace(0.0, Tend, N)
    if method == "oc":
        T = 2 * N - 1
        ts = np.empty(T); ts[0::2] = pub; ts[1::2] = 0.5 * (pub[:-1] + pub[1:])
    else:
        T, ts = N, pub
    P = NLP(f"quadrotor_{method}")
    x = [None] + [P.var(f"x{j}", T) for j in range(1, 10)]
    u = [None] + [P.var(f"u{j}", T) for j in range(1, 5)]
    d = [None] + [P.var(f"dx{j}", T) for j in range(1, 10)]
    d1 = np.sin(2 * np.pi * ts / Tend); d3 = 2 * np.sin(4 * np.pi * ts / Tend); d5 = 2 * (ts / Tend)
    P.cons += [x[j][0] for j in range(1, 10)]
    s, c, tan = sp.sin, sp.cos, sp.tan
    rhs = [None,
           lambda k: x[2][k],
           lambda k: u[1][k] * c(x[7][k]) * s(x[8][k]) * c(x[9][k]) + u[1][k] * s(x[7][k]) * s(x[9][k]),
           lambda k: x[4][k],
           lambda k: u[1][k] * c(x[7][k]) * s(x[8][k]) * s(x[9][k]) - u[1][k] * s(x[7][k]) * c(x[9][k]),
           lambda k: x[6][k],
           lambda k: u[1][k] * c(x[7][k]) * c(x[8][k]) - R(9.8),
           lambda k: u[2][k] * c(x[7][k]) / c(x[8][k]) + u[3][k] * s(x[7][k]) / c(x[8][k]),
           lambda k: -u[2][k] * s(x[7][k]) + u[3][k] * c(x[7][k]),
           lambda k: u[2][k] * c(x[7][k]) * tan(x[8][k]) + u[3][k] * s(x[7][k]) * tan(x[8][k]) + u[4][k]]
    for j in range(1, 10):
        P.cons += [d[j][k] - rhs[j](k) for k in range(T)]
    if method == "fd":
        dt = np.diff(ts)
        for j in range(1, 10):
            P.cons += [R(dt[k - 1]) * d[j][k] - x[j][k] + x[j][k - 1] for k in range(1, T)]
    else:
        # 3-node Lobatto collocation on [lb, ub] with the midpoint as internal node, written as the integral
        # form  x(node) - x(lb) = sum_m M[node, m] * dx(m)  over the two non-initial nodes:
        #   midpoint: h*(3/4*dx_mid - 1/4*dx_ub),   upper bound: h*dx_mid      (h = ub - lb)
        h = np.diff(pub)
        for j in range(1, 10):
            for e in range(N - 1):
                lb, mid, ub = 2 * e, 2 * e + 1, 2 * e + 2
                P.cons.append(R(0.75 * h[e]) * d[j][mid] + R(-0.25 * h[e]) * d[j][ub] - x[j][mid] + x[j][lb])
                P.cons.append(R(h[e]) * d[j][mid] + R(0.0) * d[j][ub] - x[j][ub] + x[j][lb])
        for j in range(1, 5):   # piecewise-constant controls: u(ub) - u(internal node) = 0 (transform.jl:565-601)
            P.cons += [u[j][2 * e + 2] - u[j][2 * e + 1] for e in range(N - 1)]
    w = trapezoid(ts)
    P.obj = sum(R(w[k]) * ((x[1][k] - R(d1[k])) ** 2 + (x[3][k] - R(d3[k])) ** 2 + (x[5][k] - R(d5[k])) ** 2
                           + x[7][k] ** 2 + x[8][k] ** 2 + x[9][k] ** 2
                           + R(0.1) * (u[1][k] ** 2 + u[2][k] ** 2 + u[3][k] ** 2 + u[4][k] ** 2)) for k in range(T))
    P.dump()


def pandemic(num_supports=4, S=2, seed=0):
    gamma, beta, Npop = 0.303, 0.727, 1e5
    extra = np.array([0.001, 0.002, 0.004, 0.008, 0.02, 0.04, 0.08, 0.2, 0.4, 0.8])
    ts = np.unique(np.concatenate([np.linspace(0, 200, num_supports), extra]))
    T = len(ts)
    xi = np.random.default_rng(seed).uniform(0.1, 0.6, S)
    P = NLP("pandemic")
    blk = lambda name: np.array(P.var(name, T * S)).reshape(S, T).T   # [i_t, i_xi], t fastest
    s, e, i_, r = blk("s"), blk("e"), blk("i"), blk("r")
    u = P.var("u", T, 0.2)
    ds, de, di, dr = blk("ds"), blk("de"), blk("di"), blk("dr")
    for v in (s, e, i_, r):
        P.cons += [v[0, j] for j in range(S)]
    both = [(k, j) for j in range(S) for k in range(T)]
    P.cons += [ds[k, j] - (-(1 - u[k]) * R(beta) * s[k, j] * i_[k, j]) for k, j in both]
    P.cons += [de[k, j] - ((1 - u[k]) * R(beta) * s[k, j] * i_[k, j] - R(xi[j]) * e[k, j]) for k, j in both]
    P.cons += [di[k, j] - (R(xi[j]) * e[k, j] - R(gamma) * i_[k, j]) for k, j in both]
    P.cons += [dr[k, j] - R(gamma) * i_[k, j] for k, j in both]
    P.cons += [i_[k, j] for k, j in both]
    dt = np.diff(ts)
    for v, dv in ((s, ds), (e, de), (i_, di), (r, dr)):
        P.cons += [R(dt[k - 1]) * dv[k, j] - v[k, j] + v[k - 1, j] for j in range(S) for k in range(1, T)]
    w = trapezoid(ts)
    P.obj = sum(R(w[k]) * u[k] for k in range(T))
    # evaluate away from x0 = 0 so that the bilinear terms have non-trivial derivatives
    rng = np.random.default_rng(7)
    P.dump(x=rng.uniform(0.05, 0.9, len(P.vars)))


def farmer(K=4, seed=42):
    rng = np.random.default_rng(seed)
    xi = np.stack([rng.uniform(0, 5, K), rng.uniform(0, 5, K), rng.uniform(10, 30, K)])
    alpha, beta, lam, dem = [150, 230, 260], [238, 210, 0], [170, 150, 36], [200, 240, 0]
    P = NLP("farmer")
    x = [P.var(f"x{c}", 1)[0] for c in range(3)]
    y = [P.var(f"y{c}", K) for c in range(3)]
    w = [P.var(f"w{c}", K) for c in range(3)]
    P.cons.append(x[0] + x[1] + x[2])
    for c in range(3):
        P.cons += [R(xi[c][k]) * x[c] + y[c][k] - w[c][k] for k in range(K)]
    P.cons += [w[2][k] for k in range(K)]
    P.cons += [y[2][k] for k in range(K)]
    P.obj = sum(alpha[c] * x[c] for c in range(3)) + sum(
        R(1.0 / K) * (beta[0] * y[0][k] + beta[1] * y[1][k] - lam[0] * w[0][k] - lam[1] * w[1][k] - lam[2] * w[2][k])
        for k in range(K))
    rng2 = np.random.default_rng(3)
    P.dump(x=rng2.uniform(0.0, 100.0, len(P.vars)))


# operator table of src/operators.jl:8-43 by mathematical definition (degree variants: argument in degrees,
# inverse degree variants: result in degrees); (function, evaluation interval)
_deg = sp.pi / 180
OPS = {
    "inv": (lambda a: 1 / a, (0.5, 2)), "sqrt": (sp.sqrt, (0.5, 2)), "cbrt": (lambda a: a ** sp.Rational(1, 3), (0.5, 2)),
    "abs": (lambda a: sp.sqrt(a * a), (0.2, 1.5)), "abs2": (lambda a: a * a, (-1, 1)), "exp": (sp.exp, (-1, 1)),
    "exp2": (lambda a: 2 ** a, (-1, 1)), "log": (sp.log, (0.5, 2)), "log2": (lambda a: sp.log(a) / sp.log(2), (0.5, 2)),
    "log10": (lambda a: sp.log(a) / sp.log(10), (0.5, 2)), "log1p": (lambda a: sp.log(1 + a), (0.2, 2)),
    "sin": (sp.sin, (-1, 1)), "cos": (sp.cos, (-1, 1)), "tan": (sp.tan, (-1, 1)), "asin": (sp.asin, (-0.7, 0.7)),
    "acos": (sp.acos, (-0.7, 0.7)), "csc": (lambda a: 1 / sp.sin(a), (0.4, 1.2)), "sec": (lambda a: 1 / sp.cos(a), (-1, 1)),
    "cot": (lambda a: sp.cos(a) / sp.sin(a), (0.4, 1.2)), "atan": (sp.atan, (-1, 1)),
    "acot": (lambda a: sp.atan(1 / a), (0.3, 2)),
    "sind": (lambda a: sp.sin(a * _deg), (-80, 80)), "cosd": (lambda a: sp.cos(a * _deg), (-80, 80)),
    "tand": (lambda a: sp.tan(a * _deg), (-50, 50)), "cscd": (lambda a: 1 / sp.sin(a * _deg), (20, 70)),
    "secd": (lambda a: 1 / sp.cos(a * _deg), (-50, 50)), "cotd": (lambda a: sp.cos(a * _deg) / sp.sin(a * _deg), (20, 70)),
    "atand": (lambda a: sp.atan(a) / _deg, (-1, 1)), "acotd": (lambda a: sp.atan(1 / a) / _deg, (0.3, 2)),
    "sinh": (sp.sinh, (-1, 1)), "cosh": (sp.cosh, (-1, 1)), "tanh": (sp.tanh, (-1, 1)),
    "csch": (lambda a: 1 / sp.sinh(a), (0.4, 1.2)), "sech": (lambda a: 1 / sp.cosh(a), (-1, 1)),
    "coth": (lambda a: sp.cosh(a) / sp.sinh(a), (0.4, 1.5)), "atanh": (sp.atanh, (-0.7, 0.7)),
    "acoth": (lambda a: sp.log((a + 1) / (a - 1)) / 2, (1.3, 3)),
}
BINARY = {
    "add": (lambda a, b: (a + b) * (a + R(0.75)), (-1, 1)),
    "sub": (lambda a, b: (a - b) * (R(0.75) - a) * (b - 2), (-1, 1)),
    "mul": (lambda a, b: a * b * a * R(0.75), (-1, 1)),
    "div": (lambda a, b: a / b + R(0.75) / a + b / R(0.75), (0.5, 2)),
    "pow_var_const": (lambda a, b: a ** 3 + b ** 2 + a ** R(0.75), (0.5, 2)),
    "pow_const_var": (lambda a, b: 2 ** a + R(0.75) ** b, (-1, 1)),
    "pow_var_var": (lambda a, b: a ** b, (0.5, 2)),
}


def operators():
    """One row per operator:  op(a*1)*b + op(0.75*a)  for the unary ones (variables a = x[2r], b = x[2r+1]);
    the objective is the sum of all rows.  tests/test_sympy_golden.py builds the same rows through nl_op."""
    P = NLP("operators")
    rng = np.random.default_rng(11)
    xs = []
    names = list(OPS) + list(BINARY)
    for nm in names:
        a, b = P.var(f"a_{nm}", 1)[0], P.var(f"b_{nm}", 1)[0]
        if nm in OPS:
            f, (lo, hi) = OPS[nm]
            row = f(a) * b + f(R(0.75) * a)
        else:
            f, (lo, hi) = BINARY[nm]
            row = f(a, b)
        P.cons.append(row)
        va, vb = rng.uniform(lo, hi, 2)
        if nm == "abs":
            va = abs(va) + 0.1
        if nm in ("acoth",):        # 0.75*a must stay inside the domain too
            va = rng.uniform(2.0, 3.0)
        xs += [va, vb]
    P.obj = sum(P.cons)
    P.names = names
    P.dump(x=np.array(xs))
    with open(os.path.join(HERE, "sympy_operators.names"), "w") as f:
        f.write("\n".join(names) + "\n")


def solve_problem(variant=-1):
    """test/solve.jl:2-27 ("Test Problem 1", ``variant=-1``, finite differences) and :48-95 ("Test Problem 2", ``variant=0``,
    and its four alternative objectives 1..4) — the models the reference checks DIFFERENTIALLY against JuMP's
    TranscriptionBackend, i.e. against the fully expanded scalar NLP, which is what is written out here.  They exercise the
    point variable y(0, 1), a DomainRestriction, the derivative of a semi-infinite variable, and the objective heuristics of
    src/transform.jl:642-767 (terms moved inside the inner measure / full expansion).  Layout: z, y (t fastest),
    d(y)/dt, d(y(0, x))/dx.  Rows are stored in the order written here; the tests match rows by value.  The returned NLP
    also carries the bounds (lvar/uvar/lcon/ucon) so that tests/test_differential_solves.py can SOLVE it."""
    nt = nx = 5
    ts, xs = np.linspace(0, 1, nt), np.linspace(-1, 1, nx)
    wt, wx = trapezoid(ts), trapezoid(xs)
    dt, dx = np.diff(ts), np.diff(xs)
    restricted = variant == -1
    P = NLP("solve_tp1" if restricted else f"solve_tp2_v{variant}")
    z = P.var("z", 1, 10.0)[0]
    y = np.array(P.var("y", nt * nx)).reshape(nx, nt).T
    dy = np.array(P.var("dy", nt * nx)).reshape(nx, nt).T
    d2 = P.var("d2", nx)                                        # d/dx of y(0, x)
    inf = float("inf")
    P.lvar = [-inf] + [0.0] * (nt * nx) + [-inf] * (nt * nx + nx); P.uvar = [inf] * len(P.vars)
    P.lcon, P.ucon = [], []

    def rows(exprs, lo, hi):
        P.cons += exprs; P.lcon += [lo] * len(exprs); P.ucon += [hi] * len(exprs)

    rows([dy[i, j] - (sp.sin(y[i, j]) + z + R(1.2)) for j in range(nx) for i in range(nt)], 0.0, 0.0)
    rows([y[i, j] + z - R(ts[i]) for j in range(nx) for i in range(nt) if (not restricted or ts[i] <= 0.5)], -inf, 42.0)
    rows([d2[j] for j in range(nx)], 5.0, 5.0)                    # == 5 (the constant lives in the set)
    rows([R(dt[i - 1]) * dy[i, j] - y[i, j] + y[i - 1, j] for j in range(nx) for i in range(1, nt)], 0.0, 0.0)
    rows([R(dx[j - 1]) * d2[j] - y[0, j] + y[0, j - 1] for j in range(1, nx)], 0.0, 0.0)
    I = [sum(R(wt[i]) * y[i, j] ** 2 for i in range(nt)) for j in range(nx)]   # ∫(y², t) at x_j
    if variant == -1:
        P.obj = sum(R(wx[j]) * I[j] for j in range(nx)) + 2 * y[0, nx - 1]          # ∫(∫(y², t), x) + 2y(0, 1)
    elif variant == 0:
        P.obj = sum(R(wx[j]) * (I[j] + 2 * z) for j in range(nx)) + 2 * y[0, nx - 1]
    elif variant == 1:
        P.obj = sum(R(wx[j]) * (I[j] + 2 * z ** 2) for j in range(nx)) + 2 * y[0, nx - 1]
    elif variant == 2:
        P.obj = sum(R(wx[j]) * (I[j] + sp.sin(z ** 2)) for j in range(nx))
    elif variant == 3:
        P.obj = sum(R(wx[j]) * (I[j] * sp.cos(z)) for j in range(nx))
    else:
        P.obj = sum(R(wx[j]) * (z * (I[j] + z ** 3)) for j in range(nx))
    return P


def solve_tests():
    rng = np.random.default_rng(21)
    for variant in (-1, 0, 1, 2, 3, 4):
        P = solve_problem(variant)
        P.dump(x=rng.uniform(0.2, 1.5, len(P.vars)))


def param_function_problem():
    """test/solve.jl:97-131 ("Parameter Function Problem"): parameter functions pf(t), pf2(t, s) (piecewise), a semi-
    infinite variable z(t, 2.5) inside a constraint, and a measure of a parameter function inside a constraint
    (c5: v*∫(pf2, s) <= 100, expanded inline — transform.jl:430-435).  Layout: v(t), z(t, s) (t fastest)."""
    nt = ns = 5
    ts, ss = np.linspace(0, 1, nt), np.linspace(2, 3, ns)
    wt, ws = trapezoid(ts), trapezoid(ss)
    ti = 0.2
    pf = np.sin(ts)
    pf2 = np.array([[np.cos(t) * s - ti if t <= 0.5 else np.sin(t) * s + ti for s in ss] for t in ts])   # [i_t, i_s]
    P = NLP("param_function_problem")
    v = P.var("v", nt)
    z = np.array(P.var("z", nt * ns)).reshape(ns, nt).T
    s25 = int(np.argmin(np.abs(ss - 2.5)))
    P.cons += [v[i] + R(pf[i]) for i in range(nt)]                                              # c1
    P.cons += [2 * v[i] + R(pf[i]) * R(pf2[i, j]) for j in range(ns) for i in range(nt)]      # c2
    P.cons += [v[i] - R(0.2) * R(pf2[i, j]) for j in range(ns) for i in range(nt)]            # c3
    P.cons += [z[i, s25] + R(pf2[i, j]) * R(pf[i]) for j in range(ns) for i in range(nt)]     # c4
    P.cons += [v[i] * sum(R(ws[j]) * R(pf2[i, j]) for j in range(ns)) for i in range(nt)]      # c5
    P.obj = sum(R(wt[i]) * v[i] * R(pf[i]) for i in range(nt)) + sum(
        R(ws[j]) * R(wt[i]) * R(0.5) * z[i, j] * R(pf2[i, j]) for j in range(ns) for i in range(nt))
    rng = np.random.default_rng(33)
    P.dump(x=rng.uniform(0.5, 5.0, len(P.vars)))


def opf_case3(K=2, seed=0):
    """ESCAPE34/opf.jl:36-286 — two-stage stochastic AC-OPF (BASELINE configs[3]) on the 3-bus / 3-branch / 3-generator
    case this repo embeds instead of the pglib download of opf.jl:15-18 (the grid NUMBERS below are that embedded case;
    the per-unit conversion, branch admittances and every constraint are restated from opf.jl and PowerModels' conventions,
    not from infiniteexamodels.jl_b200/opf.py).  K scenarios of the 2*nbus load perturbation theta ~ N(0, diag((0.1*[Pd; Qd])^2))
    (opf.jl:48-50,112).  Layout: the 24 first-stage variables in declaration order (va0, vm0, pg0, qg0, p0, q0 over
    buses / generators / arcs), then the 24 second-stage variables as blocks of K.  Rows are matched by value."""
    base = 100.0
    bus = {1: (110.0, 40.0), 2: (110.0, 40.0), 3: (95.0, 50.0)}                       # pd, qd [MW, MVAr]
    gen = {1: (1, (0.11, 5.0, 0.0)), 2: (2, (0.085, 1.2, 0.0)), 3: (3, (0.0, 0.0, 0.0))}   # bus, cost (quadratic, linear, constant)
    pmax = {1: 2000.0, 2: 2000.0, 3: 0.0}; qlim = 1000.0
    branch = {1: (1, 3, 0.065, 0.62, 0.45, 9000.0), 2: (3, 2, 0.025, 0.75, 0.7, 50.0), 3: (1, 2, 0.042, 0.9, 0.3, 9000.0)}
    buses = [1, 2, 3]
    nbus = 3
    arcs = [(l, f, t) for l, (f, t, *_) in branch.items()] + [(l, t, f) for l, (f, t, *_) in branch.items()]
    rng = np.random.default_rng(seed)
    sd = 0.1 * np.array([bus[i][0] / base for i in buses] + [bus[i][1] / base for i in buses])
    theta = rng.normal(0.0, 1.0, size=(2 * nbus, K)) * sd[:, None]

    P = NLP("opf_case3")
    va0 = {i: P.var(f"va0_{i}", 1)[0] for i in buses}
    vm0 = {i: P.var(f"vm0_{i}", 1, 1.0)[0] for i in buses}
    pg0 = {i: P.var(f"pg0_{i}", 1)[0] for i in gen}
    qg0 = {i: P.var(f"qg0_{i}", 1)[0] for i in gen}
    p0 = {a: P.var("p0_%d_%d_%d" % a, 1)[0] for a in arcs}
    q0 = {a: P.var("q0_%d_%d_%d" % a, 1)[0] for a in arcs}
    va = {i: P.var(f"va_{i}", K) for i in buses}
    vm = {i: P.var(f"vm_{i}", K, 1.0) for i in buses}
    pg = {i: P.var(f"pg_{i}", K) for i in gen}
    qg = {i: P.var(f"qg_{i}", K) for i in gen}
    p = {a: P.var("p_%d_%d_%d" % a, K) for a in arcs}
    q = {a: P.var("q_%d_%d_%d" % a, K) for a in arcs}
    P.obj = sum(R(c[0] * base ** 2) * pg0[i] ** 2 + R(c[1] * base) * pg0[i] + R(c[2]) for i, (_, c) in gen.items())

    def stage(va, vm, pg, qg, p, q, th):
        """th: None (first stage) or (k -> vector of 2*nbus perturbations); variables are scalars or functions of k"""
        rows = []
        rows.append(va[1])                                                    # reference bus: va == 0
        for l, (f, t, r, x, bch, rate) in branch.items():
            z2 = r * r + x * x
            g, b = r / z2, -x / z2                                            # series admittance y = 1/(r + jx)
            tr, ti, ttm = 1.0, 0.0, 1.0                                        # no transformer: tap 1, shift 0
            g_fr = g_to = 0.0; b_fr = b_to = bch / 2                          # line charging split over both ends
            cfr, sfr = sp.cos(va[f] - va[t]), sp.sin(va[f] - va[t])
            cto, sto = sp.cos(va[t] - va[f]), sp.sin(va[t] - va[f])
            rows.append(("pf", p[(l, f, t)] - (R((g + g_fr) / ttm) * vm[f] ** 2 + R((-g * tr + b * ti) / ttm) * (vm[f] * vm[t] * cfr)
                                                + R((-b * tr - g * ti) / ttm) * (vm[f] * vm[t] * sfr))))
            rows.append(("qf", q[(l, f, t)] - (-R((b + b_fr) / ttm) * vm[f] ** 2 - R((-b * tr - g * ti) / ttm) * (vm[f] * vm[t] * cfr)
                                                + R((-g * tr + b * ti) / ttm) * (vm[f] * vm[t] * sfr))))
            rows.append(("pt", p[(l, t, f)] - (R(g + g_to) * vm[t] ** 2 + R((-g * tr - b * ti) / ttm) * (vm[t] * vm[f] * cto)
                                                + R((-b * tr + g * ti) / ttm) * (vm[t] * vm[f] * sto))))
            rows.append(("qt", q[(l, t, f)] - (-R(b + b_to) * vm[t] ** 2 - R((-b * tr + g * ti) / ttm) * (vm[t] * vm[f] * cto)
                                                + R((-g * tr - b * ti) / ttm) * (vm[t] * vm[f] * sto))))
            rows.append(("ang", va[f] - va[t]))
            rows.append(("sf", p[(l, f, t)] ** 2 + q[(l, f, t)] ** 2))
            rows.append(("st", p[(l, t, f)] ** 2 + q[(l, t, f)] ** 2))
        for n_, i in enumerate(buses):
            pd, qd = bus[i][0] / base, bus[i][1] / base
            gens_here = [g_ for g_, (b_, _) in gen.items() if b_ == i]
            arcs_here = [a for a in arcs if a[1] == i]
            # everything moves left; affine rows keep their constant in the set (JuMP normalisation), so the function
            # part is  sum(p) - sum(pg) [- theta_i: theta is a PARAMETER of the expression, not a constant]
            fp = sum(p[a] for a in arcs_here) - sum(pg[g_] for g_ in gens_here)
            fq = sum(q[a] for a in arcs_here) - sum(qg[g_] for g_ in gens_here)
            if th is not None:
                fp, fq = fp - R(th[n_]), fq - R(th[nbus + n_])
            rows.append(("bp", fp)); rows.append(("bq", fq))
        return [r_[1] if isinstance(r_, tuple) else r_ for r_ in rows]

    P.cons += stage(va0, vm0, pg0, qg0, p0, q0, None)
    for k in range(K):
        at = lambda d: {key: v[k] for key, v in d.items()}
        P.cons += stage(at(va), at(vm), at(pg), at(qg), at(p), at(q), theta[:, k])
    for i in gen:                                                             # ramping: pg0 - pg(k), qg0 - qg(k)
        P.cons += [pg0[i] - pg[i][k] for k in range(K)]
    for i in gen:
        P.cons += [qg0[i] - qg[i][k] for k in range(K)]
    rng2 = np.random.default_rng(17)
    x = rng2.uniform(-0.3, 0.3, len(P.vars))
    for i, s_ in enumerate(P.vars):                                          # voltage magnitudes around 1
        if str(s_).startswith("vm"):
            x[i] = rng2.uniform(0.92, 1.08)
    P.dump(x=x)


if __name__ == "__main__":
    opf_case3()
    param_function_problem()
    solve_tests()
    ode_5x5()
    quadrotor("fd", 5)
    quadrotor("oc", 4)
    pandemic()
    farmer()
    operators()
