"""Hand transcriptions of the BASELINE.json configurations.

Each builder restates, for one model file of the reference, what ``build_exa_core!``
(src/transform.jl:771-796) emits into the ExaCore: x/θ layout, one generator per constraint /
derivative approximation / collocation restriction / objective term, in the reference's emission
order.  Julia RNG streams cannot be reproduced here, so random supports are drawn from the same
distributions with ``numpy.random.default_rng(seed)``.

What cannot be verified offline (InfiniteOpt / JuMP are absent): the exact term order JuMP gives
quadratic/affine expressions, and the exact shape InfiniteOpt gives the orthogonal-collocation
rows.  Both only affect slot ORDER inside a generator, never the mathematical model.
"""
from __future__ import annotations

import numpy as np

from .core import ExaCore, Itr
from .expr import DataSource, abs2, cos, sin, tan


def trapezoid_coeffs(s: np.ndarray) -> np.ndarray:
    """InfiniteOpt's default integral over sorted supports (weights used by transform.jl:625-626)."""
    s = np.asarray(s, dtype=np.float64)
    c = np.zeros_like(s)
    d = np.diff(s)
    c[:-1] += d / 2
    c[1:] += d / 2
    return c


# ------------------------------------------------------------------------------------------------
def quadrotor(num_supports: int = 100, method: str = "oc", T_end: float = 60.0, device_side: bool = False) -> ExaCore:
    """ESCAPE34/quadrotor.jl:4-76 (``method='oc'``: OrthogonalCollocation(3) + piecewise-constant
    controls, config 3) and examples/quadrotor.jl:7-77 (``method='fd'``: default backward finite
    difference, config 1).  ``device_side=True`` (OC only): the time supports, the trapezoid weights and the parameter
    functions d1, d3, d5 are DESCRIBED (generated columns, tapes) and produced on the device by the engine instead of
    being computed here and uploaded (transform.jl:2-38,161-183,618-633 on the device)."""
    N = int(num_supports)
    pub = np.linspace(0.0, T_end, N)
    if method == "oc":  # one internal Lobatto node (midpoint) per interval: transform.jl:22
        T = 2 * N - 1
        ts = np.empty(T)
        ts[0::2] = pub
        ts[1::2] = 0.5 * (pub[:-1] + pub[1:])
    else:
        T = N
        ts = pub
    core = ExaCore(minimize=True)
    ds = DataSource()
    iota = np.arange(1, T + 1)
    if device_side and method == "oc":
        from .core import linspace_mid_col, trapezoid_col
        base = Itr(T, {"group_idx1": None}, {"ip1": linspace_mid_col(0.0, T_end, N)})
        ts = base.fps["ip1"]              # numpy statement of the generated column (bit-identical by construction)
    else:
        base = Itr(T, {"group_idx1": iota}, {"ip1": ts})                       # transform.jl:31

    # x layout (transform.jl:134-158): x[1:9], u[1:4] then the derivative variables ∂x[1:9]
    x = [core.add_var(T) for _ in range(9)]
    u = [core.add_var(T, start=0.0) for _ in range(4)]
    dx = [core.add_var(T) for _ in range(9)]
    # θ layout (transform.jl:161-183): parameter functions d1, d3, d5 evaluated at every support
    if device_side and method == "oc":
        d1 = core.add_par_function(sin((2 * np.pi) * ds.ip1 / T_end), base)
        d3 = core.add_par_function(2.0 * sin((4 * np.pi) * ds.ip1 / T_end), base)
        d5 = core.add_par_function(2.0 * (ds.ip1 / T_end), base)
    else:
        d1 = core.add_par(np.sin(2 * np.pi * ts / T_end))
        d3 = core.add_par(2 * np.sin(4 * np.pi * ts / T_end))
        d5 = core.add_par(2 * (ts / T_end))

    i = ds.group_idx1
    X = [None] + [v[i] for v in x]      # 1-based like the Julia source
    U = [None] + [v[i] for v in u]
    D = [None] + [v[i] for v in dx]

    # initial conditions x[i](0) == 0: point variables -> length-1 generators (transform.jl:439-440)
    for j in range(9):
        core.add_con(x[j][1], Itr.empty(), lcon=0.0, ucon=0.0)

    # dynamics, one generator per @constraint over the time iterator (transform.jl:441-442)
    def c(expr):
        core.add_con(expr, base, lcon=0.0, ucon=0.0)

    c(D[1] + (-1.0) * X[2])                                                       # affine (:343-358)
    c(D[2] - (U[1] * cos(X[7]) * sin(X[8]) * cos(X[9]) + U[1] * sin(X[7]) * sin(X[9])))
    c(D[3] + (-1.0) * X[4])
    c(D[4] - (U[1] * cos(X[7]) * sin(X[8]) * sin(X[9]) - U[1] * sin(X[7]) * cos(X[9])))
    c(D[5] + (-1.0) * X[6])
    c(D[6] - (U[1] * cos(X[7]) * cos(X[8]) - 9.8))
    c(D[7] - (U[2] * cos(X[7]) / cos(X[8]) + U[3] * sin(X[7]) / cos(X[8])))
    c(D[8] - (((-1.0) * U[2]) * sin(X[7]) + U[3] * cos(X[7])))
    c(D[9] - (U[2] * cos(X[7]) * tan(X[8]) + U[3] * sin(X[7]) * tan(X[8]) + U[4]))

    # derivative approximations, one generator per derivative variable (transform.jl:511-562)
    if method == "oc":
        h = np.diff(pub)
        lb = np.repeat(np.arange(1, T, 2), 2)                 # index of the interval's lower bound
        node = lb + np.tile([1, 2], N - 1)                    # row centre: internal node, then ub
        m1 = np.empty(2 * (N - 1)); m2 = np.empty(2 * (N - 1))
        m1[0::2], m2[0::2] = 0.75 * h, -0.25 * h              # M = M2 * inv(M1), see DESIGN.md
        m1[1::2], m2[1::2] = h, 0.0
        oc = Itr(2 * (N - 1), {"group_idx1": node, "d_lb": lb}, {"ip1": ts[node - 1], "d_arg1": m1, "d_arg2": m2})
        ii, l = ds.group_idx1.idx(), ds.d_lb.idx()
        for j in range(9):
            core.add_con(ds.d_arg1 * dx[j][l + 1] + ds.d_arg2 * dx[j][l + 2] - x[j][ii] + x[j][l], oc)
    else:  # backward finite difference: Δt·d[i] − y[i] + y[i−1] = 0, i = 2..T
        idx = np.arange(2, T + 1)
        fd = Itr(T - 1, {"group_idx1": idx}, {"ip1": ts[1:], "d_arg1": np.diff(ts)})
        ii = ds.group_idx1.idx()
        for j in range(9):
            core.add_con(ds.d_arg1 * dx[j][ii] - x[j][ii] + x[j][ii - 1], fd)

    # collocation restrictions for the piecewise-constant controls (transform.jl:565-601)
    if method == "oc":
        ubs = np.arange(3, T + 1, 2)
        pts = np.arange(2, T, 2)
        col = Itr(N - 1, {"i1": ubs, "i2": pts}, {})
        for j in range(4):
            core.add_con(u[j][ds.i1] - u[j][ds.i2], col)

    # objective: one measure -> one generator c·(quadratic integrand) (transform.jl:693-702)
    if device_side and method == "oc":
        mitr = Itr(T, {"group_idx1": None}, {"ip1": linspace_mid_col(0.0, T_end, N), "c": trapezoid_col("ip1")})
    else:
        mitr = Itr(T, {"group_idx1": iota}, {"c": trapezoid_coeffs(ts), "ip1": ts})
    P1, P3, P5 = d1[i], d3[i], d5[i]
    quad = (abs2(X[1]) + (-2.0) * X[1] * P1 + abs2(P1)
            + abs2(X[3]) + (-2.0) * X[3] * P3 + abs2(P3)
            + abs2(X[5]) + (-2.0) * X[5] * P5 + abs2(P5)
            + abs2(X[7]) + abs2(X[8]) + abs2(X[9])
            + 0.1 * abs2(U[1]) + 0.1 * abs2(U[2]) + 0.1 * abs2(U[3]) + 0.1 * abs2(U[4]))
    core.add_obj(ds.c * quad, mitr)
    return core


# ------------------------------------------------------------------------------------------------
def ode_5x5(nt: int = 5, nx: int = 5) -> ExaCore:
    """The reference's known-answer model (test/madnlp.jl:4-11 ≡ test/ipopt.jl:4-11; SURVEY.md
    Appendix C): nvar = 51, ncon = 70, optimal objective −12.7846."""
    core = ExaCore(minimize=True)
    ds = DataSource()
    ts = np.linspace(0, 1, nt)
    xs = np.linspace(-1, 1, nx)
    it_t = Itr(nt, {"group_idx1": np.arange(1, nt + 1)}, {"ip1": ts})
    it_x = Itr(nx, {"group_idx2": np.arange(1, nx + 1)}, {"ip2": xs})
    z = core.add_var(1, start=10.0)                       # finite variables first (transform.jl:780)
    y = core.add_var(nt, nx, lvar=0.0)
    dy = core.add_var(nt, nx)
    both = Itr.product([it_t, it_x])
    i, j = ds.group_idx1, ds.group_idx2
    # ∂y/∂t == sin(y) + z + 1.2
    core.add_con(dy[i, j] - (sin(y[i, j]) + z[1] + 1.2), both, 0.0, 0.0)
    # y + z <= 42 + t   ->   y + z − t <= 42
    core.add_con(y[i, j] + z[1] + (-1.0) * ds.ip1, both, -np.inf, 42.0)
    fd = Itr(nt - 1, {"group_idx1": np.arange(2, nt + 1)}, {"ip1": ts[1:], "d_arg1": np.diff(ts)})
    ii = ds.group_idx1.idx()
    core.add_con(ds.d_arg1 * dy[ii, j] - y[ii, j] + y[ii - 1, j], Itr.product([fd, it_x]), 0.0, 0.0)
    # ∫(∫(y², t) + 2z, x): the 2z term moves inside the inner measure (transform.jl:663-686)
    wt, wx = trapezoid_coeffs(ts), trapezoid_coeffs(xs)
    K = nt * nx
    # product of measure iterators, INNER (t) measure first (transform.jl:670): c = c_t·c_x
    m = Itr(K, {"group_idx1": np.tile(np.arange(1, nt + 1), nx), "group_idx2": np.repeat(np.arange(1, nx + 1), nt)},
            {"c": np.outer(wx, wt).reshape(-1), "ip1": np.tile(ts, nx), "ip2": np.repeat(xs, nt)})
    core.add_obj(ds.c * (abs2(y[i, j]) + 2.0 * z[1]), m)
    return core


# ------------------------------------------------------------------------------------------------
def pandemic(num_supports: int = 100, num_scenarios: int = 4, seed: int = 0, xi=None) -> ExaCore:
    """ESCAPE34/pandemic.jl:4-34 — SEIR optimal control, backward FD in t, scenarios ξ~U(0.1,0.6) (``xi``: explicit
    scenario supports, e.g. the ones a Julia run drew — julia/dump_golden.jl writes them next to its dump)."""
    gamma, beta, Npop = 0.303, 0.727, 1e5
    extra = np.array([0.001, 0.002, 0.004, 0.008, 0.02, 0.04, 0.08, 0.2, 0.4, 0.8])
    ts = np.unique(np.concatenate([np.linspace(0, 200, int(num_supports)), extra]))
    T, S = len(ts), int(num_scenarios)
    xi = np.random.default_rng(seed).uniform(0.1, 0.6, S) if xi is None else np.asarray(xi, dtype=np.float64)
    assert len(xi) == S
    core = ExaCore(minimize=True)
    ds = DataSource()
    it_t = Itr(T, {"group_idx1": np.arange(1, T + 1)}, {"ip1": ts})
    it_s = Itr(S, {"group_idx2": np.arange(1, S + 1)}, {"ip2": xi})
    both = Itr.product([it_t, it_s])
    s = core.add_var(T, S, lvar=0.0)
    e = core.add_var(T, S, lvar=0.0)
    ii = core.add_var(T, S, lvar=0.0)
    r = core.add_var(T, S, lvar=0.0)
    u = core.add_var(T, lvar=0.0, uvar=0.8, start=0.2)
    ds_, de_, di_, dr_ = (core.add_var(T, S) for _ in range(4))
    i, j = ds.group_idx1, ds.group_idx2
    # initial conditions: semi-infinite variables s(0, ξ) ... over the scenario iterator (:312-319)
    core.add_con(s[1, j], it_s, 1 - 1 / Npop, 1 - 1 / Npop)
    core.add_con(e[1, j], it_s, 1 / Npop, 1 / Npop)
    core.add_con(ii[1, j], it_s, 0.0, 0.0)
    core.add_con(r[1, j], it_s, 0.0, 0.0)
    S_, E_, I_, R_, U_ = s[i, j], e[i, j], ii[i, j], r[i, j], u[i]
    core.add_con(ds_[i, j] - (-(1 - U_) * beta * S_ * I_), both, 0.0, 0.0)
    core.add_con(de_[i, j] - ((1 - U_) * beta * S_ * I_ - ds.ip2 * E_), both, 0.0, 0.0)
    core.add_con(di_[i, j] - (ds.ip2 * E_ - gamma * I_), both, 0.0, 0.0)
    core.add_con(dr_[i, j] + (-gamma) * I_, both, 0.0, 0.0)
    core.add_con(I_, both, -np.inf, 0.02)
    fd = Itr(T - 1, {"group_idx1": np.arange(2, T + 1)}, {"ip1": ts[1:], "d_arg1": np.diff(ts)})
    fdb = Itr.product([fd, it_s])
    k = ds.group_idx1.idx()
    for v, dv in ((s, ds_), (e, de_), (ii, di_), (r, dr_)):
        core.add_con(ds.d_arg1 * dv[k, j] - v[k, j] + v[k - 1, j], fdb, 0.0, 0.0)
    m = Itr(T, {"group_idx1": np.arange(1, T + 1)}, {"c": trapezoid_coeffs(ts), "ip1": ts})
    core.add_obj(ds.c * u[ds.group_idx1], m)
    return core


# ------------------------------------------------------------------------------------------------
def farmer(num_scenarios: int = 1000, seed: int = 42) -> ExaCore:
    """examples/2stage_example.jl:6-37 — two-stage stochastic farmer LP (Hessian is empty)."""
    K = int(num_scenarios)
    rng = np.random.default_rng(seed)
    xi = np.stack([rng.uniform(0, 5, K), rng.uniform(0, 5, K), rng.uniform(10, 30, K)])
    alpha, beta, lam, d = [150, 230, 260], [238, 210, 0], [170, 150, 36], [200, 240, 0]
    xbar, wbar3, ybar3 = 500, 6000, 0
    core = ExaCore(minimize=True)
    ds = DataSource()
    # one dependent-parameter group ξ[1:3]: a single iterator carrying all three values (:12-13,:31)
    it = Itr(K, {"group_idx1": np.arange(1, K + 1)}, {"dp11": xi[0], "dp12": xi[1], "dp13": xi[2]})
    x = [core.add_var(1, lvar=0.0, uvar=xbar) for _ in range(3)]
    y = [core.add_var(K, lvar=0.0) for _ in range(3)]
    w = [core.add_var(K, lvar=0.0) for _ in range(3)]
    i = ds.group_idx1
    core.add_con(x[0][1] + x[1][1] + x[2][1], Itr.empty(), -np.inf, xbar)
    for c_ in range(3):
        core.add_con(ds[f"dp1{c_ + 1}"] * x[c_][1] + y[c_][i] + (-1.0) * w[c_][i], it, d[c_], np.inf)
    core.add_con(w[2][i], it, -np.inf, wbar3)
    core.add_con(y[2][i], it, -np.inf, ybar3)
    # objective α'x + 𝔼(β'y − λ'w): three finite terms + one measure generator with c = 1/K
    for c_ in range(3):
        core.add_obj(alpha[c_] * x[c_][1], Itr.empty())
    m = Itr(K, {"group_idx1": np.arange(1, K + 1)}, {"c": np.full(K, 1.0 / K), "dp11": xi[0], "dp12": xi[1], "dp13": xi[2]})
    core.add_obj(ds.c * (beta[0] * y[0][i] + beta[1] * y[1][i] + (-lam[0]) * w[0][i] + (-lam[1]) * w[1][i]
                         + (-lam[2]) * w[2][i]), m)
    return core


# ------------------------------------------------------------------------------------------------
def rosenbrock_param(p1: float = 100.0, p2: float = 1.0, nt: int = 3):
    """test/solve.jl:134-156 ("Parameter updates"): known answers 306.4999755050365 and, after the
    in-place update p1=90, p2=1.3, 276.26497794903645.  Returns (core, P1, P2) with the finite
    parameter handles (θ layout: one entry per finite parameter, transform.jl:120-131)."""
    core = ExaCore(minimize=True)
    ds = DataSource()
    ts = np.linspace(0, 1, nt)
    P1 = core.add_par([p1])
    P2 = core.add_par([p2])
    x1 = core.add_var(nt)
    x2 = core.add_var(nt)
    it = Itr(nt, {"group_idx1": np.arange(1, nt + 1)}, {"ip1": ts})
    i = ds.group_idx1
    core.add_con(x1[i], it, -np.inf, 0.5)
    core.add_con(x2[i], it, -np.inf, 3.0)
    core.add_con(x1[i] * x2[i], it, 1.0, np.inf)
    core.add_con(abs2(x2[i]) + x1[i], it, 0.0, np.inf)
    m = Itr(nt, {"group_idx1": np.arange(1, nt + 1)}, {"c": trapezoid_coeffs(ts), "ip1": ts})
    # p1 * ∫((x2 − x1²)², t): quad term (p1, measure) -> coef·p1 moves inside (transform.jl:758-762)
    core.add_obj(ds.c * (P1[1] * (((-1.0) * abs2(x1[i]) + x2[i]) ** 2)), m)
    # ∫((p2 − x1)², t): affine² expands to p2² − 2·p2·x1 + x1²
    core.add_obj(ds.c * (abs2(P2[1]) + (-2.0) * P2[1] * x1[i] + abs2(x1[i])), m)
    return core, P1, P2


def param_function_model(offset: float = 0.2, pf1=np.sin, nt: int = 3, ns: int = 3):
    """test/solve.jl:158-209 ("Parameter function updates"): known answers 0.48292223509341475
    (pf1 = sin, pf2 = sin(t)·s + 0.2) and 0.8155916466182952 (pf1 = cos, pf2 = sin(t)·s + 0.8).
    Returns (core, PF1, PF2); θ layout is column-major, first group fastest
    (test/transcription.jl:151-167, test/solve.jl:191)."""
    core = ExaCore(minimize=True)
    ds = DataSource()
    ts = np.linspace(0, 1, nt)
    ss = np.linspace(2, 3, ns)
    it_t = Itr(nt, {"group_idx1": np.arange(1, nt + 1)}, {"ip1": ts})
    it_s = Itr(ns, {"group_idx2": np.arange(1, ns + 1)}, {"ip2": ss})
    both = Itr.product([it_t, it_s])
    v = core.add_var(nt, lvar=0.0, uvar=100.0)
    z = core.add_var(nt, ns, lvar=0.0, uvar=100.0)
    PF1 = core.add_par(pf1(ts))
    PF2 = core.add_par(np.sin(ts)[:, None] * ss[None, :] + offset)
    i, j = ds.group_idx1, ds.group_idx2
    core.add_con(v[i] + PF1[i], it_t, -np.inf, 100.0)                         # c1
    core.add_con(PF1[i] * PF2[i, j] + 2.0 * v[i], both, -np.inf, 100.0)      # c2: quad term, then aff
    core.add_con(v[i] + (-0.5) * PF2[i, j], both, 0.0, np.inf)               # c3
    s25 = int(np.argmin(np.abs(ss - 2.5))) + 1                               # support_to_index[s = 2.5]
    core.add_con(PF2[i, j] * PF1[i] + z[i, s25], both, -np.inf, 40.0)         # c4: semi-infinite z(t, 2.5)
    mt = Itr(nt, {"group_idx1": np.arange(1, nt + 1)}, {"c": trapezoid_coeffs(ts), "ip1": ts})
    core.add_obj(ds.c * (v[i] * PF1[i]), mt)
    wt, ws = trapezoid_coeffs(ts), trapezoid_coeffs(ss)
    K = nt * ns
    m2 = Itr(K, {"group_idx1": np.tile(np.arange(1, nt + 1), ns), "group_idx2": np.repeat(np.arange(1, ns + 1), nt)},
             {"c": np.outer(ws, wt).reshape(-1)})
    core.add_obj(ds.c * (0.5 * z[i, j] * PF2[i, j]), m2)
    return core, PF1, PF2
