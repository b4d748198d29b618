"""The reference's model files written against the mini modelling layer (``infopt.py``) — the same
statements, in the same order, as the Julia sources they cite — and lowered by ``transform.py``.
``models.py`` holds hand transcriptions of the same models; tests check both routes agree."""
from __future__ import annotations

import numpy as np

from . import infopt as io
from .infopt import InfiniteModel, cos, sin, tan


def ode_5x5() -> InfiniteModel:
    """test/madnlp.jl:4-11"""
    m = InfiniteModel()
    t = m.infinite_parameter(0, 1, num_supports=5)
    x = m.infinite_parameter(-1, 1, num_supports=5)
    y = m.variable(t, x, lb=0.0)
    z = m.variable(start=10.0)
    m.objective("Min", m.integral(m.integral(y ** 2, t) + 2 * z, x))
    m.constraint(m.deriv(y, t), "==", sin(y) + z + 1.2)
    m.constraint(y + z, "<=", 42 + t)
    return m


def solve_test_problem(variant: int = -1) -> InfiniteModel:
    """test/solve.jl:2-14 ("Test Problem 1", ``variant=-1``: point variable, DomainRestriction, derivative of a
    semi-infinite variable) and :48-59 / :72-77 ("Test Problem 2", ``variant=0`` and its alternative objectives 1..4).
    The reference checks these differentially against JuMP's TranscriptionBackend (solve.jl:19-26, 61-68, 79-93)."""
    m = InfiniteModel()
    t = m.infinite_parameter(0, 1, num_supports=5)
    x = m.infinite_parameter(-1, 1, num_supports=5)
    y = m.variable(t, x, lb=0.0)
    z = m.variable(start=10.0)
    inner = m.integral(y ** 2, t)
    if variant == -1:
        m.objective("Min", m.integral(m.integral(y ** 2, t), x) + 2 * y(0, 1))
    elif variant == 0:
        m.objective("Min", m.integral(inner + 2 * z, x) + 2 * y(0, 1))
    elif variant == 1:
        m.objective("Min", m.integral(inner + 2 * z ** 2, x) + 2 * y(0, 1))
    elif variant == 2:
        m.objective("Min", m.integral(inner + sin(z ** 2), x))
    elif variant == 3:
        m.objective("Min", m.integral(inner * cos(z), x))
    else:
        m.objective("Min", m.integral(z * (inner + z ** 3), x))
    m.constraint(m.deriv(y, t), "==", sin(y) + z + 1.2)
    if variant == -1:
        m.constraint(y + z, "<=", 42 + t, restriction=lambda s: 0 <= s <= 0.5, restriction_prefs=(t,))
    else:
        m.constraint(y + z, "<=", 42 + t)
    m.constraint(m.deriv(y(0, x), x), "==", 5)
    return m


def rosenbrock_param(p1v=100.0, p2v=1.0):
    """test/solve.jl:134-143"""
    m = InfiniteModel()
    t = m.infinite_parameter(0, 1, num_supports=3)
    p1 = m.finite_parameter(p1v)
    p2 = m.finite_parameter(p2v)
    x = [m.variable(t), m.variable(t)]
    m.objective("Min", p1 * m.integral((x[1] - x[0] ** 2) ** 2, t) + m.integral((p2 - x[0]) ** 2, t))
    for i, ub in enumerate([0.5, 3.0]):
        m.constraint(x[i], "<=", ub)
    m.constraint(x[0] * x[1], ">=", 1.0)
    m.constraint(x[0] + x[1] ** 2, ">=", 0.0)
    return m, p1, p2


def param_function_model(offset=0.2, pf1=np.sin):
    """test/solve.jl:167-181"""
    m = InfiniteModel()
    t = m.infinite_parameter(0, 1, num_supports=3)
    s = m.infinite_parameter(2, 3, num_supports=3)
    v = m.variable(t, lb=0, ub=100)
    z = m.variable(t, s, lb=0, ub=100)
    f1 = m.parameter_function(lambda tt: pf1(tt), t)
    f2 = m.parameter_function(lambda tt, ss: np.sin(tt) * ss + offset, t, s)
    m.constraint(v + f1, "<=", 100)
    m.constraint(v * 2 + f1 * f2, "<=", 100)
    m.constraint(v, ">=", 0.5 * f2)
    m.constraint(z(t, 2.5) + f2 * f1, "<=", 40)
    m.objective("Min", m.integral(v * f1, t) + m.integral(m.integral(0.5 * z * f2, t), s))
    return m, f1, f2


def param_function_problem(ti: float = 0.2) -> InfiniteModel:
    """test/solve.jl:97-131 ("Parameter Function Problem"), including c5 — a measure of a parameter function inside a
    constraint, which the reference expands inline with a warning (transform.jl:430-435)"""
    def param_func2(t, s):
        return np.cos(t) * s - ti if t <= 0.5 else np.sin(t) * s + ti

    m = InfiniteModel()
    t = m.infinite_parameter(0, 1, num_supports=5)
    s = m.infinite_parameter(2, 3, num_supports=5)
    v = m.variable(t, lb=0, ub=100)
    z = m.variable(t, s, lb=0, ub=100)
    pf = m.parameter_function(lambda tt: np.sin(tt), t)
    pf2 = m.parameter_function(param_func2, t, s)
    m.constraint(v + pf, "<=", 100)
    m.constraint(v * 2 + pf * pf2, "<=", 100)
    m.constraint(v, ">=", 0.2 * pf2)
    m.constraint(z(t, 2.5) + pf2 * pf, "<=", 40)
    m.constraint(v * m.integral(pf2, s), "<=", 100)
    m.objective("Min", m.integral(v * pf, t) + m.integral(m.integral(0.5 * z * pf2, t), s))
    return m


def pandemic(num_supports=100, num_scenarios=4, seed=0) -> InfiniteModel:
    """ESCAPE34/pandemic.jl:4-34"""
    gamma, beta, N = 0.303, 0.727, 1e5
    extra_ts = [0.001, 0.002, 0.004, 0.008, 0.02, 0.04, 0.08, 0.2, 0.4, 0.8]
    m = InfiniteModel()
    t = m.infinite_parameter(0, 200, num_supports=num_supports)
    xi = m.infinite_parameter(supports=np.random.default_rng(seed).uniform(0.1, 0.6, num_scenarios))
    # the reference keeps scenario supports in draw order; numpy's unique() in infinite_parameter sorts them
    m.public[xi.group - 1] = np.random.default_rng(seed).uniform(0.1, 0.6, num_scenarios)
    m.supports[xi.group - 1] = m.public[xi.group - 1]
    m.add_supports(t, extra_ts)
    s = m.variable(t, xi, lb=0); e = m.variable(t, xi, lb=0); i = m.variable(t, xi, lb=0); r = m.variable(t, xi, lb=0)
    u = m.variable(t, lb=0, ub=0.8, start=0.2)
    m.objective("Min", m.integral(u, t))
    m.constraint(s(0, xi), "==", 1 - 1 / N)
    m.constraint(e(0, xi), "==", 1 / N)
    m.constraint(i(0, xi), "==", 0)
    m.constraint(r(0, xi), "==", 0)
    m.constraint(m.deriv(s, t), "==", -(1 - u) * beta * s * i)
    m.constraint(m.deriv(e, t), "==", (1 - u) * beta * s * i - xi * e)
    m.constraint(m.deriv(i, t), "==", xi * e - gamma * i)
    m.constraint(m.deriv(r, t), "==", gamma * i)
    m.constraint(i, "<=", 0.02)
    return m


def quadrotor(num_supports=100, method="oc") -> InfiniteModel:
    """ESCAPE34/quadrotor.jl:4-76 (method='oc') / examples/quadrotor.jl:7-77 (method='fd')"""
    n, p, T = 9, 4, 60
    m = InfiniteModel()
    dm = io.OrthogonalCollocation(3) if method == "oc" else io.FiniteDifference()
    t = m.infinite_parameter(0, T, num_supports=num_supports, derivative_method=dm)
    d1 = m.parameter_function(lambda tt: np.sin(2 * np.pi * tt / T), t)
    d3 = m.parameter_function(lambda tt: 2 * np.sin(4 * np.pi * tt / T), t)
    d5 = m.parameter_function(lambda tt: 2 * (tt / T), t)
    x = [None] + [m.variable(t) for _ in range(n)]
    u = [None] + [m.variable(t, start=0) for _ in range(p)]
    m.objective("Min", m.integral(
        (x[1] - d1) ** 2 + (x[3] - d3) ** 2 + (x[5] - d5) ** 2 + x[7] ** 2 + x[8] ** 2 + x[9] ** 2
        + 0.1 * (u[1] ** 2 + u[2] ** 2 + u[3] ** 2 + u[4] ** 2), t))
    for i in range(1, n + 1):
        m.constraint(x[i](0), "==", 0)
    D = lambda v: m.deriv(v, t)
    mul = io.nl("*"); add = io.nl("+"); sub = io.nl("-"); div = io.nl("/")
    m.constraint(D(x[1]), "==", x[2])
    m.constraint(D(x[2]), "==", add(mul(u[1], cos(x[7]), sin(x[8]), cos(x[9])), mul(u[1], sin(x[7]), sin(x[9]))))
    m.constraint(D(x[3]), "==", x[4])
    m.constraint(D(x[4]), "==", sub(mul(u[1], cos(x[7]), sin(x[8]), sin(x[9])), mul(u[1], sin(x[7]), cos(x[9]))))
    m.constraint(D(x[5]), "==", x[6])
    m.constraint(D(x[6]), "==", sub(mul(u[1], cos(x[7]), cos(x[8])), 9.8))
    m.constraint(D(x[7]), "==", add(div(mul(u[2], cos(x[7])), cos(x[8])), div(mul(u[3], sin(x[7])), cos(x[8]))))
    m.constraint(D(x[8]), "==", add(mul(-u[2], sin(x[7])), mul(u[3], cos(x[7]))))
    m.constraint(D(x[9]), "==", add(mul(u[2], cos(x[7]), tan(x[8])), mul(u[3], sin(x[7]), tan(x[8])), u[4]))
    if method == "oc":
        for j in range(1, p + 1):
            m.constant_over_collocation(u[j], t)
    return m


def farmer(num_scenarios=1000, seed=42) -> InfiniteModel:
    """examples/2stage_example.jl:6-37"""
    K = num_scenarios
    rng = np.random.default_rng(seed)
    supports = np.stack([rng.uniform(0, 5, K), rng.uniform(0, 5, K), rng.uniform(10, 30, K)])
    alpha, beta, lam, d = [150, 230, 260], [238, 210, 0], [170, 150, 36], [200, 240, 0]
    xbar, wbar3, ybar3 = 500, 6000, 0
    m = InfiniteModel()
    xi = m.dependent_parameters(supports)
    x = [m.variable(lb=0, ub=xbar) for _ in range(3)]
    y = [m.variable(*xi, lb=0) for _ in range(3)]
    w = [m.variable(*xi, lb=0) for _ in range(3)]
    second = sum(beta[c] * y[c] for c in range(3) if beta[c]) - sum(lam[c] * w[c] for c in range(3))
    m.objective("Min", sum(alpha[c] * x[c] for c in range(3)) + m.expect(second, xi))
    m.constraint(x[0] + x[1] + x[2], "<=", xbar)
    for c in range(3):
        m.constraint(xi[c] * x[c] + y[c] - w[c], ">=", d[c])
    m.constraint(w[2], "<=", wbar3)
    m.constraint(y[2], "<=", ybar3)
    return m
