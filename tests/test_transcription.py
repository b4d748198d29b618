"""The Python restatement of src/transform.jl (transform.py on the infopt.py modelling layer).

Part 1 mirrors the reference's own unit tests of the builders (test/transcription.jl): iterator
count, finite-variable mapping / start / bounds, infinite-variable lengths and function-valued
bounds, derivative-variable lengths, semi-infinite / point patches, θ layout of finite parameters and
parameter functions (column-major, first group fastest), restricted iterator length.
Part 2 checks that lowering the reference's model files through this route yields the SAME ExaCore
(dimensions, sparsity, values at a random point — through the oracle) as the hand transcriptions."""
import warnings

import numpy as np
import pytest

from iexa_b200 import infmodels, infopt as io, models
from iexa_b200.transform import ExaMappingData, exa_core
from conftest import assert_close, eval_point


def _base():
    # the hand-made model of test/transcription.jl:1-24 (same shapes: t has 4 supports, x 5, ξ[1:2] 3)
    m = io.InfiniteModel()
    t = m.infinite_parameter(0, 1, num_supports=4)
    x = m.infinite_parameter(-1, 1, num_supports=5)
    xi = m.dependent_parameters(np.array([[0.1, 0.5, 0.9], [1.0, 2.0, 3.0]]))
    z = m.variable(lb=-2.0, ub=7.0, start=1.5)
    y = m.variable(t, x, lb=0.0, ub=lambda tt, xx: 10 + tt + xx, start=lambda tt, xx: tt * xx)
    q = m.variable(*xi, fix=3.0)
    w = m.variable(t)
    dy = m.deriv(y, t)
    return m, (t, x, xi, z, y, q, w, dy)


def test_base_iterators_and_variable_layout():
    m, (t, x, xi, z, y, q, w, dy) = _base()
    yt0 = y(0, x); yt0.info.ub = 4.0                 # semi-infinite with an upper bound patch
    pt = y(1, -1); pt.info.start = 9.0; pt.info.lb = 0.5   # point variable patch
    m.constraint(y + z, "<=", 1.0)
    core, data = exa_core(m)
    assert len(data.base_itrs) == 3                                        # transcription.jl:27-28
    assert data.group_alias == ["group_idx1", "group_idx2", "group_idx3"]
    assert data.param_alias[t] == "ip1" and data.param_alias[xi[1]] == "dp32"   # transform.jl:12-16
    assert data.finvar_mappings[z] == 1                                    # finite variables first (:33-37)
    assert core.x0_vec[0] == 1.5 and core.lvar_vec[0] == -2.0 and core.uvar_vec[0] == 7.0
    Y = data.infvar_mappings[y]
    assert Y.size == (4, 5) and Y.offset == 1 and Y.length == 20            # :43-57
    ub = core.uvar_vec[1:21].reshape(4, 5, order="F")
    ts, xs = np.linspace(0, 1, 4), np.linspace(-1, 1, 5)
    assert np.allclose(ub[1:, 1:], (10 + ts[:, None] + xs[None, :])[1:, 1:])  # function-valued bound, column-major
    x0 = core.x0_vec[1:21].reshape(4, 5, order="F")
    assert np.allclose(x0[:3, 1:], (ts[:, None] * xs[None, :])[:3, 1:])
    Q = data.infvar_mappings[q]
    assert Q.size == (3,) and (core.lvar_vec[Q.offset:Q.offset + 3] == 3.0).all() and (core.uvar_vec[Q.offset:Q.offset + 3] == 3.0).all()
    assert data.infvar_mappings[dy].size == (4, 5)                          # derivative variables last (:59-62)
    assert data.infvar_mappings[dy].offset == 1 + 20 + 3 + 4
    # semi-infinite patch: y(0, x) <= 4 for every x (:66-87); point patch: y(1, -1)
    assert (ub[0, :] == 4.0).all()
    assert data.finvar_mappings[pt] == Y.index(4, 1) and core.x0_vec[Y.index(4, 1) - 1] == 9.0 and core.lvar_vec[Y.index(4, 1) - 1] == 0.5


def test_parameter_layouts():
    m, (t, x, xi, z, y, q, w, dy) = _base()
    p1 = m.finite_parameter(42.0); p2 = m.finite_parameter(-3.0)
    f = m.parameter_function(lambda tt, xx: 10 * tt + xx, t, x)
    m.constraint(y + p1 * z + f, "<=", p2)
    core, data = exa_core(m)
    assert list(core.theta_vec[:2]) == [42.0, -3.0]                         # one entry per finite parameter (:108-126)
    F = data.param_mappings[f]
    assert F.offset == 2 and F.size == (4, 5)
    ts, xs = np.linspace(0, 1, 4), np.linspace(-1, 1, 5)
    assert np.allclose(core.theta_vec[2:], (10 * ts[:, None] + xs[None, :]).reshape(-1, order="F"))   # first group fastest (:151-167)


def test_domain_restriction_filters_the_iterator():
    m, (t, x, xi, z, y, q, w, dy) = _base()
    m.constraint(w, "<=", 1.0, restriction=lambda tt: tt <= 0.5, restriction_prefs=(t,))
    core, data = exa_core(m)
    assert core.cons[0].itr.K == 2                                         # transcription.jl:215-217
    assert list(core.cons[0].itr.ints["group_idx1"]) == [1, 2]


def test_objective_heuristics_warn_only_when_expanding():
    m, (t, x, xi, z, y, q, w, dy) = _base()
    m.objective("Min", m.integral(m.integral(y ** 2, t) + 2 * z, x))         # terms move inside: no warning (:186-208)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        core, _ = exa_core(m)
    assert len(core.objs) == 1 and core.objs[0].itr.K == 20
    m2, (t, x, xi, z, y, q, w, dy) = _base()
    inner = m2.integral(y, t)
    m2.objective("Min", m2.integral(io.sin(inner), x))                       # nonlinear in the inner measure: expand + warn
    with pytest.warns(UserWarning, match="Unable to convert objective measures"):
        core2, _ = exa_core(m2)
    assert core2.objs[0].itr.K == 5


def test_unsupported_operator_errors_like_the_reference():
    m, (t, x, xi, z, y, q, w, dy) = _base()
    m.constraint(io.nl("erf")(y), "<=", 1.0)
    with pytest.raises(ValueError, match="does not support the nonlinear operator `erf`"):   # operators.jl:50-53
        exa_core(m)


PAIRS = {
    "ode_5x5": (lambda: exa_core(infmodels.ode_5x5())[0], lambda: models.ode_5x5()),
    "rosenbrock_param": (lambda: exa_core(infmodels.rosenbrock_param()[0])[0], lambda: models.rosenbrock_param()[0]),
    "param_function": (lambda: exa_core(infmodels.param_function_model()[0])[0], lambda: models.param_function_model()[0]),
    "pandemic": (lambda: exa_core(infmodels.pandemic(6, 3))[0], lambda: models.pandemic(6, 3)),
    "quadrotor_fd": (lambda: exa_core(infmodels.quadrotor(9, "fd"))[0], lambda: models.quadrotor(9, "fd")),
    "quadrotor_oc": (lambda: exa_core(infmodels.quadrotor(7, "oc"))[0], lambda: models.quadrotor(7, "oc")),
    "farmer": (lambda: exa_core(infmodels.farmer(9))[0], lambda: models.farmer(9)),
}


@pytest.mark.parametrize("name", list(PAIRS))
def test_lowering_matches_hand_transcription(name):
    from oracle.oracle import OracleModel
    a, b = PAIRS[name][0](), PAIRS[name][1]()
    assert (a.nvar, a.ncon, a.npar) == (b.nvar, b.ncon, b.npar)
    assert np.array_equal(a.x0_vec, b.x0_vec) and np.array_equal(a.lvar_vec, b.lvar_vec) and np.array_equal(a.uvar_vec, b.uvar_vec)
    assert np.allclose(a.theta_vec, b.theta_vec, rtol=0, atol=1e-15)
    oa, ob = OracleModel(a), OracleModel(b)
    assert np.array_equal(oa.lcon, ob.lcon) and np.array_equal(oa.ucon, ob.ucon)
    assert (oa.nnzj, oa.nnzh) == (ob.nnzj, ob.nnzh)
    x, y = eval_point(b, seed=2)
    assert_close(oa.obj(x), ob.obj(x), "obj"); assert_close(oa.cons(x), ob.cons(x), "cons")
    assert_close(oa.grad(x), ob.grad(x), "grad")
    ra, ca = oa.jac_structure(); rb, cb = ob.jac_structure()
    Ja = np.zeros((oa.ncon, oa.nvar)); np.add.at(Ja, (ra - 1, ca - 1), oa.jac_coord(x))
    Jb = np.zeros((ob.ncon, ob.nvar)); np.add.at(Jb, (rb - 1, cb - 1), ob.jac_coord(x))
    assert np.allclose(Ja, Jb, rtol=1e-13, atol=1e-14)
    ha, hb = oa.hess_structure(), ob.hess_structure()
    Ha = np.zeros((oa.nvar, oa.nvar)); np.add.at(Ha, (ha[0] - 1, ha[1] - 1), oa.hess_coord(x, y, 0.7))
    Hb = np.zeros((ob.nvar, ob.nvar)); np.add.at(Hb, (hb[0] - 1, hb[1] - 1), ob.hess_coord(x, y, 0.7))
    assert np.allclose(Ha, Hb, rtol=1e-13, atol=1e-14)


def test_constrained_measure_is_expanded_inline_with_a_warning():
    """transform.jl:430-435: a measure inside a constraint is expanded into the explicit weighted sum
    (point variables), with the reference's warning; values checked against a hand-written model."""
    import iexa_b200 as ex
    from oracle.oracle import OracleModel
    m = io.InfiniteModel()
    t = m.infinite_parameter(0, 1, num_supports=5)
    y = m.variable(t, start=1.0)
    z = m.variable(start=2.0)
    m.constraint(m.integral(y ** 2, t) + z, "<=", 3.0)
    m.objective("Min", z)
    with pytest.warns(UserWarning, match="Constrained measures can lead to poor performance"):
        core, data = exa_core(m)
    assert core.ncon == 1 and core.cons[0].itr.K == 1                      # a finite constraint over [(;)]
    om = OracleModel(core)
    assert om.nnzj == 6                                                    # z + the 5 point variables y(t_k)
    x = np.array([2.0, 0.3, -0.5, 0.7, 1.1, -0.2])
    w = np.array([0.125, 0.25, 0.25, 0.25, 0.125])
    assert abs(om.cons(x)[0] - (np.dot(w, x[1:] ** 2) + x[0])) < 1e-14
    assert om.lcon[0] == -np.inf and om.ucon[0] == 3.0
    r, c = om.jac_structure()
    J = np.zeros(6); np.add.at(J, c - 1, om.jac_coord(x))
    assert np.allclose(J, np.r_[1.0, 2 * w * x[1:]], atol=1e-14)


@pytest.mark.parametrize("method,degree", [(io.FiniteDifference("backward"), 1), (io.FiniteDifference("forward"), 1),
                                           (io.FiniteDifference("central"), 1), (io.OrthogonalCollocation(3), 2),
                                           (io.OrthogonalCollocation(4), 3), (io.OrthogonalCollocation(5), 4)])
def test_derivative_approximation_rows_are_exact_for_polynomials(method, degree):
    """The derivative-approximation generators (transform.jl:511-562 with InfiniteOpt's derivative_expr_data: finite
    differences, orthogonal collocation with num_nodes Lobatto nodes per interval) vanish when y is a polynomial of the
    degree the scheme integrates exactly (on NON-uniform public supports: 1 for the finite differences, num_nodes - 1 for
    collocation) and the derivative variable holds y' — evaluated
    through the oracle at x = (y(t_i), y'(t_i))."""
    from oracle.oracle import OracleModel
    m = io.InfiniteModel()
    pub = np.array([0.0, 0.3, 0.35, 1.1, 2.0, 2.05, 3.0])
    t = m.infinite_parameter(supports=pub, derivative_method=method)
    y = m.variable(t)
    m.constraint(m.deriv(y, t), "==", 1.0)           # any constraint that brings the derivative variable in
    m.objective("Min", m.integral(y ** 2, t))
    core, data = exa_core(m)
    ts = data.supports[0]
    T = len(ts)
    coef = np.array([0.7, -1.3, 0.5, 0.25, -0.125])[:degree + 1]
    p = np.polynomial.Polynomial(coef)
    x = np.concatenate([p(ts), p.deriv()(ts)])       # layout: y block, then the derivative variable (transform.jl:141-144)
    assert core.nvar == 2 * T
    om = OracleModel(core)
    c = om.cons(x)
    rows = c[T:]                                      # the approximation rows follow the T rows of the constraint
    assert len(rows) >= T - 2 and np.max(np.abs(rows)) <= 1e-12, np.max(np.abs(rows))
    if isinstance(method, io.OrthogonalCollocation):
        assert T == len(pub) + (method.num_nodes - 2) * (len(pub) - 1)      # internal nodes (transform.jl:22)


def test_error_paths_read_like_the_reference():
    """integer variables (transform.jl:41-45) and unsupported constraint sets (transform.jl:408-411)"""
    m = io.InfiniteModel()
    t = m.infinite_parameter(0, 1, num_supports=3)
    m.variable(t, integer=True)
    with pytest.raises(ValueError, match="Integer variables are not supported by ExaModels."):
        exa_core(m)
    m2 = io.InfiniteModel()
    z = m2.variable(binary=True)
    with pytest.raises(ValueError, match="Integer variables are not supported by ExaModels."):
        exa_core(m2)
    m3 = io.InfiniteModel()
    w = m3.variable()
    with pytest.raises(ValueError, match="is not supported by InfiniteExaModels"):
        m3.constraint(w, "in SecondOrderCone", 1.0)
