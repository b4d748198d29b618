"""Every operator of the reference's table (src/operators.jl:2-46) to second order: value, first
and second derivatives from the engine's programs against the oracle AND against central finite
differences of the oracle's value (an independent check of both derivative tables)."""
import numpy as np
import pytest

import iexa_b200 as ex
from iexa_b200.expr import _OP_MAPPINGS, nl_op
from conftest import assert_close

# evaluation intervals inside every operator's domain
DOMAIN = {"sqrt": (0.5, 2), "log": (0.5, 2), "log2": (0.5, 2), "log10": (0.5, 2), "log1p": (0.2, 2), "cbrt": (0.5, 2),
          "asin": (-0.7, 0.7), "acos": (-0.7, 0.7), "atanh": (-0.7, 0.7), "acoth": (1.3, 3), "inv": (0.5, 2),
          "csc": (0.4, 1.2), "cot": (0.4, 1.2), "csch": (0.4, 1.2), "coth": (0.4, 1.5), "cscd": (20, 70), "cotd": (20, 70),
          "secd": (-50, 50), "tand": (-50, 50), "sind": (-80, 80), "cosd": (-80, 80), "acot": (0.3, 2), "acotd": (0.3, 2),
          "tan": (-1, 1), "sec": (-1, 1), "abs": (0.2, 1.5)}
UNARY = [s for s in _OP_MAPPINGS if s not in "+-*/^"]


def _model(build, K=7, lo=-1.0, hi=1.0, seed=0):
    rng = np.random.default_rng(seed)
    core = ex.ExaCore()
    a = core.add_var(K); b = core.add_var(K)
    it = ex.Itr(K, {"i": np.arange(1, K + 1)}, {"p": rng.uniform(0.5, 1.5, K)})
    ds = ex.DataSource()
    core.add_con(build(a[ds.i], b[ds.i], ds.p), it)
    core.add_obj(build(a[ds.i], b[ds.i], ds.p), it)
    x = rng.uniform(lo, hi, 2 * K)
    return core, x


def _check(core, x, L):
    from oracle.oracle import OracleModel
    om = OracleModel(core)
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    assert (m.meta.nnzj, m.meta.nnzh) == (om.nnzj, om.nnzh)
    y = np.linspace(0.5, 1.5, om.ncon)

    def hc(which, n, yy=None, s=1.0):
        out = np.zeros(max(n, 1))
        assert L.hostcheck_eval_groups(m.h, which, x.ctypes.data, None if yy is None else yy.ctypes.data, s, out.ctypes.data, None) == 0
        return out[:n]

    assert_close(hc(2, om.ncon), om.cons(x), "cons")
    assert_close(hc(3, om.nnzj), om.jac_coord(x), "jac")
    assert_close(hc(4, om.nnzh, y, 0.3), om.hess_coord(x, y, 0.3), "hess")
    assert_close(hc(1, om.nvar), om.grad(x), "grad")
    # finite differences of the oracle value: independent of both derivative tables
    eps = 1e-6
    g = om.grad(x)
    gfd = np.array([(om.obj(x + eps * e) - om.obj(x - eps * e)) / (2 * eps) for e in np.eye(om.nvar)])
    assert np.allclose(g, gfd, rtol=2e-6, atol=2e-7), np.abs(g - gfd).max()
    r, c = om.hess_structure()
    H = np.zeros((om.nvar, om.nvar)); np.add.at(H, (r - 1, c - 1), om.hess_coord(x, None, 1.0)); H = H + np.tril(H, -1).T
    Hfd = np.array([(om.grad(x + eps * e) - om.grad(x - eps * e)) / (2 * eps) for e in np.eye(om.nvar)])
    assert np.allclose(H, Hfd, rtol=2e-5, atol=2e-6), np.abs(H - Hfd).max()


@pytest.mark.parametrize("sym", UNARY)
def test_unary_operator(sym, hostcheck_lib):
    lo, hi = DOMAIN.get(sym, (-1.0, 1.0))
    f = nl_op(sym, compat=(sym != "csch"))  # the TRUE csch here; the compat mapping is tested below
    core, x = _model(lambda a, b, p: f(a * 1.0) * b + f(p * a), lo=lo, hi=hi)
    if sym in ("abs",):
        x = np.abs(x) + 0.1
    _check(core, x, hostcheck_lib)


@pytest.mark.parametrize("name,build,lo,hi", [
    ("add", lambda a, b, p: (a + b) * (a + p), -1, 1),
    ("sub", lambda a, b, p: (a - b) * (p - a) * (b - 2.0), -1, 1),
    ("mul", lambda a, b, p: a * b * a * p, -1, 1),
    ("div", lambda a, b, p: a / b + p / a + b / p, 0.5, 2),
    ("pow_var_const", lambda a, b, p: a ** 3.0 + b ** 2 + a ** p, 0.5, 2),
    ("pow_const_var", lambda a, b, p: 2.0 ** a + p ** b, -1, 1),
    ("pow_var_var", lambda a, b, p: a ** b, 0.5, 2),
    ("neg_pos", lambda a, b, p: -(a * b) + (+(a * a)), -1, 1),
    ("nested", lambda a, b, p: ex.sin(a * b) / ex.cos(b) + ex.exp(a - b) * ex.tan(a) - ex.sqrt(a * a + 1.0), -0.8, 0.8),
])
def test_binary_and_nested(name, build, lo, hi, hostcheck_lib):
    core, x = _model(build, lo=lo, hi=hi)
    _check(core, x, hostcheck_lib)


def test_reference_csch_quirk_is_reproducible():
    """src/operators.jl:41 maps :csch to csc; nl_op reproduces that by default, and offers the true
    csch behind compat=False."""
    a = ex.Var(ex.IndexExpr(1))
    assert nl_op("csch")(a).op == ex.OP["CSC"]
    assert nl_op("csch", compat=False)(a).op == ex.OP["CSCH"]
    with pytest.raises(ValueError, match="does not support the nonlinear operator"):
        nl_op("erf")  # operators.jl:50-53


def test_nary_fold_is_left_deep():
    a, b, c = (ex.Var(ex.IndexExpr(i)) for i in (1, 2, 3))
    e = nl_op("*")(a, b, c)
    assert isinstance(e.a, ex.expr.Binary) and e.a.a is a and e.a.b is b and e.b is c
