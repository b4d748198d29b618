"""NaN / Inf semantics.  The reference's evaluator is a run-time AD: it multiplies numbers, so a structural zero times an
infinite or NaN partial is NaN.  The engine differentiates symbolically; by DEFAULT (IEXA_OPT_STRICT_IEEE = 1) it keeps
structural zeros as run-time multiplications, and the NaN PATTERN of every callback equals the oracle's on poisoned inputs —
for every BASELINE configuration.  IEXA_OPT_STRICT_IEEE = 0 folds 0*x -> 0: identical results for finite inputs, a SUBSET
of the NaNs otherwise (1.5-2.5 % faster on B200).
abs'(0) = +1 (sign(+0) taken as +1) in oracle and engine alike (src/operators.jl:14)."""
import ctypes as C

import numpy as np
import pytest

import iexa_b200 as ex
from iexa_b200 import models
from iexa_b200.core import ExaCore, Itr
from iexa_b200.expr import DataSource
from conftest import eval_point


def _opf():
    from iexa_b200 import opf
    from iexa_b200.transform import exa_core
    return exa_core(opf.opf(None, num_supports=5))[0]


CASES = {"config1_quadrotor_fd": lambda: models.quadrotor(12, "fd"), "config3_quadrotor_oc": lambda: models.quadrotor(9, "oc"),
         "config2_pandemic": lambda: models.pandemic(7, 3), "config4_opf": _opf, "config5_farmer": lambda: models.farmer(11),
         "ode_5x5": lambda: models.ode_5x5()}


def _hc(L, m, which, n, x, y=None, sig=1.0, fn="hostcheck_eval_groups"):
    out = np.zeros(max(n, 1))
    args = [m.h, which, x.ctypes.data, None if y is None else y.ctypes.data, sig, out.ctypes.data]
    if fn == "hostcheck_eval_groups":
        args.append(C.byref(C.c_int32()))
    assert getattr(L, fn)(*args) == 0
    return out[:n]


def _compare(ref, got):
    nr, ng = np.isnan(ref), np.isnan(got)
    fin = ~nr & ~ng
    with np.errstate(invalid="ignore"):
        same = (ref[fin] == got[fin]) | (np.abs(ref[fin] - got[fin]) <= 1e-14 + 1e-12 * np.abs(ref[fin]))
    return int((nr != ng).sum()), int((ng & ~nr).sum()), bool(same.all())


@pytest.mark.parametrize("poison", [np.nan, np.inf, -np.inf])
@pytest.mark.parametrize("name", list(CASES))
def test_strict_mode_reproduces_the_oracles_nan_pattern(name, poison, hostcheck_lib):
    from oracle.oracle import OracleModel
    L = hostcheck_lib
    core = CASES[name]()
    om = OracleModel(core)
    strict = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)   # the default IS strict
    folded = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L, strict_ieee=False)
    x, y = eval_point(core)
    x = np.where(np.isfinite(x), x, 0.0)
    rng = np.random.default_rng(1)
    xp = x.copy()
    xp[rng.choice(core.nvar, size=max(1, core.nvar // 10), replace=False)] = poison
    for which, n, ref in ((2, om.ncon, om.cons(xp)), (3, om.nnzj, om.jac_coord(xp)), (4, om.nnzh, om.hess_coord(xp, y, 0.7)),
                          (1, om.nvar, om.grad(xp))):
        sig = 0.7 if which == 4 else 1.0
        for fn in ("hostcheck_eval", "hostcheck_eval_groups"):
            mism, extra, fin_ok = _compare(ref, _hc(L, strict, which, n, xp, y, sig, fn))
            assert mism == 0 and fin_ok, (name, which, fn, "strict", mism)
        # default (folded) mode: never a NaN the reference does not have, identical where both are finite
        mism, extra, fin_ok = _compare(ref, _hc(L, folded, which, n, xp, y, sig))
        assert extra == 0 and fin_ok, (name, which, "folded", extra)


def test_abs_derivative_at_zero_is_plus_one(hostcheck_lib):
    """d|u|/du at u = +0.0 is +1 in oracle and engine (the convention is pinned here; upstream's is unverified)"""
    from oracle.oracle import OracleModel
    from iexa_b200.expr import nl_op
    L = hostcheck_lib
    core = ExaCore(minimize=True)
    ds = DataSource()
    v = core.add_var(3, start=0.0)
    it = Itr(3, {"group_idx1": np.arange(1, 4)}, {})
    core.add_con(nl_op("abs")(v[ds.group_idx1]), it, 0.0, 1.0)
    core.add_obj(nl_op("abs")(v[ds.group_idx1]), it)
    om = OracleModel(core)
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    x = np.array([0.0, -0.0, -2.0])
    ref = om.jac_coord(x)
    assert list(ref) == [1.0, 1.0, -1.0]            # -0.0 >= 0 is true: +1
    assert list(_hc(L, m, 3, om.nnzj, x)) == list(ref)
    assert list(_hc(L, m, 1, om.nvar, x)) == list(om.grad(x))
