"""The COO slot-ORDER policy, written out by hand for small expressions.

SURVEY.md Appendix A is a from-memory restatement of how ExaModels orders its sparse entries (A.2: first-order slots are
the distinct Var leaves in traversal order; A.4: `hrpass0` emits nothing through top-level `+ - const*` chains, below the
first nonlinear node `hrpass` emits a diagonal slot per leaf — even when its value is identically zero — and a binary node
walks child 1, then child 2, then the cross pairs of the two subtrees; identical index pairs share a slot; lower triangle).
None of it can be checked against the real ExaModels here (parity unpinned).  This file states what that
hypothesis MEANS on concrete expressions, derived by hand from the rules, and checks that the oracle AND the plan compiler
implement exactly that — so that the day a `*.golden` dump disagrees, the rule to change is identified by the case that
breaks, and a change of policy cannot slip in unnoticed."""
import numpy as np
import pytest

import iexa_b200 as ex

X = np.array([0.3, 0.7, -0.4, 1.1])

CASES = {
    # x1*x2 + sin(x1): `+` at the top emits nothing; x1*x2 -> diag(x1) [value 0], diag(x2) [value 0], cross (x2, x1);
    # sin(x1) -> diag(x1), the same index pair as the first slot, so it shares it: value -sin(x1)
    "x1*x2 + sin(x1)": (lambda a, b, c, d: a * b + ex.sin(a), [1, 2], [(1, 1), (2, 2), (2, 1)],
                        lambda x: [-np.sin(x[0]), 0.0, 1.0]),
    # ((u*cos(x7))*sin(x8)): child 1 = u*cos(x7) -> diag(u), diag(x7), cross (x7, u); child 2 = sin(x8) -> diag(x8);
    # cross pass of the outer product over (leaves of child 1) x (leaves of child 2): (x8, u), (x8, x7)
    "u*cos(x7)*sin(x8)": (lambda a, b, c, d: a * ex.cos(b) * ex.sin(c), [1, 2, 3],
                          [(1, 1), (2, 2), (2, 1), (3, 3), (3, 1), (3, 2)],
                          lambda x: [0.0, -x[0] * np.cos(x[1]) * np.sin(x[2]), -np.sin(x[1]) * np.sin(x[2]),
                                     -x[0] * np.cos(x[1]) * np.sin(x[2]), np.cos(x[1]) * np.cos(x[2]),
                                     -x[0] * np.sin(x[1]) * np.cos(x[2])]),
    # affine: Jacobian slots in traversal order, no Hessian slots at all
    "2*x1 - x2 + 3": (lambda a, b, c, d: 2.0 * a - b + 3.0, [1, 2], [], lambda x: []),
    # the shape of the quadrotor rows  dx - (u*cos(x) - 9.8): the linear leaf x1 gets NO slot, the product does
    "x1 - (x2*cos(x3) - 9.8)": (lambda a, b, c, d: a - (b * ex.cos(c) - 9.8), [1, 2, 3], [(2, 2), (3, 3), (3, 2)],
                                lambda x: [0.0, x[1] * np.cos(x[2]), np.sin(x[2])]),   # minus (minus x2*cos x3), minus (minus sin x3)
    # JuMP quadratic form  abs2(x1) + (-2)*x1*x4  (transform.jl:365-381)
    "abs2(x1) + (-2)*x1*x4": (lambda a, b, c, d: ex.abs2(a) + (-2.0) * a * d, [1, 4], [(1, 1), (4, 4), (4, 1)],
                              lambda x: [2.0, 0.0, -2.0]),
    "x1/x2": (lambda a, b, c, d: a / b, [1, 2], [(1, 1), (2, 2), (2, 1)],
              lambda x: [0.0, 2 * x[0] / x[1] ** 3, -1 / x[1] ** 2]),
}


def _core(build):
    core = ex.ExaCore()
    v = [core.add_var(1)[1] for _ in range(4)]
    core.add_con(build(*v), ex.Itr.empty())
    return core


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_orders_slots_as_the_hypothesis_says(name):
    from oracle.oracle import OracleModel
    build, jcols, hpairs, hvals = CASES[name]
    om = OracleModel(_core(build))
    jr, jc = om.jac_structure()
    hr, hc = om.hess_structure()
    assert jc.tolist() == jcols and jr.tolist() == [1] * len(jcols)
    assert list(zip(hr.tolist(), hc.tolist())) == hpairs
    got = om.hess_coord(X, np.array([1.0]), 0.0)
    assert np.allclose(got, hvals(X), rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name", list(CASES))
def test_plan_compiler_orders_slots_as_the_hypothesis_says(name, hostcheck_lib):
    L = hostcheck_lib
    build, jcols, hpairs, hvals = CASES[name]
    m = ex.ExaModel(_core(build), flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    assert (m.meta.nnzj, m.meta.nnzh) == (len(jcols), len(hpairs))
    r = np.zeros(max(len(jcols), 1), dtype=np.int64); c = np.zeros_like(r)
    L.hostcheck_structure(m.h, 0, r.ctypes.data, c.ctypes.data)
    assert c[:len(jcols)].tolist() == jcols
    r = np.zeros(max(len(hpairs), 1), dtype=np.int64); c = np.zeros_like(r)
    L.hostcheck_structure(m.h, 1, r.ctypes.data, c.ctypes.data)
    assert list(zip(r[:len(hpairs)].tolist(), c[:len(hpairs)].tolist())) == hpairs
    out = np.zeros(max(len(hpairs), 1))
    y = np.array([1.0])
    assert L.hostcheck_eval_groups(m.h, 4, X.ctypes.data, y.ctypes.data, 0.0, out.ctypes.data, None) == 0
    assert np.allclose(out[:len(hpairs)], hvals(X), rtol=1e-13, atol=1e-15)
