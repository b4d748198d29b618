"""The oracle against central finite differences on every config model (the oracle is "parity
unpinned" against ExaModels itself, so it is pinned against calculus here and against the
reference's solve-level goldens in test_golden_solves.py)."""
import numpy as np
import pytest

from iexa_b200 import models
from conftest import eval_point

CASES = {
    "ode_5x5": lambda: models.ode_5x5(),
    "quadrotor_oc": lambda: models.quadrotor(6, "oc"),
    "quadrotor_fd": lambda: models.quadrotor(7, "fd"),
    "pandemic": lambda: models.pandemic(5, 2),
    "farmer": lambda: models.farmer(5),
    "rosenbrock_param": lambda: models.rosenbrock_param()[0],
    "param_function": lambda: models.param_function_model()[0],
}


@pytest.mark.parametrize("name", list(CASES))
def test_first_and_second_derivatives(name):
    from oracle.oracle import OracleModel
    core = CASES[name]()
    om = OracleModel(core)
    x, y = eval_point(core, seed=11)
    x = x + 0.05  # stay away from the bounds' kinks
    n, eps = om.nvar, 1e-6
    E = np.eye(n)
    g = om.grad(x)
    gfd = np.array([(om.obj(x + eps * e) - om.obj(x - eps * e)) / (2 * eps) for e in E])
    assert np.allclose(g, gfd, rtol=1e-6, atol=1e-6)
    r, c = om.jac_structure()
    J = np.zeros((om.ncon, n)); np.add.at(J, (r - 1, c - 1), om.jac_coord(x))
    Jfd = np.array([(om.cons(x + eps * e) - om.cons(x - eps * e)) / (2 * eps) for e in E]).T
    assert np.allclose(J, Jfd, rtol=1e-6, atol=1e-6)
    hr, hc = om.hess_structure()
    assert (hr >= hc).all()
    H = np.zeros((n, n)); np.add.at(H, (hr - 1, hc - 1), om.hess_coord(x, y, 0.7)); H = H + np.tril(H, -1).T
    lg = lambda z: 0.7 * om.grad(z) + om.jtprod(z, y)
    Hfd = np.array([(lg(x + eps * e) - lg(x - eps * e)) / (2 * eps) for e in E]).T
    assert np.allclose(H, Hfd, rtol=1e-5, atol=1e-5)
    v = np.linspace(-1, 1, n)
    assert np.allclose(om.jprod(x, v), J @ v, atol=1e-12)
    assert np.allclose(om.jtprod(x, y), J.T @ y, atol=1e-12)
    assert np.allclose(om.hprod(x, y, v, 0.7), H @ v, atol=1e-10)
