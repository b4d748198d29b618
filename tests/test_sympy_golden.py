"""Known-answer vectors from an INDEPENDENT symbolic restatement (tests/golden/make_sympy_golden.py: sympy,
40-digit arithmetic, written from the reference's model files, imports nothing from this repo) against

  * the oracle (oracle/oracle.c, the CPU restatement of ExaModels' algorithm),
  * the plan compiler's register programs run on the host (tests/hostcheck), and
  * the CUDA engine through the C ABI (`-m gpu`).

The fixtures pin values, not slot order: COO entries are summed into (row, col) form before comparing.
Tolerance is the north star's: 1e-12 relative / 1e-14 absolute — relative to the largest TERM of an entry
that is a sum of duplicates (cancellation inside a sum is not an evaluator error)."""
import os

import numpy as np
import pytest

import iexa_b200 as ex
from iexa_b200 import models
from iexa_b200.expr import nl_op
from conftest import ROOT, has_gpu

G = os.path.join(ROOT, "tests", "golden")


def operators_core():
    names = open(os.path.join(G, "sympy_operators.names")).read().split()
    core = ex.ExaCore()
    rows = []
    binary = {
        "add": lambda a, b: (a + b) * (a + 0.75),
        "sub": lambda a, b: (a - b) * (0.75 - a) * (b - 2.0),
        "mul": lambda a, b: a * b * a * 0.75,
        "div": lambda a, b: a / b + 0.75 / a + b / 0.75,
        "pow_var_const": lambda a, b: a ** 3.0 + b ** 2 + a ** 0.75,
        "pow_const_var": lambda a, b: 2.0 ** a + 0.75 ** b,
        "pow_var_var": lambda a, b: a ** b,
    }
    for nm in names:
        a, b = core.add_var(1)[1], core.add_var(1)[1]
        if nm in binary:
            rows.append(binary[nm](a, b))
        else:
            f = nl_op(nm, compat=(nm != "csch"))     # the mathematical csch (the reference's :csch => csc quirk is tested elsewhere)
            rows.append(f(a * 1.0) * b + f(0.75 * a))
    for r in rows:
        core.add_con(r, ex.Itr.empty())
    for r in rows:
        core.add_obj(r, ex.Itr.empty())
    return core


BUILDERS = {
    "ode_5x5": lambda: models.ode_5x5(),
    "quadrotor_fd": lambda: models.quadrotor(5, "fd"),
    "quadrotor_oc": lambda: models.quadrotor(4, "oc"),
    "pandemic": lambda: models.pandemic(4, 2, seed=0),
    "farmer": lambda: models.farmer(4, seed=42),
    "operators": operators_core,
}


def _solve_problem(variant):
    from iexa_b200 import infmodels
    from iexa_b200.transform import exa_core
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # the slow-path measure expansion warns like the reference (transform.jl:683,716)
        return exa_core(infmodels.solve_test_problem(variant))[0]


def _param_function_problem():
    from iexa_b200 import infmodels
    from iexa_b200.transform import exa_core
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return exa_core(infmodels.param_function_problem())[0]


def _opf_case3():
    from iexa_b200 import opf
    from iexa_b200.transform import exa_core
    return exa_core(opf.opf(None, num_supports=2, seed=0))[0]


BUILDERS["opf_case3"] = _opf_case3
BUILDERS["param_function_problem"] = _param_function_problem
BUILDERS["solve_tp1"] = lambda: _solve_problem(-1)
for _v in range(5):
    BUILDERS[f"solve_tp2_v{_v}"] = (lambda v: (lambda: _solve_problem(v)))(_v)
ROW_MATCHED = {n for n in BUILDERS if n.startswith("solve_")} | {"param_function_problem", "opf_case3"}   # fixture rows are in the generator script's order


def load(name):
    return dict(np.load(os.path.join(G, f"sympy_{name}.npz")))


def dense(n0, n1, r, c, v):
    """sum of duplicates and the largest |term| per entry"""
    A = np.zeros((n0, n1)); M = np.zeros((n0, n1))
    np.add.at(A, (r, c), v)
    np.maximum.at(M, (r, c), np.abs(v))
    return A, M


def close(a, b, scale, what):
    a, b, scale = np.asarray(a, float), np.asarray(b, float), np.asarray(scale, float)
    tol = 1e-14 + 1e-12 * np.maximum(np.maximum(np.abs(a), np.abs(b)), scale)
    bad = ~(np.abs(a - b) <= tol)
    assert not bad.any(), f"{what}: {int(bad.sum())} entries differ, worst {np.abs(a - b)[bad].max():.3e}"


def match_rows(cons, gold):
    """row permutation between the engine's constraint order and the fixture's, recovered from the (distinct) values"""
    a, b = np.argsort(cons), np.argsort(gold)
    assert np.all(np.abs(np.sort(cons) - np.sort(gold)) <= 1e-14 + 1e-12 * np.abs(np.sort(gold))), "constraint VALUES differ as multisets"
    # rows with EQUAL values (e.g. z(0, 2.5) + pf2*pf at t = 0, where pf = sin(0) = 0, for every s) are matched in sorted
    # order; a wrong tie-break would show up in the Jacobian / Hessian comparison that follows
    perm = np.empty(len(cons), dtype=np.int64)
    perm[a] = b          # engine row i  <->  fixture row perm[i]
    return perm


def check(d, nvar, ncon, obj, grad, cons, jr, jc, jv, hr, hc, hv, obj_scale=None, perm=None):
    assert (nvar, ncon) == (int(d["nvar"]), int(d["ncon"]))
    if perm is not None:   # bring the fixture's rows into the engine's order
        inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
        d = dict(d); d["cons"] = d["cons"][perm]; d["jr"] = inv[d["jr"]]
    close(obj, d["obj"], np.abs(d["cons"]).sum() if obj_scale is None else obj_scale, "obj")
    close(grad, d["grad"], 0.0, "grad")
    close(cons, d["cons"], np.abs(d["x"]).max(), "cons")
    J, JM = dense(ncon, nvar, jr - 1, jc - 1, jv)
    Jg, _ = dense(ncon, nvar, d["jr"], d["jc"], d["jv"])
    close(J, Jg, JM, "jacobian")
    assert (hr >= hc).all(), "Hessian is not lower-triangular"
    H, HM = dense(nvar, nvar, hr - 1, hc - 1, hv)
    Hg, _ = dense(nvar, nvar, d["hr"], d["hc"], d["hv"])
    close(H, Hg, HM, "hessian of the Lagrangian")


def rows_and_multipliers(name, d, cons):
    """(multipliers in the engine's row order, row permutation or None)"""
    if name not in ROW_MATCHED:
        return np.ascontiguousarray(d["y"]), None
    perm = match_rows(np.asarray(cons), d["cons"])
    return np.ascontiguousarray(d["y"][perm]), perm


@pytest.mark.parametrize("name", list(BUILDERS))
def test_oracle_against_sympy(name):
    from oracle.oracle import OracleModel
    d = load(name)
    core = BUILDERS[name]()
    om = OracleModel(core)
    x, s = d["x"], float(d["sigma"])
    y, perm = rows_and_multipliers(name, d, om.cons(x))
    jr, jc = om.jac_structure()
    hr, hc = om.hess_structure()
    check(d, om.nvar, om.ncon, om.obj(x), om.grad(x), om.cons(x), jr, jc, om.jac_coord(x), hr, hc, om.hess_coord(x, y, s), perm=perm)


@pytest.mark.parametrize("name", list(BUILDERS))
def test_compiled_programs_against_sympy(name, hostcheck_lib):
    L = hostcheck_lib
    d = load(name)
    core = BUILDERS[name]()
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    x, s = np.ascontiguousarray(d["x"]), float(d["sigma"])

    def hc_(which, n, yy=None, sg=1.0):
        out = np.zeros(max(n, 1))
        assert L.hostcheck_eval_groups(m.h, which, x.ctypes.data, None if yy is None else yy.ctypes.data, sg, out.ctypes.data, None) == 0
        return out[:n]

    from oracle.oracle import OracleModel
    om = OracleModel(core)            # structure only (bit-exactness of the engine's structure vs the oracle is tested elsewhere)
    jr, jc = om.jac_structure()
    hr, hcol = om.hess_structure()
    obj = hc_(0, 1)[0]
    y, perm = rows_and_multipliers(name, d, hc_(2, m.meta.ncon))
    check(d, m.meta.nvar, m.meta.ncon, obj, hc_(1, m.meta.nvar), hc_(2, m.meta.ncon), jr, jc, hc_(3, m.meta.nnzj),
          hr, hcol, hc_(4, m.meta.nnzh, y, s), perm=perm)


@pytest.mark.gpu
@pytest.mark.skipif(not has_gpu(), reason="needs a CUDA device")
@pytest.mark.parametrize("interp", [False, True])
@pytest.mark.parametrize("name", list(BUILDERS))
def test_cuda_engine_against_sympy(name, interp):
    import torch
    d = load(name)
    core = BUILDERS[name]()
    m = ex.ExaModel(core, device=0, flags=ex.lib.IEXA_F_NO_SPECIALISE if interp else ex.lib.IEXA_F_DEFAULT)
    if not interp:
        assert m.cmeta.n_kernels_specialised > 0, m.L.iexa_engine_note(m.h).decode()
    dev = "cuda"
    x, s = torch.from_numpy(d["x"]).to(dev), float(d["sigma"])
    n, mc = m.meta.nvar, m.meta.ncon
    c = torch.zeros(max(mc, 1), dtype=torch.float64, device=dev)
    g = torch.zeros(n, dtype=torch.float64, device=dev)
    jv = torch.zeros(max(m.meta.nnzj, 1), dtype=torch.float64, device=dev)
    hv = torch.zeros(max(m.meta.nnzh, 1), dtype=torch.float64, device=dev)
    jr = torch.zeros(max(m.meta.nnzj, 1), dtype=torch.int64, device=dev); jc = torch.zeros_like(jr)
    hr = torch.zeros(max(m.meta.nnzh, 1), dtype=torch.int64, device=dev); hc = torch.zeros_like(hr)
    ex.cons_(m, x, c)
    yh, perm = rows_and_multipliers(name, d, c.cpu().numpy()[:mc])
    y = torch.from_numpy(yh).to(dev)
    ex.grad_(m, x, g); ex.jac_coord_(m, x, jv); ex.hess_coord_(m, x, y, hv, s)
    ex.jac_structure_(m, jr, jc); ex.hess_structure_(m, hr, hc)
    f = ex.obj(m, x)
    cpu = lambda t, k: t.cpu().numpy()[:k]
    check(d, n, mc, f, cpu(g, n), cpu(c, mc), cpu(jr, m.meta.nnzj), cpu(jc, m.meta.nnzj), cpu(jv, m.meta.nnzj),
          cpu(hr, m.meta.nnzh), cpu(hc, m.meta.nnzh), cpu(hv, m.meta.nnzh), perm=perm)
