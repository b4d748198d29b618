"""Compares the engine with dumps of the REAL ExaModels evaluator (julia/dump_golden.jl).  No dump can
be produced in this container (no Julia); the test is skipped until one is dropped into
tests/golden/, which is what converts "parity unpinned" into pinned parity."""
import glob
import os

import numpy as np
import pytest

import iexa_b200 as ex
from iexa_b200 import models
from conftest import ROOT, assert_close

def _pandemic():
    xi_file = os.path.join(ROOT, "tests", "golden", "pandemic_50x4.xi")   # scenario supports the Julia run drew
    xi = np.fromfile(xi_file, dtype="<f8") if os.path.exists(xi_file) else None
    return models.pandemic(50, 4, xi=xi)


BUILDERS = {"ode_5x5": lambda: models.ode_5x5(), "quadrotor_oc_40": lambda: models.quadrotor(40, "oc"),
            "pandemic_50x4": _pandemic}
DUMPS = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.golden")))


def load(path):
    raw = np.fromfile(path, dtype=np.uint8)
    pos = 0

    def take(dt, n):
        nonlocal pos
        a = np.frombuffer(raw, dtype=dt, count=n, offset=pos); pos += a.nbytes; return a

    nvar, ncon, nnzj, nnzh = take("<i8", 4)
    d = dict(x=take("<f8", nvar), y=take("<f8", ncon), sigma=float(take("<f8", 1)[0]), obj=float(take("<f8", 1)[0]),
             grad=take("<f8", nvar), cons=take("<f8", ncon), jr=take("<i8", nnzj), jc=take("<i8", nnzj), jv=take("<f8", nnzj),
             hr=take("<i8", nnzh), hc=take("<i8", nnzh), hv=take("<f8", nnzh))
    return d


@pytest.mark.skipif(not DUMPS, reason="no ExaModels dump present (needs Julia; see julia/dump_golden.jl)")
@pytest.mark.parametrize("path", DUMPS)
def test_against_exa_models_dump(path, hostcheck_lib):
    compare_with_dump(path, hostcheck_lib)


def test_dump_harness_on_a_dump_written_by_the_oracle(tmp_path, hostcheck_lib):
    """The comparison above has never seen a real dump (no Julia here).  This runs the SAME loader and comparison on a
    file in julia/dump_golden.jl's binary format whose contents come from the oracle, so that the day a real dump is
    dropped into tests/golden/ the harness itself is known to work — it pins nothing about ExaModels."""
    from oracle.oracle import OracleModel
    from conftest import eval_point
    core = BUILDERS["quadrotor_oc_40"]()
    om = OracleModel(core)
    x, y = eval_point(core, seed=0)
    jr, jc = om.jac_structure(); hr, hc = om.hess_structure()
    path = tmp_path / "quadrotor_oc_40.golden"
    with open(path, "wb") as f:
        for a in (np.array([om.nvar, om.ncon, om.nnzj, om.nnzh], dtype="<i8"), x, y, np.array([0.7]), np.array([om.obj(x)]),
                  om.grad(x), om.cons(x), jr.astype("<i8"), jc.astype("<i8"), om.jac_coord(x), hr.astype("<i8"),
                  hc.astype("<i8"), om.hess_coord(x, y, 0.7)):
            f.write(np.ascontiguousarray(a).tobytes())
    compare_with_dump(str(path), hostcheck_lib)


def compare_with_dump(path, hostcheck_lib):
    name = os.path.splitext(os.path.basename(path))[0]
    d = load(path)
    L = hostcheck_lib
    m = ex.ExaModel(BUILDERS[name](), flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    assert (m.meta.nvar, m.meta.ncon, m.meta.nnzj, m.meta.nnzh) == (len(d["x"]), len(d["y"]), len(d["jv"]), len(d["hv"]))
    r = np.zeros(len(d["jv"]), dtype=np.int64); c = np.zeros_like(r)
    L.hostcheck_structure(m.h, 0, r.ctypes.data, c.ctypes.data)
    assert (r == d["jr"]).all() and (c == d["jc"]).all(), "Jacobian structure must be bit-exact"
    r = np.zeros(len(d["hv"]), dtype=np.int64); c = np.zeros_like(r)
    L.hostcheck_structure(m.h, 1, r.ctypes.data, c.ctypes.data)
    assert (r == d["hr"]).all() and (c == d["hc"]).all(), "Hessian structure must be bit-exact"
    x, y = np.ascontiguousarray(d["x"]), np.ascontiguousarray(d["y"])
    for which, key, n in ((2, "cons", len(y)), (3, "jv", len(d["jv"])), (4, "hv", len(d["hv"])), (1, "grad", len(x))):
        out = np.zeros(max(n, 1))
        L.hostcheck_eval_groups(m.h, which, x.ctypes.data, y.ctypes.data, d["sigma"], out.ctypes.data, None)
        assert_close(out[:n], d[key], key)
