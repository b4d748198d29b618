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

BUILDERS = {"ode_5x5": lambda: models.ode_5x5(), "quadrotor_oc_40": lambda: models.quadrotor(40, "oc"),
            "pandemic_50x4": lambda: models.pandemic(50, 4)}
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
