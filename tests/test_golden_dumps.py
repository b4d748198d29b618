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


@pytest.mark.parametrize("policy", [0, 1])
def test_dump_harness_on_a_dump_written_by_the_oracle(tmp_path, hostcheck_lib, policy):
    """The comparison above has never seen a real dump (no Julia here).  This runs the SAME loader and comparison on a
    file in julia/dump_golden.jl's binary format whose contents come from the oracle — under either slot-order policy: the
    harness must IDENTIFY the policy from the structure and then hold the values — so that the day a real dump is
    dropped into tests/golden/ the harness itself is known to work.  It pins nothing about ExaModels."""
    from oracle.oracle import OracleModel
    from conftest import eval_point
    core = BUILDERS["quadrotor_oc_40"]()
    om = OracleModel(core, slot_order=policy)
    x, y = eval_point(core, seed=0)
    jr, jc = om.jac_structure(); hr, hc = om.hess_structure()
    path = tmp_path / "quadrotor_oc_40.golden"
    with open(path, "wb") as f:
        for a in (np.array([om.nvar, om.ncon, om.nnzj, om.nnzh], dtype="<i8"), x, y, np.array([0.7]), np.array([om.obj(x)]),
                  om.grad(x), om.cons(x), jr.astype("<i8"), jc.astype("<i8"), om.jac_coord(x), hr.astype("<i8"),
                  hc.astype("<i8"), om.hess_coord(x, y, 0.7)):
            f.write(np.ascontiguousarray(a).tobytes())
    compare_with_dump(str(path), hostcheck_lib)


POLICIES = {0: "IEXA_SLOT_ORDER_LEFT_TO_RIGHT", 1: "IEXA_SLOT_ORDER_RIGHT_TO_LEFT"}


def identify_policy(name, d, L):
    """the slot-order policy (iexa_set_option IEXA_OPT_SLOT_ORDER) under which the plan compiler reproduces the dump's COO
    structure bit for bit — the slot ORDER of ExaModels is data here, so a real dump only has to select a value"""
    tried = []
    for policy in POLICIES:
        m = ex.ExaModel(BUILDERS[name](), flags=ex.lib.IEXA_F_NO_DEVICE, library=L, slot_order=policy)
        if (m.meta.nvar, m.meta.ncon, m.meta.nnzj, m.meta.nnzh) != (len(d["x"]), len(d["y"]), len(d["jv"]), len(d["hv"])):
            tried.append(f"{POLICIES[policy]}: dimensions differ")
            continue
        ok = True
        for which, rk, ck in ((0, "jr", "jc"), (1, "hr", "hc")):
            r = np.zeros(max(len(d[rk]), 1), dtype=np.int64); c = np.zeros_like(r)
            L.hostcheck_structure(m.h, which, r.ctypes.data, c.ctypes.data)
            ok = ok and (r[: len(d[rk])] == d[rk]).all() and (c[: len(d[ck])] == d[ck]).all()
        if ok:
            return policy, m
        tried.append(f"{POLICIES[policy]}: structure differs")
    raise AssertionError("no slot-order policy reproduces the dump's structure: " + "; ".join(tried) +
                         " — add the order as a policy in gen.hpp (GenCompiler::kids / jr / hr) and oracle.c (kids)")


def compare_with_dump(path, hostcheck_lib, expect_policy=None):
    name = os.path.splitext(os.path.basename(path))[0]
    d = load(path)
    L = hostcheck_lib
    policy, m = identify_policy(name, d, L)
    print(f"{name}: structure reproduced bit for bit under {POLICIES[policy]}")
    if expect_policy is not None:
        assert policy == expect_policy
    x, y = np.ascontiguousarray(d["x"]), np.ascontiguousarray(d["y"])
    for which, key, n in ((2, "cons", len(y)), (3, "jv", len(d["jv"])), (4, "hv", len(d["hv"])), (1, "grad", len(x))):
        out = np.zeros(max(n, 1))
        L.hostcheck_eval_groups(m.h, which, x.ctypes.data, y.ctypes.data, d["sigma"], out.ctypes.data, None)
        assert_close(out[:n], d[key], key)
