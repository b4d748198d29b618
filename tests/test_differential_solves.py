"""The reference's DIFFERENTIAL solve tests (test/solve.jl:2-27 "Test Problem 1", :48-95 "Test Problem 2" and its four
alternative objectives): solve the model once through JuMP's TranscriptionBackend — the fully expanded scalar NLP — and
once through ExaTranscriptionBackend, and compare objective and variable values at 1e-6.

Here the expanded NLP is the independent sympy restatement of tests/golden/make_sympy_golden.py (lambdified), the
generator route is model statements -> transform.py -> plan -> callbacks (oracle on the CPU, CUDA engine with `-m gpu`),
and both are solved by the same scipy driver that only sees NLPModels-style callbacks (tests/nlp_solve.py)."""
import os
import sys
import warnings

import numpy as np
import pytest
import sympy as sp

from conftest import ROOT, has_gpu
from nlp_solve import Callbacks, from_examodel, from_oracle, solve

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_sympy_golden as gold  # noqa: E402  (imports nothing from this repo)

TOL = 1e-6   # test/solve.jl:1


def from_sympy(P) -> Callbacks:
    """the expanded NLP behind the same callback interface: symbolic derivatives, lambdified once"""
    xs = P.vars
    n, m = len(xs), len(P.cons)
    pos = {s: i for i, s in enumerate(xs)}
    f = sp.lambdify([xs], P.obj, "numpy")
    g = sp.lambdify([xs], [sp.diff(P.obj, s) for s in xs], "numpy")
    c = sp.lambdify([xs], P.cons, "numpy")
    jr, jc, je = [], [], []
    for i, ci in enumerate(P.cons):
        for s in sorted(ci.free_symbols, key=lambda s: pos[s]):
            jr.append(i + 1); jc.append(pos[s] + 1); je.append(sp.diff(ci, s))
    jf = sp.lambdify([xs], je, "numpy")
    # Hessian of the Lagrangian: entries (row, col, owner, expr); owner -1 = objective
    hr, hc, howner, he = [], [], [], []
    for owner, e in [(-1, P.obj)] + list(enumerate(P.cons)):
        fs = sorted(e.free_symbols, key=lambda s: pos[s])
        for a in fs:
            da = sp.diff(e, a)
            for b in fs:
                if pos[b] > pos[a]:
                    continue
                d2 = sp.diff(da, b)
                if d2 != 0:
                    hr.append(pos[a] + 1); hc.append(pos[b] + 1); howner.append(owner); he.append(d2)
    hf = sp.lambdify([xs], he, "numpy")
    howner = np.array(howner)
    arr = lambda v, k: np.array([float(t) for t in v], dtype=np.float64).reshape(k)

    def hess_coord(x, y, sigma):
        v = arr(hf(x), len(he))
        w = np.where(howner < 0, sigma, (np.zeros(m) if y is None else np.asarray(y))[np.maximum(howner, 0)])
        return v * w

    x0 = np.array(P.x0, dtype=np.float64)
    return Callbacks(n, m, np.array(P.lvar), np.array(P.uvar), np.array(P.lcon), np.array(P.ucon), x0,
                     lambda x: float(f(x)), lambda x: arr(g(x), n), lambda x: arr(c(x), m), lambda x: arr(jf(x), len(je)),
                     hess_coord, (np.array(jr), np.array(jc)), (np.array(hr), np.array(hc)))


def _generator_core(variant):
    from iexa_b200 import infmodels
    from iexa_b200.transform import exa_core
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return exa_core(infmodels.solve_test_problem(variant))[0]


_EXPANDED = {}


def _expanded_solution(variant):
    if variant not in _EXPANDED:
        _EXPANDED[variant] = solve(from_sympy(gold.solve_problem(variant)))
    return _EXPANDED[variant]


@pytest.mark.parametrize("variant", [-1, 0, 1, 2, 3, 4])
def test_generator_route_and_expanded_nlp_solve_to_the_same_point(variant):
    from oracle.oracle import OracleModel
    ref = _expanded_solution(variant)
    res = solve(from_oracle(OracleModel(_generator_core(variant))))
    assert abs(res.fun - ref.fun) < TOL, (res.fun, ref.fun)                           # test/solve.jl:22,40,63,88
    assert np.max(np.abs(res.x - ref.x)) < 1e-4, np.max(np.abs(res.x - ref.x))        # same x layout on both sides (:23-26)


@pytest.mark.gpu
@pytest.mark.skipif(not has_gpu(), reason="needs a CUDA device")
@pytest.mark.parametrize("variant", [-1, 3])
def test_cuda_engine_solves_to_the_expanded_nlp_solution(variant):
    import iexa_b200 as ex
    ref = _expanded_solution(variant)
    m = ex.ExaModel(_generator_core(variant), device=0)
    res = solve(from_examodel(m))
    assert abs(res.fun - ref.fun) < TOL, (res.fun, ref.fun)
    assert np.max(np.abs(res.x - ref.x)) < 1e-4
