"""``ExaTranscriptionBackend`` (backend.py) driven like the reference's own integration tests
(test/solve.jl:134-209, test/madnlp.jl:4-18): build from an InfiniteModel, optimize, update parameters and
start values IN PLACE (no plan rebuild), re-optimize, query values — on the GPU engine through the C ABI."""
import numpy as np
import pytest

from iexa_b200 import infmodels
from iexa_b200.backend import ExaTranscriptionBackend, Results
from nlp_solve import from_examodel, solve

pytestmark = pytest.mark.gpu


def scipy_solver(model, x0, y0, **options):
    """stand-in for MadNLP / Ipopt: an interior-point method fed only through the NLPModels callbacks"""
    res = solve(from_examodel(model), x0=x0)
    mult = np.asarray(res.v[0]) if len(res.v) else np.zeros(model.meta.ncon)
    return Results(solution=res.x, multipliers=mult, objective=float(res.fun), iter=int(res.nit))


def test_build_optimize_and_query():
    m = infmodels.ode_5x5()
    b = ExaTranscriptionBackend(scipy_solver, device=0).build_transformation_backend(m)
    assert b.transformation_backend_ready()
    assert (b.model.meta.nvar, b.model.meta.ncon) == (51, 70)                  # test/ipopt.jl:183-186
    res = b.optimize()
    assert abs(res.objective - (-1.2784599867885884e+01)) < 1e-6              # test/madnlp.jl:42
    z = m.finite_vars[0]; y = m.infinite_vars[0]
    assert b.map_value(y).shape == (5, 5) and isinstance(b.map_value(z), float)
    assert (b.map_value(y) >= -1e-8).all()
    assert len(b.map_dual(m.constraints[0])) == 25
    # warm start: the previous solution becomes x0 / y0 (infiniteopt_backend.jl:595-603)
    assert b.warmstart_backend()
    assert np.array_equal(b.model.meta.x0, res.solution)
    res2 = b.optimize()
    assert abs(res2.objective - res.objective) < 1e-6   # (Ipopt's 8 -> 5 iteration drop, test/ipopt.jl:180,195, needs Ipopt)


def test_parameter_updates_do_not_rebuild():
    m, p1, p2 = infmodels.rosenbrock_param(100.0, 1.0)
    b = ExaTranscriptionBackend(scipy_solver, device=0).build_transformation_backend(m)
    for v in m.infinite_vars:
        assert b.update_start_value(v, 1.0)
    handle = b.model.h.value
    assert abs(b.optimize().objective - 306.4999755050365) < 1e-4             # test/solve.jl:146
    assert b.update_parameter_value(p1, 90.0) and b.update_parameter_value(p2, 1.3)
    assert b.model.h.value == handle, "the plan must be updated in place"
    assert list(b.model.θ) == [90.0, 1.3]
    b.warmstart_backend()
    assert abs(b.optimize().objective - 276.26497794903645) < 1e-4            # test/solve.jl:154
    p3 = m.finite_parameter(43.0)                                             # unknown to the built backend
    assert b.update_parameter_value(p3, 50.0) is False                        # -> needs a rebuild (test/solve.jl:157-160)


def test_parameter_function_updates():
    m, f1, f2 = infmodels.param_function_model(0.2, np.sin)
    b = ExaTranscriptionBackend(scipy_solver, device=0).build_transformation_backend(m)
    assert abs(b.optimize().objective - 0.48292223509341475) < 1e-6          # test/solve.jl:187
    assert np.allclose(b.map_value(f1), np.sin([0.0, 0.5, 1.0]))
    assert b.update_parameter_value(f1, np.cos)
    assert b.update_parameter_value(f2, lambda t, s: np.sin(t) * s + 0.8)
    expected = [0.8, 1.758851077208406, 2.4829419696157933, 0.8, 1.9985638465105076, 2.9036774620197416, 0.8,
                2.238276615812609, 3.324412954423689]                          # test/solve.jl:202
    assert np.allclose(b.map_value(f2).reshape(-1, order="F"), expected, rtol=0, atol=1e-15)
    b.warmstart_backend()
    assert abs(b.optimize().objective - 0.8155916466182952) < 1e-6           # test/solve.jl:206


def test_optimize_without_solver_errors():
    b = ExaTranscriptionBackend(None, device=0).build_transformation_backend(infmodels.ode_5x5())
    with pytest.raises(RuntimeError, match="No solver"):
        b.optimize()
