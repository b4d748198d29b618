"""SURVEY §8 a13 (solver hand-off): the reference hands the model to MadNLP / Ipopt (ext/InfiniteExaModelsMadNLP.jl:49-50,
ext/InfiniteExaModelsIpopt.jl:48-49).  Neither can run here, so tests/gpu_ipm.py is a primal-dual interior-point loop with the same
structure whose iterate never leaves the device: per iteration ONE fused `iexa_eval3` launch (cons! + jac_coord! + hess_coord!),
`iexa_grad`, `iexa_obj_device` / `iexa_cons` per line-search trial, COO values scattered into the KKT matrix on the device.

Known answers: the reference's own solve-level goldens (test/madnlp.jl:18,42; test/solve.jl:146,154,187,206) at the reference's
tolerance (1e-6), the analytic optimum of test/solve.jl:134-156 (306.5 / 276.265) at 1e-7, parameter updates IN PLACE between solves
(no re-planning: infiniteopt_backend.jl:511-548)."""
import numpy as np
import pytest

import iexa_b200 as ex
from iexa_b200 import models
from gpu_ipm import DeviceCallbacks, OracleCallbacks, solve_on_device

TOL = 1e-6


def test_the_interior_point_loop_reproduces_the_goldens_on_the_oracle():
    """the ALGORITHM, on the CPU oracle (no GPU needed)"""
    from oracle.oracle import OracleModel
    res = solve_on_device(OracleCallbacks(OracleModel(models.ode_5x5())))
    assert abs(res["f"] - (-1.2784599867885884e+01)) < TOL and res["err"] < 1e-8          # test/madnlp.jl:42
    core, p1, p2 = models.rosenbrock_param(100.0, 1.0)
    om = OracleModel(core)
    res = solve_on_device(OracleCallbacks(om), x0=np.full(core.nvar, 1.0))
    assert abs(res["f"] - 306.4999755050365) < 1e-4 and abs(res["f"] - 306.5) < 1e-7        # test/solve.jl:146 / analytic
    om.set_parameter(p1.offset, [90.0]); om.set_parameter(p2.offset, [1.3])
    res = solve_on_device(OracleCallbacks(om), x0=res["x"].numpy())
    assert abs(res["f"] - 276.26497794903645) < 1e-4 and abs(res["f"] - 276.265) < 1e-7     # test/solve.jl:154 / analytic


@pytest.mark.gpu
def test_device_resident_solve_of_the_5x5_model():
    import torch
    m = ex.ExaModel(models.ode_5x5(), device=0)
    cb = DeviceCallbacks(m, ex)
    res = solve_on_device(cb)
    assert res["x"].is_cuda and res["y"].is_cuda
    assert abs(res["f"] - (-1.2784599867885884e+01)) < TOL, res["f"]                        # test/madnlp.jl:18,42
    assert res["err"] < 1e-8 and res["iters"] < 60
    assert res["counts"]["eval3"] == res["iters"]      # ONE fused cons + jac + hess launch per iteration
    # the same answer as the host-buffer (Ipopt-style) route through scipy
    from nlp_solve import from_examodel, solve
    assert abs(solve(from_examodel(m)).fun - res["f"]) < TOL


@pytest.mark.gpu
def test_device_resident_resolve_after_in_place_parameter_updates():
    core, p1, p2 = models.rosenbrock_param(100.0, 1.0)
    m = ex.ExaModel(core, device=0)
    res = solve_on_device(DeviceCallbacks(m, ex), x0=np.full(core.nvar, 1.0))
    assert abs(res["f"] - 306.4999755050365) < 1e-4 and abs(res["f"] - 306.5) < 1e-7, res["f"]   # test/solve.jl:146
    m.set_parameter(p1, [90.0]); m.set_parameter(p2, [1.3])                                        # no re-planning
    res = solve_on_device(DeviceCallbacks(m, ex), x0=res["x"].cpu().numpy())
    assert abs(res["f"] - 276.26497794903645) < 1e-4 and abs(res["f"] - 276.265) < 1e-7, res["f"]  # test/solve.jl:154


@pytest.mark.gpu
def test_device_resident_resolve_after_parameter_function_updates():
    core, pf1, pf2 = models.param_function_model(0.2, np.sin)
    m = ex.ExaModel(core, device=0)
    res = solve_on_device(DeviceCallbacks(m, ex))
    assert abs(res["f"] - 0.48292223509341475) < TOL, res["f"]                                     # test/solve.jl:187
    ts, ss = np.linspace(0, 1, 3), np.linspace(2, 3, 3)
    m.set_parameter(pf1, np.cos(ts)); m.set_parameter(pf2, np.sin(ts)[:, None] * ss[None, :] + 0.8)
    res = solve_on_device(DeviceCallbacks(m, ex), x0=res["x"].cpu().numpy())
    assert abs(res["f"] - 0.8155916466182952) < TOL, res["f"]                                      # test/solve.jl:206


@pytest.mark.gpu
def test_device_resident_solve_of_the_two_stage_farmer_problem():
    """examples/2stage_example.jl:20-37 (BASELINE configs[4]) at 20 scenarios — an LP with 123 variables, 101 rows, bounds on
    everything: the device-resident loop against the same loop on the oracle, feasibility and objective checked independently"""
    core = models.farmer(20)
    m = ex.ExaModel(core, device=0)
    res = solve_on_device(DeviceCallbacks(m, ex), tol=1e-8, max_iter=300)
    assert res["err"] < 1e-7, res
    from oracle.oracle import OracleModel
    om = OracleModel(core)
    x = res["x"].cpu().numpy()
    c = om.cons(x)
    assert (c >= om.lcon - 1e-6).all() and (c <= om.ucon + 1e-6).all() and (x >= om.lvar - 1e-9).all() and (x <= om.uvar + 1e-9).all()
    assert abs(om.obj(x) - res["f"]) <= 1e-12 * max(1.0, abs(res["f"]))
    ref = solve_on_device(OracleCallbacks(om), tol=1e-8, max_iter=300)
    assert abs(ref["f"] - res["f"]) < 1e-6 * max(1.0, abs(ref["f"])), (ref["f"], res["f"])
