"""Edge cases: empty iterators (a domain restriction that removes every support), constant-only
generators (ExaModels.Null, transform.jl:393,741), models without constraints / without objective,
a single support, supports counts around the block size."""
import numpy as np
import pytest

import iexa_b200 as ex
from iexa_b200 import models
from conftest import assert_close


def _edge_core(K=5):
    core = ex.ExaCore()
    x = core.add_var(K, start=0.5)
    z = core.add_var(1, start=2.0)
    ds = ex.DataSource()
    it = ex.Itr(K, {"i": np.arange(1, K + 1)}, {"p": np.linspace(1, 2, K)})
    empty = it.filtered(np.zeros(K, dtype=bool))              # restriction removed every support: K = 0
    core.add_con(ex.sin(x[ds.i]) * z[1], empty, 0.0, 0.0)     # contributes no rows
    core.add_con(ex.Null(3.5), it, 0.0, 5.0)                  # constant rows, no Jacobian entries
    core.add_con(x[ds.i] * x[ds.i] + ds.p * z[1], it, -1.0, 1.0)
    core.add_obj(ex.Null(7.25))                               # constant objective term
    core.add_obj(ex.exp(x[ds.i]) + z[1], it)
    return core


def test_edge_core_host(hostcheck_lib):
    from oracle.oracle import OracleModel
    L = hostcheck_lib
    core = _edge_core()
    om = OracleModel(core)
    assert om.ncon == 10
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    assert (m.meta.ncon, m.meta.nnzj, m.meta.nnzh) == (om.ncon, om.nnzj, om.nnzh)
    x = np.linspace(0.1, 0.9, core.nvar); y = np.linspace(-1, 1, om.ncon)
    for which, ref in ((0, [om.obj(x)]), (1, om.grad(x)), (2, om.cons(x)), (3, om.jac_coord(x)), (4, om.hess_coord(x, y, 0.7))):
        for fn in ("hostcheck_eval", "hostcheck_eval_groups"):
            out = np.zeros(max(len(ref), 1))
            args = [m.h, which, x.ctypes.data, y.ctypes.data, 0.7 if which == 4 else 1.0, out.ctypes.data]
            if fn.endswith("groups"):
                args.append(None)
            assert getattr(L, fn)(*args) == 0
            assert_close(out[:len(ref)], np.asarray(ref), f"{fn} {which}")
    assert abs(om.obj(x) - (7.25 + np.exp(x[:5]).sum() + 5 * x[5])) < 1e-12
    assert np.allclose(om.cons(x)[:5], 3.5)


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [ex.lib.IEXA_F_NO_SPECIALISE, ex.lib.IEXA_F_DEFAULT])
def test_edge_core_gpu(flags):
    import torch
    from oracle.oracle import OracleModel
    core = _edge_core()
    om = OracleModel(core)
    m = ex.ExaModel(core, device=0, flags=flags)
    x = np.linspace(0.1, 0.9, core.nvar); y = np.linspace(-1, 1, om.ncon)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    c = torch.full((om.ncon,), 9.0, dtype=torch.float64, device="cuda")
    jv = torch.full((om.nnzj,), 9.0, dtype=torch.float64, device="cuda")
    hv = torch.full((om.nnzh,), 9.0, dtype=torch.float64, device="cuda")
    g = torch.full((om.nvar,), 9.0, dtype=torch.float64, device="cuda")
    assert_close(ex.obj(m, xd), om.obj(x), "obj")
    assert_close(ex.cons_(m, xd, c).cpu().numpy(), om.cons(x), "cons")
    assert_close(ex.jac_coord_(m, xd, jv).cpu().numpy(), om.jac_coord(x), "jac")
    assert_close(ex.hess_coord_(m, xd, yd, hv, 0.7).cpu().numpy(), om.hess_coord(x, y, 0.7), "hess")
    assert_close(ex.grad_(m, xd, g).cpu().numpy(), om.grad(x), "grad")


@pytest.mark.gpu
def test_models_without_constraints_or_objective():
    import torch
    from oracle.oracle import OracleModel
    # objective only
    core = ex.ExaCore()
    x = core.add_var(3, start=1.0)
    core.add_obj(ex.abs2(x[1]) + x[2] * x[3])
    m = ex.ExaModel(core, device=0)
    om = OracleModel(core)
    xv = np.array([1.0, 2.0, 3.0]); xd = torch.from_numpy(xv).cuda()
    assert (m.meta.ncon, m.meta.nnzj) == (0, 0)
    assert_close(ex.obj(m, xd), om.obj(xv), "obj")
    hv = torch.zeros(om.nnzh, dtype=torch.float64, device="cuda")
    assert_close(ex.hess_coord_(m, xd, None, hv, 2.0).cpu().numpy(), om.hess_coord(xv, None, 2.0), "hess")
    ex.cons_(m, xd, torch.zeros(1, dtype=torch.float64, device="cuda"))       # no rows: a no-op, not an error
    # constraints only (feasibility problem: no objective, transform.jl:790-794)
    core = ex.ExaCore()
    x = core.add_var(2, start=1.0)
    core.add_con(x[1] * x[2], None, 1.0, 1.0)
    m = ex.ExaModel(core, device=0)
    om = OracleModel(core)
    xv = np.array([2.0, 3.0]); xd = torch.from_numpy(xv).cuda()
    assert ex.obj(m, xd) == 0.0
    g = torch.full((2,), 5.0, dtype=torch.float64, device="cuda")
    assert (ex.grad_(m, xd, g).cpu().numpy() == 0).all()
    c = torch.zeros(1, dtype=torch.float64, device="cuda")
    assert_close(ex.cons_(m, xd, c).cpu().numpy(), om.cons(xv), "cons")


@pytest.mark.gpu
@pytest.mark.parametrize("N", [2, 3, 64, 65, 127, 128, 129, 257])
def test_support_counts_around_the_block_size(N):
    """ragged last blocks / last warps of the staged write-out"""
    import torch
    from oracle.oracle import OracleModel
    core = models.quadrotor(N, "oc")
    om = OracleModel(core)
    m = ex.ExaModel(core, device=0)
    rng = np.random.default_rng(N)
    x = core.x0_vec + 0.1 * rng.uniform(-1, 1, core.nvar); y = rng.uniform(-1, 1, core.ncon)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    # guard cells behind the outputs must stay untouched
    jv = torch.full((om.nnzj + 64,), -7.0, dtype=torch.float64, device="cuda")
    hv = torch.full((om.nnzh + 64,), -7.0, dtype=torch.float64, device="cuda")
    c = torch.full((om.ncon + 64,), -7.0, dtype=torch.float64, device="cuda")
    ex.cons_(m, xd, c); ex.jac_coord_(m, xd, jv); ex.hess_coord_(m, xd, yd, hv, 1.0)
    assert_close(c.cpu().numpy()[: om.ncon], om.cons(x), "cons")
    assert_close(jv.cpu().numpy()[: om.nnzj], om.jac_coord(x), "jac")
    assert_close(hv.cpu().numpy()[: om.nnzh], om.hess_coord(x, y, 1.0), "hess")
    assert (c[om.ncon:] == -7.0).all() and (jv[om.nnzj:] == -7.0).all() and (hv[om.nnzh:] == -7.0).all()
