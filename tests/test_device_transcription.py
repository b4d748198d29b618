"""Device-side transcription (src/transform.jl:2-38 iterators, :161-183 parameter functions, :618-633 measure
coefficients): iterator columns described by a closed form are generated ON the device, parameter functions are tapes the
engine evaluates into theta on the device.  The closed forms restate numpy's / InfiniteOpt's arithmetic operation by
operation, so the generated columns are BIT-IDENTICAL to what the host path uploads; theta of a transcendental parameter
function differs by the device libm's rounding (<= 2 ulp), of a polynomial one not at all."""
import numpy as np
import pytest

import iexa_b200 as ex
from iexa_b200 import models
from iexa_b200.core import ExaCore, Itr, linspace_col, linspace_mid_col, trapezoid_col, const_col
from iexa_b200.expr import DataSource, sin, cos
from conftest import assert_close, eval_point


def _columns(L, m, itr_id, n, K):
    out = []
    for c in range(n):
        a = np.zeros(K)
        assert L.iexa_debug_get_column(m.h, itr_id, c, a.ctypes.data) == 0, L.iexa_last_error()
        out.append(a)
    return out


def _toy(N):
    core = ExaCore()
    ds = DataSource()
    T = 2 * N - 1
    it = Itr(T, {"group_idx1": None}, {"t": linspace_mid_col(0.0, 60.0, N), "c": trapezoid_col("t"), "one": const_col(1.5)})
    it2 = Itr(N, {"j": None}, {"s": linspace_col(-1.0, 3.0, N), "w": trapezoid_col("s")})
    x = core.add_var(T)
    z = core.add_var(N)
    p = core.add_par_function(2.0 * (ds.t / 60.0) + ds.one, it)
    q = core.add_par_function(sin((2 * np.pi) * ds.t / 60.0) * cos(ds.t), it)
    core.add_con(x[ds.group_idx1] * p[ds.group_idx1] + q[ds.group_idx1] * ds.t, it, 0.0, 0.0)
    core.add_con(z[ds.j] * ds.s, it2, 0.0, 1.0)
    core.add_obj(ds.c * (x[ds.group_idx1] - p[ds.group_idx1]) ** 2, it)
    core.add_obj(ds.w * z[ds.j] ** 2, it2)
    return core, it, it2, p, q


def test_closed_forms_equal_numpy_bit_for_bit_on_the_host(hostcheck_lib):
    """the plan's own statement of the closed forms (used by host-only plans and by the test executor)"""
    L = hostcheck_lib
    for N in (2, 3, 17, 1000):
        core, it, it2, p, q = _toy(N)
        m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
        for got, name in zip(_columns(L, m, 1, 3, it.K), ("t", "c", "one")):
            assert np.array_equal(got, it.fps[name]), (N, name)
        for got, name in zip(_columns(L, m, 2, 2, it2.K), ("s", "w")):
            assert np.array_equal(got, it2.fps[name]), (N, name)


def test_described_quadrotor_is_the_same_model_as_the_uploaded_one(hostcheck_lib):
    a, b = models.quadrotor(40, "oc"), models.quadrotor(40, "oc", device_side=True)
    assert np.array_equal(a.theta_vec, b.theta_vec) and np.array_equal(a.x0_vec, b.x0_vec)
    assert len(b.par_functions) == 3
    ma = ex.ExaModel(a, flags=ex.lib.IEXA_F_NO_DEVICE, library=hostcheck_lib)
    mb = ex.ExaModel(b, flags=ex.lib.IEXA_F_NO_DEVICE, library=hostcheck_lib)
    assert (ma.meta.nvar, ma.meta.ncon, ma.meta.nnzj, ma.meta.nnzh, ma.cmeta.npar) == (mb.meta.nvar, mb.meta.ncon, mb.meta.nnzj, mb.meta.nnzh, mb.cmeta.npar)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [2, 5, 1000, 100_003])
def test_generated_columns_and_theta_on_the_device(N):
    L = ex.lib.load()
    core, it, it2, p, q = _toy(N)
    m = ex.ExaModel(core, device=0)
    for got, name in zip(_columns(L, m, 1, 3, it.K), ("t", "c", "one")):
        assert np.array_equal(got, it.fps[name]), f"generated column {name} is not bit-identical to numpy"
    for got, name in zip(_columns(L, m, 2, 2, it2.K), ("s", "w")):
        assert np.array_equal(got, it2.fps[name]), f"generated column {name} is not bit-identical to numpy"
    th = m.θ
    ref = core.theta_vec
    assert np.array_equal(th[p.offset:p.offset + p.length], ref[p.offset:p.offset + p.length]), "polynomial parameter function must be exact"
    d = np.abs(th[q.offset:q.offset + q.length] - ref[q.offset:q.offset + q.length])
    assert (d <= 4 * np.spacing(np.maximum(np.abs(ref[q.offset:q.offset + q.length]), 1e-300)) + 1e-17).all(), d.max()


@pytest.mark.gpu
def test_device_side_quadrotor_evaluates_like_the_host_built_one():
    import torch
    from oracle.oracle import OracleModel
    a, b = models.quadrotor(3000, "oc"), models.quadrotor(3000, "oc", device_side=True)
    om = OracleModel(a)
    m = ex.ExaModel(b, device=0)
    assert m.cmeta.n_kernels_specialised > 0
    x, y = eval_point(a, seed=3)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    z = lambda n: torch.full((n,), 7.0, dtype=torch.float64, device="cuda")
    assert_close(ex.cons_(m, xd, z(om.ncon)).cpu().numpy(), om.cons(x), "cons")
    assert_close(ex.jac_coord_(m, xd, z(om.nnzj)).cpu().numpy(), om.jac_coord(x), "jac_coord")
    assert_close(ex.hess_coord_(m, xd, yd, z(om.nnzh), 0.7).cpu().numpy(), om.hess_coord(x, y, 0.7), "hess_coord")
    assert_close(ex.obj(m, xd), om.obj(x), "obj")
    assert_close(ex.grad_(m, xd, z(om.nvar)).cpu().numpy(), om.grad(x), "grad")
