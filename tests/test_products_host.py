"""CPU tests of the matrix-free product programs (jprod! / jtprod! / hprod!): the plan compiler builds them from the
symbolic first / second order slots (gen.hpp: build_products) — per generator (what the AOT interpreter kernels run)
and per fused group with the two-phase scatter protocol (what the NVRTC kernels run: zero ranges, phase-0 plain
stores, phase-1 adds; plan.hpp: analyse_scatter) — checked against the oracle's restatement of ExaModels'
(out, v) reverse passes through the test-only host executor.  No GPU."""
import ctypes as C

import numpy as np
import pytest

import iexa_b200 as ex
from iexa_b200 import models
from conftest import assert_close, eval_point

CASES = {
    "ode_5x5": lambda: models.ode_5x5(),
    "quadrotor_oc": lambda: models.quadrotor(9, "oc"),
    "quadrotor_fd": lambda: models.quadrotor(12, "fd"),
    "pandemic": lambda: models.pandemic(7, 3),
    "farmer": lambda: models.farmer(11),
}


def _prod(L, m, which, use_groups, x, y, v, sigma, n):
    out = np.full(max(n, 1), 123.0)
    st = np.zeros(4, dtype=np.int64)
    rc = L.hostcheck_prod(m.h, which, use_groups, x.ctypes.data, None if y is None else y.ctypes.data, v.ctypes.data,
                          sigma, out.ctypes.data, st.ctypes.data)
    assert rc == 0
    return out[:n], st


@pytest.mark.parametrize("use_groups", [0, 1])
@pytest.mark.parametrize("name", list(CASES))
def test_product_programs_match_oracle(name, use_groups, hostcheck_lib):
    from oracle.oracle import OracleModel
    L = hostcheck_lib
    core = CASES[name]()
    om = OracleModel(core)
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    x, y = eval_point(core)
    rng = np.random.default_rng(5)
    v, w = rng.uniform(-1, 1, om.nvar), rng.uniform(-1, 1, om.ncon)
    assert_close(_prod(L, m, 5, use_groups, x, None, v, 1.0, om.ncon)[0], om.jprod(x, v), "jprod")
    assert_close(_prod(L, m, 6, use_groups, x, None, w, 1.0, om.nvar)[0], om.jtprod(x, w), "jtprod")
    assert_close(_prod(L, m, 7, use_groups, x, y, v, 0.7, om.nvar)[0], om.hprod(x, y, v, 0.7), "hprod")
    assert_close(_prod(L, m, 7, use_groups, x, None, v, 1.3, om.nvar)[0], om.hprod(x, None, v, 1.3), "hprod (objective only)")


def test_products_agree_with_coo_values(hostcheck_lib):
    """J v, J' w and H v from the product programs == the same products formed from jac_coord / hess_coord + structure"""
    from oracle.oracle import OracleModel
    L = hostcheck_lib
    core = models.quadrotor(15, "oc")
    om = OracleModel(core)
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    x, y = eval_point(core, seed=2)
    rng = np.random.default_rng(1)
    v, w = rng.uniform(-1, 1, om.nvar), rng.uniform(-1, 1, om.ncon)
    jr, jc = om.jac_structure(); jv = om.jac_coord(x)
    Jv = np.zeros(om.ncon); np.add.at(Jv, jr - 1, jv * v[jc - 1])
    Jtw = np.zeros(om.nvar); np.add.at(Jtw, jc - 1, jv * w[jr - 1])
    hr, hc = om.hess_structure(); hv = om.hess_coord(x, y, 0.7)
    Hv = np.zeros(om.nvar); np.add.at(Hv, hr - 1, hv * v[hc - 1])
    off = hr != hc
    np.add.at(Hv, hc[off] - 1, hv[off] * v[hr[off] - 1])
    np.testing.assert_allclose(_prod(L, m, 5, 1, x, None, v, 1.0, om.ncon)[0], Jv, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(_prod(L, m, 6, 1, x, None, w, 1.0, om.nvar)[0], Jtw, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(_prod(L, m, 7, 1, x, y, v, 0.7, om.nvar)[0], Hv, rtol=1e-12, atol=1e-13)


def test_scatter_phases_quadrotor(hostcheck_lib):
    """jtprod! of the quadrotor: the fused ODE rows (K = T) write the 19 variable blocks they reference as single writers
    (phase 0, plain stores, not zero-filled); x1, x3, x5 appear in no ODE row (only their derivatives do): those three
    blocks are zero-filled; the collocation / restriction / initial-condition rows add on top in phase 1"""
    L = hostcheck_lib
    core = models.quadrotor(9, "oc")
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    x, y = eval_point(core)
    w = np.ones(m.meta.ncon)
    _, st = _prod(L, m, 6, 1, x, None, w, 1.0, m.meta.nvar)
    T = 2 * 9 - 1
    assert st[0] == 1 and st[1] == 19 and st[2] == 3 * T and st[3] == 4, st
    _, st = _prod(L, m, 7, 1, x, y, np.ones(m.meta.nvar), 1.0, m.meta.nvar)
    assert st[0] >= 1 and st[1] >= 7, st


def test_hprod_rider_quadrotor(hostcheck_lib):
    """hprod! of the quadrotor: the objective group (K = T) stores its 10 diagonal blocks directly; the fused ODE rows walk the
    same supports and ride on it — `out[i] += v` by the thread that just stored out[i] — so the whole product is ONE launch
    without atomics or a second phase; jtprod! keeps its second phase (the collocation rows walk other supports)"""
    L = hostcheck_lib
    m = ex.ExaModel(models.quadrotor(9, "oc"), flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    info = np.zeros(3, dtype=np.int64)
    assert L.hostcheck_scatter_info(m.h, 1, info.ctypes.data) == 0
    assert list(info) == [1, 1, 0], info
    assert L.hostcheck_scatter_info(m.h, 0, info.ctypes.data) == 0
    assert info[0] == 1 and info[2] == 3, info


def test_scatter_phases_product_iterator(hostcheck_lib):
    """pandemic: the column-major index of a variable over the (t, xi) product iterator is linear in k (mixed-radix digits
    recombine), so the scenario blocks are single-writer; u(t) is shared by the scenarios -> atomics on a zeroed range"""
    L = hostcheck_lib
    core = models.pandemic(7, 3)
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    x, _ = eval_point(core)
    _, st = _prod(L, m, 6, 1, x, None, np.ones(m.meta.ncon), 1.0, m.meta.nvar)
    assert st[0] >= 1 and st[1] >= 4 and 0 < st[2] < m.meta.nvar, st


def test_products_shape_classes(hostcheck_lib):
    from oracle.oracle import OracleModel
    from iexa_b200 import opf
    from iexa_b200.transform import exa_core
    L = hostcheck_lib
    core, _ = exa_core(opf.opf(opf.synthetic_grid(12), num_supports=4))
    om = OracleModel(core)
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    x, y = eval_point(core)
    x = np.where(np.isfinite(x), x, 0.0)
    rng = np.random.default_rng(5)
    v, w = rng.uniform(-1, 1, om.nvar), rng.uniform(-1, 1, om.ncon)
    for mode in (0, 1):
        assert L.iexa_debug_set_class_mode(m.h, mode) == 0
        assert_close(_prod(L, m, 5, 1, x, None, v, 1.0, om.ncon)[0], om.jprod(x, v), "jprod")
        assert_close(_prod(L, m, 6, 1, x, None, w, 1.0, om.nvar)[0], om.jtprod(x, w), "jtprod")
        assert_close(_prod(L, m, 7, 1, x, y, v, 0.7, om.nvar)[0], om.hprod(x, y, v, 0.7), "hprod")
    # the product module (second NVRTC translation unit) cross-compiles for sm_100a
    n = L.iexa_debug_codegen_source(m.h, None, -1)
    assert n > 1000


@pytest.mark.parametrize("name", ["quadrotor_oc", "pandemic", "farmer"])
def test_product_kernels_compile_for_sm100a(name, hostcheck_lib):
    L = hostcheck_lib
    m = ex.ExaModel(CASES[name](), flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    n = L.iexa_debug_codegen_source(m.h, None, -1)       # cap < 0: the product module
    buf = C.create_string_buffer(n + 1)
    L.iexa_debug_codegen_source(m.h, buf, -(n + 1))
    src = buf.value.decode()
    assert "iexa_cb_jprod" in src and "iexa_cb_jtprod_p" in src and "iexa_cb_hprod_p" in src or name == "farmer"
    assert "atomicAdd" in src or name != "quadrotor_oc"
    nb = C.c_int64(-1)                                     # in: -1 selects the product module
    rc = L.iexa_debug_codegen_compile(m.h, C.byref(nb))
    assert rc == 0, L.iexa_last_error().decode()
    assert nb.value > 1000
