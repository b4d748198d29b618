"""Programs with hundreds of live values (what an inline-expanded measure can produce,
transform.jl:430-435): the interpreter must switch to its global-scratch register file
(> 256 registers) and the specialised path must still compile and agree."""
import numpy as np
import pytest

import iexa_b200 as ex
from conftest import assert_close


def _product_model(n=300, K=5):
    """c_k = Π_i x[i,k]  and  f = Σ_k Π_i x[i,k]: a left-deep product keeps every partial product alive
    until the reverse sweep"""
    core = ex.ExaCore()
    x = core.add_var(n, K, start=1.0)
    it = ex.Itr(K, {"j": np.arange(1, K + 1)}, {})
    ds = ex.DataSource()
    e = x[1, ds.j]
    for i in range(2, n + 1):
        e = e * x[i, ds.j]
    core.add_con(e, it, 0.0, 2.0)
    core.add_obj(e, it)
    rng = np.random.default_rng(0)
    xv = 1.0 + 0.01 * rng.uniform(-1, 1, core.nvar)
    return core, xv


def test_register_count_exceeds_local_file(hostcheck_lib):
    core, x = _product_model()
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=hostcheck_lib)
    st = np.zeros(8, dtype=np.int64)
    assert hostcheck_lib.hostcheck_gen_stats(m.h, 0, 0, st.ctypes.data) == 0
    assert st[5] > 256, f"first-order program uses {st[5]} registers; the test must exceed 256"
    from oracle.oracle import OracleModel
    om = OracleModel(core)
    out = np.zeros(om.nnzj)
    assert hostcheck_lib.hostcheck_eval(m.h, 3, x.ctypes.data, None, 1.0, out.ctypes.data) == 0
    assert_close(out, om.jac_coord(x), "jac")


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [ex.lib.IEXA_F_NO_SPECIALISE, ex.lib.IEXA_F_DEFAULT])
def test_big_program_on_gpu(flags):
    import torch
    from oracle.oracle import OracleModel
    core, x = _product_model(n=300, K=37)
    om = OracleModel(core)
    m = ex.ExaModel(core, device=0, flags=flags)
    xd = torch.from_numpy(x).cuda()
    y = np.linspace(-1, 1, om.ncon)
    c = torch.zeros(om.ncon, dtype=torch.float64, device="cuda")
    jv = torch.zeros(om.nnzj, dtype=torch.float64, device="cuda")
    g = torch.zeros(om.nvar, dtype=torch.float64, device="cuda")
    assert_close(ex.cons_(m, xd, c).cpu().numpy(), om.cons(x), "cons")
    assert_close(ex.jac_coord_(m, xd, jv).cpu().numpy(), om.jac_coord(x), "jac")
    assert_close(ex.grad_(m, xd, g).cpu().numpy(), om.grad(x), "grad")
    assert_close(ex.obj(m, xd), om.obj(x), "obj")
    if om.nnzh <= 200000:
        hv = torch.zeros(om.nnzh, dtype=torch.float64, device="cuda")
        assert_close(ex.hess_coord_(m, xd, torch.from_numpy(y).cuda(), hv, 0.5).cpu().numpy(), om.hess_coord(x, y, 0.5), "hess")
