"""The BASELINE.json configurations at their FULL sizes on the GPU, through the C ABI:

  * direct parity with the oracle (all host threads) — the oracle finishes these sizes in seconds;
  * size-independent properties: hess_coord! is linear in (y, obj_weight) — exactly so for a power-of-two factor; a
    world-2 sharding of the plan tiles the unsharded result bit for bit; J·v and H·v (jprod! / hprod!) agree with
    central differences of cons! / of the gradient of the Lagrangian; the COO -> CSR pass preserves the value checksum.

Tolerance (north star): structure bit-exact, values 1e-12 relative / 1e-14 absolute."""
import os

import numpy as np
import pytest

import iexa_b200 as ex
from iexa_b200 import models
from conftest import assert_close, eval_point

pytestmark = pytest.mark.gpu


def _threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _eval_all(m, x, y, sigma):
    import torch
    xd, yd = _dev(x), _dev(y)
    z = lambda n: torch.zeros(max(int(n), 1), dtype=torch.float64, device="cuda")
    c, jv, hv, g = z(m.loc_ncon), z(m.loc_nnzj), z(m.loc_nnzh), z(m.meta.nvar)
    ex.cons_(m, xd, c); ex.jac_coord_(m, xd, jv); ex.hess_coord_(m, xd, yd, hv, sigma); ex.grad_(m, xd, g)
    f = ex.obj(m, xd)
    return f, g, c, jv, hv


def _parity(core, seed=0):
    from oracle import oracle as orc
    from oracle.oracle import OracleModel
    orc.set_threads(_threads())
    om = OracleModel(core)
    m = ex.ExaModel(core, device=0)
    assert m.cmeta.n_kernels_specialised > 0, m.L.iexa_engine_note(m.h).decode()
    assert (m.meta.nvar, m.meta.ncon, m.meta.nnzj, m.meta.nnzh) == (om.nvar, om.ncon, om.nnzj, om.nnzh)
    x, y = eval_point(core, seed=seed)
    f, g, c, jv, hv = _eval_all(m, x, y, 0.7)
    n = lambda t, k: t.cpu().numpy()[:k]
    fo = om.obj(x)
    assert abs(f - fo) <= 1e-14 + 1e-12 * max(abs(f), abs(fo), np.abs(x).sum() * 1e-3), ("obj", f, fo)
    assert_close(n(c, om.ncon), om.cons(x), "cons")
    assert_close(n(jv, om.nnzj), om.jac_coord(x), "jac_coord")
    assert_close(n(hv, om.nnzh), om.hess_coord(x, y, 0.7), "hess_coord")
    go = om.grad(x)
    assert np.allclose(n(g, om.nvar), go, rtol=1e-12, atol=1e-13), "grad"
    return m, om, x, y


def test_config3_quadrotor_1e6_supports_parity_structure_and_properties():
    """ESCAPE34/quadrotor.jl, OrthogonalCollocation(3), 10^6 public time supports (BASELINE configs[2])"""
    import torch
    core = models.quadrotor(1_000_000, "oc")
    m, om, x, y = _parity(core)
    # structure, bit-exact (int32 buffers, 1-based)
    for which, nnz, fn, ref in ((0, om.nnzj, ex.jac_structure_, om.jac_structure), (1, om.nnzh, ex.hess_structure_, om.hess_structure)):
        r = torch.zeros(nnz, dtype=torch.int32, device="cuda"); c = torch.zeros_like(r)
        fn(m, r, c)
        ro, co = ref()
        assert np.array_equal(r.cpu().numpy(), ro) and np.array_equal(c.cpu().numpy(), co), ("structure", which)
        del r, c
    # linearity of hess_coord! in (y, obj_weight): a factor 2 is exact in binary floating point
    xd, yd = _dev(x), _dev(y)
    h1 = torch.zeros(om.nnzh, dtype=torch.float64, device="cuda"); h2 = torch.zeros_like(h1)
    ex.hess_coord_(m, xd, yd, h1, 0.7); ex.hess_coord_(m, xd, 2.0 * yd, h2, 1.4)
    assert torch.equal(2.0 * h1, h2), "hess_coord! is not linear in (y, obj_weight)"
    del h1, h2
    # J v against central differences of cons!
    rng = np.random.default_rng(5)
    v = rng.uniform(-1, 1, om.nvar)
    vd = _dev(v)
    Jv = torch.zeros(om.ncon, dtype=torch.float64, device="cuda")
    ex.jprod_(m, xd, vd, Jv)
    eps = 1e-6
    cp = torch.zeros(om.ncon, dtype=torch.float64, device="cuda"); cm = torch.zeros_like(cp)
    ex.cons_(m, xd + eps * vd, cp); ex.cons_(m, xd - eps * vd, cm)
    fd = (cp - cm) / (2 * eps)
    err = (fd - Jv).abs().max().item()
    assert err <= 1e-7 * max(1.0, Jv.abs().max().item()), f"J v vs central differences: {err:.3e}"
    # H v against central differences of the gradient of the Lagrangian  sigma*grad f + J' y
    def grad_lag(xx):
        g = torch.zeros(om.nvar, dtype=torch.float64, device="cuda"); jt = torch.zeros_like(g)
        ex.grad_(m, xx, g); ex.jtprod_(m, xx, yd, jt)
        return 0.7 * g + jt
    Hv = torch.zeros(om.nvar, dtype=torch.float64, device="cuda")
    ex.hprod_(m, xd, yd, vd, Hv, 0.7)
    fdh = (grad_lag(xd + eps * vd) - grad_lag(xd - eps * vd)) / (2 * eps)
    errh = (fdh - Hv).abs().max().item()
    assert errh <= 1e-6 * max(1.0, Hv.abs().max().item()), f"H v vs central differences: {errh:.3e}"
    # the fused product kernels against the oracle's (out, v) reverse passes at the FULL size, north-star tolerance;
    # jprod! (no atomics) is bit-reproducible run to run
    assert_close(Jv.cpu().numpy(), om.jprod(x, v), "jprod")
    Jv2 = torch.full_like(Jv, 3.0)
    ex.jprod_(m, xd, vd, Jv2)
    assert torch.equal(Jv, Jv2), "jprod! is not bit-reproducible"
    del Jv2, cp, cm, fd, fdh
    assert_close(Hv.cpu().numpy(), om.hprod(x, y, v, 0.7), "hprod")
    w = rng.uniform(-1, 1, om.ncon)
    Jtw = torch.full((om.nvar,), 7.0, dtype=torch.float64, device="cuda")
    ex.jtprod_(m, xd, _dev(w), Jtw)
    assert_close(Jtw.cpu().numpy(), om.jtprod(x, w), "jtprod")
    assert m.L.iexa_engine_note(m.h) == b""


def test_config3_world2_sharding_tiles_the_full_model_bit_for_bit():
    import torch
    core = models.quadrotor(1_000_000, "oc")
    x, y = eval_point(core, seed=1)
    m1 = ex.ExaModel(core, device=0)
    f, g, c, jv, hv = _eval_all(m1, x, y, 0.7)
    ref = [t.cpu().numpy() for t in (c, jv, hv)]
    del m1, c, jv, hv, g
    got = [np.full_like(r, np.nan) for r in ref]
    fsum = 0.0
    for rank in range(2):
        m = ex.ExaModel(core, device=0, rank=rank, world=2)
        segs = []
        for which in range(3):
            arr = (ex.lib.Segment * 4096)()
            n = m.L.iexa_segments(m.h, which, arr, 4096)
            segs.append([(s.global_start, s.local_start, s.length) for s in arr[:n]])
        yl = np.zeros(max(m.loc_ncon, 1))
        for gs, ls, ln in segs[0]:
            yl[ls:ls + ln] = y[gs:gs + ln]
        fr, _, cl, jl, hl = _eval_all(m, x, yl, 0.7)
        fsum += fr
        for which, loc in enumerate((cl, jl, hl)):
            l = loc.cpu().numpy()
            for gs, ls, ln in segs[which]:
                got[which][gs:gs + ln] = l[ls:ls + ln]
        del m, cl, jl, hl
    for name, a, b in zip(("cons", "jac_coord", "hess_coord"), got, ref):
        assert np.array_equal(a, b), f"{name}: the two shards do not tile the unsharded result"
    assert abs(fsum - f) <= 1e-12 * max(1.0, abs(f))


def test_config3_general_lowering_builds_the_same_plan_at_full_size():
    """the model STATEMENTS of ESCAPE34/quadrotor.jl lowered by transform.py (the restatement of src/transform.jl) at 10^6
    supports give the same plan dimensions, x0 and θ bit for bit, and the same evaluations within the north-star tolerance
    as the hand transcription the benchmark uses (the collocation coefficients differ in the last bit: M2·inv(M1) evaluated
    numerically vs the closed form 0.75h, -0.25h, h, 0)"""
    import torch
    from iexa_b200 import infmodels
    from iexa_b200.transform import exa_core
    hand = models.quadrotor(1_000_000, "oc")
    x, y = eval_point(hand, seed=6)
    m1 = ex.ExaModel(hand, device=0)
    ref = [t.clone() for t in _eval_all(m1, x, y, 0.7)[2:]]
    dims = (m1.meta.nvar, m1.meta.ncon, m1.meta.nnzj, m1.meta.nnzh)
    del m1
    general, _ = exa_core(infmodels.quadrotor(1_000_000, "oc"))
    assert np.array_equal(general.x0_vec, hand.x0_vec) and np.array_equal(general.theta_vec, hand.theta_vec)
    m2 = ex.ExaModel(general, device=0)
    assert (m2.meta.nvar, m2.meta.ncon, m2.meta.nnzj, m2.meta.nnzh) == dims
    got = _eval_all(m2, x, y, 0.7)[2:]
    for name, a, b in zip(("cons", "jac_coord", "hess_coord"), got, ref):
        assert_close(a.cpu().numpy(), b.cpu().numpy(), name)


def test_config2_pandemic_1e5_time_supports_parity():
    """ESCAPE34/pandemic.jl, 10^5 time supports (+10 extra), 4 scenarios (BASELINE configs[1])"""
    _parity(models.pandemic(100_000, 4), seed=2)


def test_config4_opf_1e5_scenarios_parity():
    """ESCAPE34/opf.jl, two-stage stochastic AC-OPF, embedded 3-bus case, 10^5 scenarios (BASELINE configs[3])"""
    from iexa_b200 import opf
    from iexa_b200.transform import exa_core
    _parity(exa_core(opf.opf(None, num_supports=100_000))[0], seed=3)


def test_config5_farmer_1e5_scenarios_parity_and_csr_checksum():
    """examples/2stage_example.jl, 10^5 scenarios (BASELINE configs[4]); plus: COO -> CSR keeps the value checksum"""
    import torch
    m, om, x, y = _parity(models.farmer(100_000), seed=4)
    r = torch.zeros(om.nnzj, dtype=torch.int32, device="cuda"); c = torch.zeros_like(r)
    ex.jac_structure_(m, r, c)
    jv = torch.zeros(om.nnzj, dtype=torch.float64, device="cuda")
    ex.jac_coord_(m, _dev(x), jv)
    import ctypes as C
    L = ex.lib.load()
    h = C.c_void_p()
    st = torch.cuda.current_stream().cuda_stream
    assert L.iexa_csr_create(C.byref(h), om.ncon, om.nvar, om.nnzj, r.data_ptr(), c.data_ptr(), 4, 1, 0) == 0, L.iexa_last_error()
    out = torch.zeros(L.iexa_csr_nnz(h), dtype=torch.float64, device="cuda")
    assert L.iexa_csr_apply(h, jv.data_ptr(), out.data_ptr(), 1, st) == 0
    torch.cuda.synchronize()
    assert abs(out.sum().item() - jv.sum().item()) <= 1e-9 * max(1.0, jv.abs().sum().item())
    L.iexa_csr_destroy(h)
