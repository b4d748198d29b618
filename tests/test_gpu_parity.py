"""GPU parity: every NLPModels callback through the C ABI (CUDA kernels) against the oracle, on the
same seeded inputs, for the AOT tape-interpreter kernels and the NVRTC-specialised kernels,
with device buffers (MadNLP-style) and host buffers (Ipopt-style).

Bar (BASELINE.json north_star): sparsity structure bit-exact; fp64 values within 1e-12 relative
/ 1e-14 absolute."""
import numpy as np
import pytest

import iexa_b200 as ex
from iexa_b200 import models
from conftest import assert_close, eval_point

pytestmark = pytest.mark.gpu

CASES = {
    "ode_5x5": lambda: models.ode_5x5(),
    "quadrotor_oc_40": lambda: models.quadrotor(40, "oc"),
    "quadrotor_fd_100": lambda: models.quadrotor(100, "fd"),     # BASELINE configs[0]
    "quadrotor_oc_ragged": lambda: models.quadrotor(333, "oc"),  # support count not a multiple of the block
    "pandemic_50x4": lambda: models.pandemic(50, 4),
    "pandemic_100x128": lambda: models.pandemic(100, 128),       # largest pandemic case of the reference's study grid (ESCAPE34/run_cases_gpu.jl:100)
    "farmer_1000": lambda: models.farmer(1000),
}
MODES = {"interp": ex.lib.IEXA_F_NO_SPECIALISE, "nvrtc": ex.lib.IEXA_F_DEFAULT}


@pytest.fixture(scope="module")
def oracle_cache():
    return {}


def _oracle(cache, name):
    from oracle.oracle import OracleModel
    if name not in cache:
        core = CASES[name]()
        cache[name] = (core, OracleModel(core))
    return cache[name]


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("name", list(CASES))
def test_callbacks_device_buffers(name, mode, oracle_cache):
    import torch
    core, om = _oracle(oracle_cache, name)
    m = ex.ExaModel(core, device=0, flags=MODES[mode])
    if mode == "nvrtc":
        assert m.cmeta.n_kernels_specialised > 0, "NVRTC specialisation did not produce kernels"
    else:
        assert m.cmeta.n_kernels_specialised == 0
    assert (m.meta.nvar, m.meta.ncon, m.meta.nnzj, m.meta.nnzh) == (om.nvar, om.ncon, om.nnzj, om.nnzh)
    x, y = eval_point(core)
    dev = torch.device("cuda:0")
    xd = torch.from_numpy(x).to(dev)
    yd = torch.from_numpy(y).to(dev)
    # structure: bit-exact, int32 and int64
    for dt in (torch.int32, torch.int64):
        r = torch.zeros(max(om.nnzj, 1), dtype=dt, device=dev); c = torch.zeros_like(r)
        ex.jac_structure_(m, r, c)
        ro, co = om.jac_structure()
        assert (r.cpu().numpy()[: om.nnzj] == ro).all() and (c.cpu().numpy()[: om.nnzj] == co).all()
        r = torch.zeros(max(om.nnzh, 1), dtype=dt, device=dev); c = torch.zeros_like(r)
        ex.hess_structure_(m, r, c)
        ro, co = om.hess_structure()
        assert (r.cpu().numpy()[: om.nnzh] == ro).all() and (c.cpu().numpy()[: om.nnzh] == co).all()
    assert_close(ex.obj(m, xd), om.obj(x), "obj")
    g = torch.full((om.nvar,), 7.0, dtype=torch.float64, device=dev)
    assert_close(ex.grad_(m, xd, g).cpu().numpy(), om.grad(x), "grad")
    c = torch.full((max(om.ncon, 1),), 7.0, dtype=torch.float64, device=dev)
    assert_close(ex.cons_(m, xd, c).cpu().numpy()[: om.ncon], om.cons(x), "cons")
    jv = torch.full((max(om.nnzj, 1),), 7.0, dtype=torch.float64, device=dev)
    assert_close(ex.jac_coord_(m, xd, jv).cpu().numpy()[: om.nnzj], om.jac_coord(x), "jac_coord")
    hv = torch.full((max(om.nnzh, 1),), 7.0, dtype=torch.float64, device=dev)
    assert_close(ex.hess_coord_(m, xd, yd, hv, obj_weight=0.7).cpu().numpy()[: om.nnzh], om.hess_coord(x, y, 0.7), "hess_coord")
    # objective-only Hessian (y = nothing)
    assert_close(ex.hess_coord_(m, xd, None, hv, obj_weight=1.3).cpu().numpy()[: om.nnzh], om.hess_coord(x, None, 1.3), "hess_coord(obj only)")
    # matrix-free products
    rng = np.random.default_rng(5)
    v = rng.uniform(-1, 1, om.nvar); w = rng.uniform(-1, 1, om.ncon)
    vd, wd = torch.from_numpy(v).to(dev), torch.from_numpy(w).to(dev)
    # (fused product kernels: outputs start as garbage — every entry must be (re)written, the bar is that of the values)
    out = torch.full((max(om.ncon, 1),), 7.0, dtype=torch.float64, device=dev)
    got = ex.jprod_(m, xd, vd, out).cpu().numpy()[: om.ncon]
    assert_close(got, om.jprod(x, v), "jprod")
    out2 = torch.full((max(om.ncon, 1),), -3.0, dtype=torch.float64, device=dev)
    assert (ex.jprod_(m, xd, vd, out2).cpu().numpy()[: om.ncon] == got).all(), "jprod! must be bit-reproducible"
    out = torch.full((om.nvar,), 7.0, dtype=torch.float64, device=dev)
    assert_close(ex.jtprod_(m, xd, wd, out).cpu().numpy(), om.jtprod(x, w), "jtprod")
    out = torch.full((om.nvar,), 7.0, dtype=torch.float64, device=dev)
    assert_close(ex.hprod_(m, xd, yd, vd, out, obj_weight=0.7).cpu().numpy(), om.hprod(x, y, v, 0.7), "hprod")
    out = torch.full((om.nvar,), 7.0, dtype=torch.float64, device=dev)
    assert_close(ex.hprod_(m, xd, None, vd, out, obj_weight=1.3).cpu().numpy(), om.hprod(x, None, v, 1.3), "hprod (objective only)")
    if mode == "nvrtc":
        assert m.L.iexa_engine_note(m.h) == b"", m.L.iexa_engine_note(m.h)


@pytest.mark.parametrize("name", ["ode_5x5", "quadrotor_fd_100", "pandemic_50x4"])
def test_callbacks_host_buffers(name, oracle_cache):
    """Ipopt-style call: numpy buffers, copies inside the C-ABI call."""
    core, om = _oracle(oracle_cache, name)
    m = ex.ExaModel(core, device=0)
    x, y = eval_point(core, seed=3)
    assert_close(ex.obj(m, x), om.obj(x), "obj")
    assert_close(ex.grad_(m, x, np.zeros(om.nvar)), om.grad(x), "grad")
    assert_close(ex.cons_(m, x, np.zeros(om.ncon)), om.cons(x), "cons")
    assert_close(ex.jac_coord_(m, x, np.zeros(om.nnzj)), om.jac_coord(x), "jac")
    assert_close(ex.hess_coord_(m, x, y, np.zeros(om.nnzh), 0.5), om.hess_coord(x, y, 0.5), "hess")
    r, c = np.zeros(om.nnzj, dtype=np.int64), np.zeros(om.nnzj, dtype=np.int64)
    ex.jac_structure_(m, r, c)
    ro, co = om.jac_structure()
    assert (r == ro).all() and (c == co).all()


def test_host_buffers_same_x_flag(oracle_cache):
    """IEXA_MEM_HOST_SAME_X (Ipopt's new_x == false): the device copy of x is reused by the calls that say so, and a
    call with new_x picks up a changed x."""
    from iexa_b200.model import bind
    core, om = _oracle(oracle_cache, "quadrotor_oc_40")
    m = ex.ExaModel(core, device=0)
    x, y = eval_point(core)
    xb = x.copy()
    c, jv, hv = np.zeros(om.ncon), np.zeros(om.nnzj), np.zeros(om.nnzh)
    f_cons = bind(m, "cons", xb, c)
    f_jac = bind(m, "jac_coord", xb, jv, new_x=False)
    f_hess = bind(m, "hess_coord", xb, hv, y, 0.7, new_x=False)
    f_jac()                                   # nothing uploaded yet: falls back to a plain upload
    assert_close(jv, om.jac_coord(x), "jac_coord before any upload")
    for trial in range(2):
        f_cons(); f_jac(); f_hess()
        assert_close(c, om.cons(xb), "cons"); assert_close(jv, om.jac_coord(xb), "jac_coord (same x)")
        assert_close(hv, om.hess_coord(xb, y, 0.7), "hess_coord (same x)")
        xb += 0.05 * np.cos(np.arange(xb.size))   # next iterate: cons! is called with new_x and re-uploads


@pytest.mark.parametrize("mode", list(MODES))
def test_callbacks_capture_into_a_cuda_graph(mode, oracle_cache):
    """Device-memory callbacks only enqueue kernels on the caller's stream (no allocation, no synchronisation), so a
    solver can capture an iteration's cons! + jac_coord! + hess_coord! into a CUDA graph and replay it: the replay
    must see the CURRENT contents of x / y and reproduce the oracle."""
    import torch
    from iexa_b200.model import bind
    core, om = _oracle(oracle_cache, "quadrotor_oc_40")
    m = ex.ExaModel(core, device=0, flags=MODES[mode])
    x, y = eval_point(core, seed=11)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    c = torch.zeros(om.ncon, dtype=torch.float64, device="cuda")
    jv = torch.zeros(om.nnzj, dtype=torch.float64, device="cuda")
    hv = torch.zeros(om.nnzh, dtype=torch.float64, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):  # bound on the capture stream
        f_c, f_j, f_h = bind(m, "cons", xd, c), bind(m, "jac_coord", xd, jv), bind(m, "hess_coord", xd, hv, yd, 0.7)
        f_c(); f_j(); f_h()     # warm-up outside the capture
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        f_c(); f_j(); f_h()
    for trial in range(2):
        c.zero_(); jv.zero_(); hv.zero_()
        g.replay(); torch.cuda.synchronize()
        xh = xd.cpu().numpy()
        assert_close(c.cpu().numpy(), om.cons(xh), "cons (graph replay)")
        assert_close(jv.cpu().numpy(), om.jac_coord(xh), "jac_coord (graph replay)")
        assert_close(hv.cpu().numpy(), om.hess_coord(xh, y, 0.7), "hess_coord (graph replay)")
        xd += 0.01 * torch.cos(torch.arange(xd.numel(), device="cuda", dtype=torch.float64))  # next iterate, same buffers


def test_device_objective_and_explicit_page_locking(oracle_cache):
    """iexa_obj_device (no host synchronisation: the objective lands in a device double, what a GPU-resident solver or
    the multi-GPU all-reduce consumes) and iexa_host_register / iexa_host_unregister (explicit page-locking of the
    caller's host vectors for the Ipopt-style path)."""
    import ctypes as C
    import torch
    core, om = _oracle(oracle_cache, "pandemic_50x4")
    m = ex.ExaModel(core, device=0)
    L = m.L
    x, y = eval_point(core, seed=5)
    xd = torch.from_numpy(x).cuda()
    fd = torch.zeros(1, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert L.iexa_obj_device(m.h, C.c_void_p(xd.data_ptr()), C.c_void_p(fd.data_ptr()), C.c_void_p(st)) == 0, L.iexa_last_error()
    assert_close(fd.item(), om.obj(x), "obj_device")
    # page-lock the caller's vectors, evaluate through the host path, unlock
    xh = x.copy(); jv = np.zeros(om.nnzj)
    for buf in (xh, jv):
        assert L.iexa_host_register(m.h, C.c_void_p(buf.ctypes.data), buf.nbytes) == 0, L.iexa_last_error()
    assert_close(ex.jac_coord_(m, xh, jv), om.jac_coord(x), "jac_coord through registered host buffers")
    assert L.iexa_host_register(m.h, C.c_void_p(xh.ctypes.data), xh.nbytes) == 0       # registering twice is a no-op
    for buf in (xh, jv):
        assert L.iexa_host_unregister(m.h, C.c_void_p(buf.ctypes.data)) == 0, L.iexa_last_error()
    assert L.iexa_host_unregister(m.h, C.c_void_p(jv.ctypes.data)) == 0                # so is unregistering twice
    assert_close(ex.jac_coord_(m, xh, jv), om.jac_coord(x), "jac_coord after unregistering")


def test_parameter_update_in_place(oracle_cache):
    """set_parameter! semantics (infiniteopt_backend.jl:511-548): θ changes, no plan rebuild."""
    import torch
    core = models.quadrotor(20, "oc")
    from oracle.oracle import OracleModel
    om = OracleModel(core)
    m = ex.ExaModel(core, device=0)
    x, _ = eval_point(core)
    xd = torch.from_numpy(x).cuda()
    f0 = ex.obj(m, xd)
    assert_close(f0, om.obj(x), "obj before")
    newvals = np.cos(np.linspace(0, 1, 39))
    from iexa_b200.core import Parameter
    d1 = Parameter(0, (39,))
    m.set_parameter(d1, newvals)
    om.set_parameter(0, newvals)
    assert np.array_equal(m.θ[:39], newvals)
    f1 = ex.obj(m, xd)
    assert abs(f1 - f0) > 1e-6
    assert_close(f1, om.obj(x), "obj after")
    g = torch.zeros(om.nvar, dtype=torch.float64, device="cuda")
    assert_close(ex.grad_(m, xd, g).cpu().numpy(), om.grad(x), "grad after")


def test_nan_propagates():
    import torch
    core = models.quadrotor(16, "fd")
    m = ex.ExaModel(core, device=0)
    x = torch.zeros(core.nvar, dtype=torch.float64, device="cuda")
    x[5] = float("nan")
    c = torch.zeros(core.ncon, dtype=torch.float64, device="cuda")
    ex.cons_(m, x, c)
    assert torch.isnan(c).any() and not torch.isnan(c).all()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_plans_tile_the_model_on_one_gpu(world, oracle_cache):
    """the CUDA path of every rank of a W-rank sharding, run one after the other on cuda:0: the local
    slices assembled through iexa_segments reproduce the unsharded oracle result"""
    import torch
    core = models.quadrotor(101, "oc")
    from oracle.oracle import OracleModel
    om = OracleModel(core)
    x, y = eval_point(core, seed=9)
    xd = torch.from_numpy(x).cuda()
    cg, jg, hg = np.zeros(om.ncon), np.zeros(om.nnzj), np.zeros(om.nnzh)
    f, g = 0.0, np.zeros(om.nvar)
    for rank in range(world):
        m = ex.ExaModel(core, device=0, rank=rank, world=world)
        segs = {}
        for which in range(3):
            arr = (ex.lib.Segment * 4096)()
            n = m.L.iexa_segments(m.h, which, arr, 4096)
            segs[which] = [(s.global_start, s.local_start, s.length) for s in arr[:n]]
        yl = np.zeros(max(m.loc_ncon, 1))
        for gs, ls, ln in segs[0]:
            yl[ls:ls + ln] = y[gs:gs + ln]
        c = torch.zeros(max(m.loc_ncon, 1), dtype=torch.float64, device="cuda")
        jv = torch.zeros(max(m.loc_nnzj, 1), dtype=torch.float64, device="cuda")
        hv = torch.zeros(max(m.loc_nnzh, 1), dtype=torch.float64, device="cuda")
        gd = torch.zeros(om.nvar, dtype=torch.float64, device="cuda")
        ex.cons_(m, xd, c); ex.jac_coord_(m, xd, jv); ex.hess_coord_(m, xd, torch.from_numpy(yl).cuda(), hv, 0.7)
        f += ex.obj(m, xd); g += ex.grad_(m, xd, gd).cpu().numpy()
        for (arrg, loc, which) in ((cg, c, 0), (jg, jv, 1), (hg, hv, 2)):
            l = loc.cpu().numpy()
            for gs, ls, ln in segs[which]:
                arrg[gs:gs + ln] = l[ls:ls + ln]
        # local structure: rows are global row numbers of the owned rows
        r = torch.zeros(max(m.loc_nnzj, 1), dtype=torch.int64, device="cuda"); cc = torch.zeros_like(r)
        ex.jac_structure_(m, r, cc)
        ro, co = om.jac_structure()
        rl, cl = r.cpu().numpy(), cc.cpu().numpy()
        for gs, ls, ln in segs[1]:
            assert (rl[ls:ls + ln] == ro[gs:gs + ln]).all() and (cl[ls:ls + ln] == co[gs:gs + ln]).all()
    assert_close(cg, om.cons(x), "cons"); assert_close(jg, om.jac_coord(x), "jac"); assert_close(hg, om.hess_coord(x, y, 0.7), "hess")
    assert abs(f - om.obj(x)) <= 1e-12 * abs(om.obj(x)) + 1e-14
    assert np.allclose(g, om.grad(x), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("device_side", [False, True])
def test_a_rank_keeps_its_slices_of_columns_and_theta_on_the_device(device_side):
    """world > 1 on the device: iterator columns are uploaded (or generated) as the slice the rank's supports visit, theta is a
    full-length virtual range with only the rank's granules backed (CUDA virtual-memory API) — sized so that path is taken
    (theta = 3 x 8*10^5 doubles = 19 MB, 10 granules, of which a rank of 4 maps <= 6).  Results: bit-identical to the unsharded
    model on the same GPU, slice by slice; set_parameter! on a block reaches the resident part and is ignored elsewhere."""
    import torch
    N, world = 400_000, 4
    core = models.quadrotor(N, "oc", device_side=device_side)
    full = ex.ExaModel(core, device=0)
    x, y = eval_point(core, seed=3)
    xd = torch.from_numpy(x).cuda()
    dev = torch.device("cuda:0")
    z = lambda n: torch.zeros(max(n, 1), dtype=torch.float64, device=dev)
    cF, jF, hF = z(full.meta.ncon), z(full.meta.nnzj), z(full.meta.nnzh)
    yd = torch.from_numpy(y).cuda()
    ex.cons_(full, xd, cF); ex.jac_coord_(full, xd, jF); ex.hess_coord_(full, xd, yd, hF, 0.7)
    b1 = ex.device_bytes(full)
    assert b1["columns"] == b1["columns_unsharded"] and b1["theta"] == b1["theta_unsharded"]
    for rank in (1, 3):
        m = ex.ExaModel(core, device=0, rank=rank, world=world)
        b = ex.device_bytes(m)
        # device_side: the parameter functions, too, are evaluated over the rank's share only
        assert b["columns"] <= b["columns_unsharded"] / world + 4096, b
        assert b["theta"] <= 0.75 * b["theta_unsharded"], b
        if device_side:
            T = core.npar // 3
            k0, k1 = (T * rank) // world, (T * (rank + 1)) // world
            th, ref = m.θ, full.θ
            for blk in range(3):       # the rank's own share of every block is there, bit-identical to the whole-block evaluation
                assert np.array_equal(th[blk * T + k0 + 2: blk * T + k1 - 2], ref[blk * T + k0 + 2: blk * T + k1 - 2]), (rank, blk)
        segs = {}
        for which in range(3):
            arr = (ex.lib.Segment * 4096)()
            n = m.L.iexa_segments(m.h, which, arr, 4096)
            segs[which] = [(s.global_start, s.local_start, s.length) for s in arr[:n]]
        yl = z(m.loc_ncon)
        for gs, ls, ln in segs[0]:
            yl[ls:ls + ln] = yd[gs:gs + ln]
        c, jv, hv = z(m.loc_ncon), z(m.loc_nnzj), z(m.loc_nnzh)
        ex.cons_(m, xd, c); ex.jac_coord_(m, xd, jv); ex.hess_coord_(m, xd, yl, hv, 0.7)
        for (ref, loc, which) in ((cF, c, 0), (jF, jv, 1), (hF, hv, 2)):
            for gs, ls, ln in segs[which]:
                assert torch.equal(loc[ls:ls + ln], ref[gs:gs + ln]), (rank, which)
        if not device_side:
            # update the first theta block (d1 of ESCAPE34/quadrotor.jl:16) on both models: same results again
            T = core.npar // 3
            new = np.linspace(-1.0, 1.0, T)
            for mm in (full, m):
                ex.lib.check(mm.L, mm.L.iexa_set_par(mm.h, 0, T, new.ctypes.data))
            ex.cons_(full, xd, cF); ex.cons_(m, xd, c)
            for gs, ls, ln in segs[0]:
                assert torch.equal(c[ls:ls + ln], cF[gs:gs + ln]), rank
            old = np.ascontiguousarray(core.theta_vec[:T])
            ex.lib.check(full.L, full.L.iexa_set_par(full.h, 0, T, old.ctypes.data))
            ex.cons_(full, xd, cF)
        del m


def test_opf_shape_classes_and_budget_fallback(oracle_cache, monkeypatch):
    """BASELINE configs[3] (ESCAPE34/opf.jl) through the general lowering: the embedded 3-bus case (fused
    groups), a 30-bus grid with 722 generators (its fused source exceeds the NVRTC budget -> generators
    of identical shape are canonicalised into class kernels with an instance axis), and the same grid
    with a tiny budget (-> AOT interpreter kernels, with the reason reported).  All match the oracle."""
    import torch
    from iexa_b200 import opf
    from iexa_b200.transform import exa_core
    from oracle.oracle import OracleModel
    for case, K, budget, expect_spec in ((None, 64, None, True), (opf.synthetic_grid(30), 9, None, True),
                                         (opf.synthetic_grid(30), 9, "1000", False)):
        if budget:
            monkeypatch.setenv("IEXA_CODEGEN_MAX_BYTES", budget)
        else:
            monkeypatch.delenv("IEXA_CODEGEN_MAX_BYTES", raising=False)
        core, _ = exa_core(opf.opf(case, num_supports=K))
        om = OracleModel(core)
        m = ex.ExaModel(core, device=0)
        note = m.L.iexa_engine_note(m.h).decode()
        assert (m.cmeta.n_kernels_specialised > 0) == expect_spec, note
        if not expect_spec:
            assert "budget" in note
        x, y = eval_point(core, seed=1)
        x = np.where(np.isfinite(x), x, 0.0)
        xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
        c = torch.zeros(om.ncon, dtype=torch.float64, device="cuda")
        jv = torch.zeros(om.nnzj, dtype=torch.float64, device="cuda")
        hv = torch.zeros(om.nnzh, dtype=torch.float64, device="cuda")
        g = torch.zeros(om.nvar, dtype=torch.float64, device="cuda")
        assert_close(ex.cons_(m, xd, c).cpu().numpy(), om.cons(x), "cons")
        assert_close(ex.jac_coord_(m, xd, jv).cpu().numpy(), om.jac_coord(x), "jac")
        assert_close(ex.hess_coord_(m, xd, yd, hv, 0.9).cpu().numpy(), om.hess_coord(x, y, 0.9), "hess")
        assert_close(ex.grad_(m, xd, g).cpu().numpy(), om.grad(x), "grad")
        assert_close(ex.obj(m, xd), om.obj(x), "obj")
        r = torch.zeros(om.nnzh, dtype=torch.int32, device="cuda"); cc = torch.zeros_like(r)
        ex.hess_structure_(m, r, cc)
        ro, co = om.hess_structure()
        assert (r.cpu().numpy() == ro).all() and (cc.cpu().numpy() == co).all()
        # fused products and the fused eval3 kernel through the same (fused-class / class / interpreter) paths
        rng = np.random.default_rng(2)
        v, w = rng.uniform(-1, 1, om.nvar), rng.uniform(-1, 1, om.ncon)
        vd, wd = torch.from_numpy(v).cuda(), torch.from_numpy(w).cuda()
        z = lambda n: torch.full((max(n, 1),), 7.0, dtype=torch.float64, device="cuda")
        assert_close(ex.jprod_(m, xd, vd, z(om.ncon)).cpu().numpy()[: om.ncon], om.jprod(x, v), "jprod")
        assert_close(ex.jtprod_(m, xd, wd, z(om.nvar)).cpu().numpy(), om.jtprod(x, w), "jtprod")
        assert_close(ex.hprod_(m, xd, yd, vd, z(om.nvar), 0.9).cpu().numpy(), om.hprod(x, y, v, 0.9), "hprod")
        c3, j3, h3 = z(om.ncon), z(om.nnzj), z(om.nnzh)
        ex.eval3_(m, xd, yd, c3, j3, h3, 0.9)
        assert_close(c3.cpu().numpy()[: om.ncon], om.cons(x), "eval3 cons")
        assert_close(j3.cpu().numpy()[: om.nnzj], om.jac_coord(x), "eval3 jac")
        assert_close(h3.cpu().numpy()[: om.nnzh], om.hess_coord(x, y, 0.9), "eval3 hess")
        if expect_spec:
            assert m.L.iexa_engine_note(m.h) == b"", m.L.iexa_engine_note(m.h)


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("name", ["quadrotor_oc_40", "pandemic_50x4", "ode_5x5"])
def test_callbacks_under_the_alternative_slot_order_policy(name, mode):
    """IEXA_OPT_SLOT_ORDER = right-to-left (the alternative of SURVEY App. A.2): structure bit-exact and values at the
    north-star tolerance against the oracle under the SAME policy — the slot order is data, the kernels are generated from it"""
    import torch
    from oracle.oracle import OracleModel
    core = CASES[name]()
    om = OracleModel(core, slot_order=1)
    m = ex.ExaModel(core, device=0, flags=MODES[mode], slot_order=1)
    x, y = eval_point(core, seed=8)
    dev = torch.device("cuda:0")
    xd, yd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    for fn, ref, n in ((ex.jac_structure_, om.jac_structure, om.nnzj), (ex.hess_structure_, om.hess_structure, om.nnzh)):
        r = torch.zeros(max(n, 1), dtype=torch.int64, device=dev); c = torch.zeros_like(r)
        fn(m, r, c)
        ro, co = ref()
        assert (r.cpu().numpy()[:n] == ro).all() and (c.cpu().numpy()[:n] == co).all()
    z = lambda n: torch.full((max(n, 1),), 7.0, dtype=torch.float64, device=dev)
    assert_close(ex.jac_coord_(m, xd, z(om.nnzj)).cpu().numpy()[: om.nnzj], om.jac_coord(x), "jac_coord")
    assert_close(ex.hess_coord_(m, xd, yd, z(om.nnzh), 0.7).cpu().numpy()[: om.nnzh], om.hess_coord(x, y, 0.7), "hess_coord")
    assert_close(ex.grad_(m, xd, z(om.nvar)).cpu().numpy(), om.grad(x), "grad")
    rng = np.random.default_rng(5)
    v = rng.uniform(-1, 1, om.nvar)
    assert_close(ex.hprod_(m, xd, yd, torch.from_numpy(v).to(dev), z(om.nvar), 0.7).cpu().numpy(), om.hprod(x, y, v, 0.7), "hprod")


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("name", ["quadrotor_oc_40", "pandemic_50x4", "ode_5x5", "opf_30bus_x9"])
def test_row_sorted_policy_writes_the_csr_value_array(name, mode):
    """IEXA_SLOT_ORDER_JAC_ROW_SORTED on the GPU: structure bit-exact and values at the north-star tolerance against the oracle
    under the same policy; device row pointers + jac_structure's columns + the untouched jac_coord! output are scipy's CSR of the
    same triplets; the Jacobian products agree with the oracle as well (their programs follow the permuted slots)."""
    import torch
    import scipy.sparse as sp
    from oracle.oracle import OracleModel
    if name == "opf_30bus_x9":      # 722 generators -> shape-class kernels: the permutation is part of the shape key
        from iexa_b200 import opf
        from iexa_b200.transform import exa_core
        core = exa_core(opf.opf(opf.synthetic_grid(30), num_supports=9))[0]
    else:
        core = CASES[name]()
    om = OracleModel(core, slot_order=2)
    m = ex.ExaModel(core, device=0, flags=MODES[mode], slot_order=2)
    assert ex.jac_is_csr(m) and om.L.orc_jac_is_csr(om.h) == 1
    x, y = eval_point(core, seed=8)
    x = np.where(np.isfinite(x), x, 0.0)
    dev = torch.device("cuda:0")
    xd, yd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    r = torch.zeros(om.nnzj, dtype=torch.int32, device=dev); c = torch.zeros_like(r)
    ex.jac_structure_(m, r, c)
    ro, co = om.jac_structure()
    assert (r.cpu().numpy() == ro).all() and (c.cpu().numpy() == co).all()
    z = lambda n: torch.full((max(n, 1),), 7.0, dtype=torch.float64, device=dev)
    jv = ex.jac_coord_(m, xd, z(om.nnzj)).cpu().numpy()
    assert_close(jv, om.jac_coord(x), "jac_coord")
    assert_close(ex.hess_coord_(m, xd, yd, z(om.nnzh), 0.7).cpu().numpy()[: om.nnzh], om.hess_coord(x, y, 0.7), "hess_coord")
    rp = torch.zeros(om.ncon + 1, dtype=torch.int32, device=dev)
    ex.jac_csr_rowptr_(m, rp)
    rp_h = np.zeros(om.ncon + 1, dtype=np.int64)
    ex.jac_csr_rowptr_(m, rp_h)
    assert (rp.cpu().numpy() == rp_h).all()
    mine = sp.csr_matrix((jv, co - 1, rp_h), shape=(om.ncon, om.nvar))
    ref = sp.coo_matrix((jv, (ro - 1, co - 1)), shape=(om.ncon, om.nvar)).tocsr()
    ref.sort_indices()
    assert (mine.indptr == ref.indptr).all() and (mine.indices == ref.indices).all() and np.array_equal(mine.data, ref.data)
    rng = np.random.default_rng(5)
    v, w = rng.uniform(-1, 1, om.nvar), rng.uniform(-1, 1, om.ncon)
    assert_close(ex.jprod_(m, xd, torch.from_numpy(v).to(dev), z(om.ncon)).cpu().numpy(), om.jprod(x, v), "jprod")
    assert_close(ex.jtprod_(m, xd, torch.from_numpy(w).to(dev), z(om.nvar)).cpu().numpy(), om.jtprod(x, w), "jtprod")
    cc, jj, hh = z(om.ncon), z(om.nnzj), z(om.nnzh)
    ex.eval3_(m, xd, yd, cc, jj, hh, 0.7)
    assert_close(jj.cpu().numpy(), jv, "eval3 Jacobian vs jac_coord! under the row-sorted policy")


@pytest.mark.parametrize("poison", [float("nan"), float("inf")])
@pytest.mark.parametrize("name", ["quadrotor_oc_40", "pandemic_50x4"])
def test_strict_ieee_mode_reproduces_the_oracles_nan_pattern_on_the_gpu(name, poison):
    """IEXA_OPT_STRICT_IEEE: structural zeros are multiplied at run time — the NaN pattern of cons / jac_coord / hess_coord on a
    poisoned x equals the oracle's (tests/test_nan_semantics.py does the same for every BASELINE config on the host executor)"""
    import torch
    from oracle.oracle import OracleModel
    core = CASES[name]()
    om = OracleModel(core)
    m = ex.ExaModel(core, device=0)   # IEEE-strict is the default
    x, y = eval_point(core, seed=2)
    rng = np.random.default_rng(1)
    x[rng.choice(core.nvar, size=max(1, core.nvar // 10), replace=False)] = poison
    dev = torch.device("cuda:0")
    xd, yd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    z = lambda n: torch.full((max(n, 1),), 7.0, dtype=torch.float64, device=dev)
    for got, ref in ((ex.cons_(m, xd, z(om.ncon)).cpu().numpy()[: om.ncon], om.cons(x)),
                     (ex.jac_coord_(m, xd, z(om.nnzj)).cpu().numpy()[: om.nnzj], om.jac_coord(x)),
                     (ex.hess_coord_(m, xd, yd, z(om.nnzh), 0.7).cpu().numpy()[: om.nnzh], om.hess_coord(x, y, 0.7))):
        assert np.array_equal(np.isnan(got), np.isnan(ref)), "NaN pattern differs from the oracle's"
        inf = np.isinf(ref)
        assert np.array_equal(got[inf], ref[inf]), "infinite entries differ"
        fin = np.isfinite(ref)
        assert_close(got[fin], ref[fin], "finite entries")


@pytest.mark.parametrize("name", ["quadrotor_oc_40", "quadrotor_oc_ragged", "pandemic_50x4", "farmer_1000", "ode_5x5"])
def test_eval3_fused_kernel_matches_the_oracle_and_the_three_callbacks(name, oracle_cache):
    """iexa_eval3: cons! + jac_coord! + hess_coord! at one (x, y) in ONE launch (one fused program per constraint group)"""
    import torch
    core, om = _oracle(oracle_cache, name)
    m = ex.ExaModel(core, device=0)
    x, y = eval_point(core, seed=6)
    dev = torch.device("cuda:0")
    xd, yd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    z = lambda n: torch.full((max(n, 1),), 7.0, dtype=torch.float64, device=dev)
    c, jv, hv = z(om.ncon), z(om.nnzj), z(om.nnzh)
    ex.eval3_(m, xd, yd, c, jv, hv, 0.7)
    assert_close(c.cpu().numpy()[: om.ncon], om.cons(x), "cons")
    assert_close(jv.cpu().numpy()[: om.nnzj], om.jac_coord(x), "jac_coord")
    assert_close(hv.cpu().numpy()[: om.nnzh], om.hess_coord(x, y, 0.7), "hess_coord")
    c2, j2, h2 = z(om.ncon), z(om.nnzj), z(om.nnzh)
    ex.cons_(m, xd, c2); ex.jac_coord_(m, xd, j2); ex.hess_coord_(m, xd, yd, h2, 0.7)
    assert_close(c.cpu().numpy(), c2.cpu().numpy(), "eval3 vs cons!")
    assert_close(jv.cpu().numpy(), j2.cpu().numpy(), "eval3 vs jac_coord!")
    assert_close(hv.cpu().numpy(), h2.cpu().numpy(), "eval3 vs hess_coord!")
    # host buffers: the three callbacks in sequence behind the same entry point
    ch, jh, hh = np.zeros(max(om.ncon, 1)), np.zeros(max(om.nnzj, 1)), np.zeros(max(om.nnzh, 1))
    ex.eval3_(m, x, y, ch, jh, hh, 0.7)
    assert_close(ch[: om.ncon], om.cons(x), "cons (host)")
    assert_close(hh[: om.nnzh], om.hess_coord(x, y, 0.7), "hess_coord (host)")
