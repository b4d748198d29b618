"""CPU tests of the host logic: the plan compiler (symbolic sparsity, AD register programs, generator
fusion, layout, sharding segments, byte accounting, NVRTC source generation) checked against the
oracle through the test-only host executor (tests/hostcheck).  No GPU, no product kernels."""
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
    "pandemic_100x128": lambda: models.pandemic(100, 128),   # the largest pandemic case of the reference's study grid (ESCAPE34/run_cases_gpu.jl:100)
    "farmer": lambda: models.farmer(11),
}


def _hc(L, m, fn, which, n, x, y=None, sig=1.0):
    out = np.zeros(max(n, 1))
    args = [m.h, which, x.ctypes.data, None if y is None else y.ctypes.data, sig, out.ctypes.data]
    if fn == "hostcheck_eval_groups":
        args.append(C.byref(C.c_int32()))
    assert getattr(L, fn)(*args) == 0
    return out[:n]


@pytest.mark.parametrize("fn", ["hostcheck_eval", "hostcheck_eval_groups"])
@pytest.mark.parametrize("name", list(CASES))
def test_programs_match_oracle(name, fn, hostcheck_lib):
    from oracle.oracle import OracleModel
    L = hostcheck_lib
    core = CASES[name]()
    om = OracleModel(core)
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    assert (m.meta.nvar, m.meta.ncon, m.meta.nnzj, m.meta.nnzh) == (om.nvar, om.ncon, om.nnzj, om.nnzh)
    x, y = eval_point(core)
    assert_close(_hc(L, m, fn, 0, 1, x)[0], om.obj(x), "obj")
    assert_close(_hc(L, m, fn, 1, om.nvar, x), om.grad(x), "grad")
    assert_close(_hc(L, m, fn, 2, om.ncon, x), om.cons(x), "cons")
    assert_close(_hc(L, m, fn, 3, om.nnzj, x), om.jac_coord(x), "jac")
    assert_close(_hc(L, m, fn, 4, om.nnzh, x, y, 0.7), om.hess_coord(x, y, 0.7), "hess")
    assert_close(_hc(L, m, fn, 4, om.nnzh, x, None, 1.3), om.hess_coord(x, None, 1.3), "hess (objective only)")


@pytest.mark.parametrize("name", list(CASES))
def test_structure_bit_exact(name, hostcheck_lib):
    from oracle.oracle import OracleModel
    L = hostcheck_lib
    core = CASES[name]()
    om = OracleModel(core)
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    for which, (ro, co) in ((0, om.jac_structure()), (1, om.hess_structure())):
        n = len(ro)
        r = np.zeros(max(n, 1), dtype=np.int64); c = np.zeros_like(r)
        assert L.hostcheck_structure(m.h, which, r.ctypes.data, c.ctypes.data) == 0
        assert (r[:n] == ro).all() and (c[:n] == co).all()
        if which == 1 and n:
            assert (r[:n] >= c[:n]).all(), "Hessian must be lower-triangular"


@pytest.mark.parametrize("name", list(CASES))
def test_generated_cuda_compiles_for_sm100a(name, hostcheck_lib):
    """NVRTC cross-compiles the specialised kernels without a GPU."""
    L = hostcheck_lib
    m = ex.ExaModel(CASES[name](), flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    nb = C.c_int64()
    rc = L.iexa_debug_codegen_compile(m.h, C.byref(nb))
    assert rc == 0, L.iexa_last_error().decode()
    assert nb.value > 1000
    n = L.iexa_debug_codegen_source(m.h, None, 0)
    buf = C.create_string_buffer(n + 1)
    L.iexa_debug_codegen_source(m.h, buf, n + 1)
    src = buf.value.decode()
    assert "iexa_cb_cons" in src and "__constant__ long long CI[" in src


def test_fusion_groups_quadrotor(hostcheck_lib):
    """the 9 ODE rows share one iterator -> one fused group; so do the 9 collocation rows"""
    L = hostcheck_lib
    m = ex.ExaModel(models.quadrotor(9, "oc"), flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    ng = C.c_int32()
    x = np.zeros(m.meta.nvar); out = np.zeros(1)
    L.hostcheck_eval_groups(m.h, 0, x.ctypes.data, None, 1.0, out.ctypes.data, C.byref(ng))
    # objective, K=1 initial conditions, ODE rows, OC rows, control collocation rows
    assert ng.value == 5


def test_evaluation_without_gpu_fails_loudly(hostcheck_lib):
    core = models.ode_5x5()
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=hostcheck_lib)
    with pytest.raises(ex.lib.IexaError) as e:
        ex.obj(m, core.x0_vec)
    assert "no CPU fallback" in str(e.value)


def test_sharding_segments_partition_the_model(hostcheck_lib):
    """world=4: the local segments of the 4 ranks tile rows / Jacobian / Hessian slots exactly once"""
    L = hostcheck_lib
    core = models.quadrotor(21, "oc")
    full = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    totals = {0: full.meta.ncon, 1: full.meta.nnzj, 2: full.meta.nnzh}
    cover = {w: np.zeros(t, dtype=np.int32) for w, t in totals.items()}
    for rank in range(4):
        m = ex.ExaModel(core, rank=rank, world=4, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
        for which, loc in ((0, m.loc_ncon), (1, m.loc_nnzj), (2, m.loc_nnzh)):
            segs = (ex.lib.Segment * 1024)()
            n = L.iexa_segments(m.h, which, segs, 1024)
            pos = 0
            for s in segs[:n]:
                assert s.local_start == pos
                pos += s.length
                cover[which][s.global_start:s.global_start + s.length] += 1
            assert pos == loc
    for w in cover:
        assert (cover[w] == 1).all()


def test_algorithmic_bytes_positive_and_consistent(hostcheck_lib):
    L = hostcheck_lib
    m = ex.ExaModel(models.quadrotor(50, "oc"), flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    b = [L.iexa_algorithmic_bytes(m.h, w) for w in range(5)]
    assert all(v > 0 for v in b)
    # outputs alone are a lower bound
    assert b[2] >= 8 * m.meta.ncon and b[3] >= 8 * m.meta.nnzj and b[4] >= 8 * m.meta.nnzh
    # ... and inputs cannot exceed x + theta + every column entry once
    assert b[3] <= 8 * (m.meta.nnzj + m.meta.nvar + m.cmeta.npar) + 8 * 6 * 99


def test_bad_tapes_are_rejected(hostcheck_lib):
    L = hostcheck_lib
    h = C.c_void_p()
    assert L.iexa_plan_create(C.byref(h), 1) == 0
    nodes = np.zeros(1, dtype=ex.expr.NODE_DTYPE)
    nodes[0] = (999, 0, 0, 0, 0.0)
    off = C.c_int64()
    rc = L.iexa_add_con(h, nodes.ctypes.data, 1, None, 0, 0, 0.0, 0.0, C.byref(off))
    assert rc != 0 and b"unsupported operator" in L.iexa_last_error()
    nodes[0] = (ex.OP["VAR"], 3, 0, 0, 0.0)  # index id out of range
    rc = L.iexa_add_con(h, nodes.ctypes.data, 1, None, 0, 0, 0.0, 0.0, C.byref(off))
    assert rc != 0
    L.iexa_plan_destroy(h)


def test_shape_classes_match_oracle(hostcheck_lib):
    """class mode: generators of identical shape share one parametrised program (plan.hpp::build_groups)"""
    from oracle.oracle import OracleModel
    from iexa_b200 import opf
    from iexa_b200.transform import exa_core
    L = hostcheck_lib
    core, _ = exa_core(opf.opf(opf.synthetic_grid(12), num_supports=4))
    om = OracleModel(core)
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    n_fused = L.iexa_debug_codegen_source(m.h, None, 0)
    assert L.iexa_debug_set_class_mode(m.h, 1) == 0
    n_class = L.iexa_debug_codegen_source(m.h, None, 0)
    assert n_class < n_fused / 3, (n_class, n_fused)
    x, y = eval_point(core)
    x = np.where(np.isfinite(x), x, 0.0)
    for which, ref in ((0, [om.obj(x)]), (1, om.grad(x)), (2, om.cons(x)), (3, om.jac_coord(x)), (4, om.hess_coord(x, y, 0.7))):
        assert_close(_hc(L, m, "hostcheck_eval_groups", which, len(ref), x, y, 0.7 if which == 4 else 1.0), np.asarray(ref), str(which))
    nb = C.c_int64()
    assert L.iexa_debug_codegen_compile(m.h, C.byref(nb)) == 0, L.iexa_last_error().decode()


def test_cubin_disk_cache_roundtrip_and_corruption(tmp_path):
    """compiled images are kept on disk keyed by the generated source: a second PROCESS loads instead of compiling; a
    truncated / foreign file under the same name is ignored (and replaced)"""
    import os
    import subprocess
    import sys
    code = ("import sys, ctypes as C; sys.path.insert(0, %r); import iexa_b200 as ex; from iexa_b200 import models\n"
            "m = ex.ExaModel(models.quadrotor(20, 'oc'), flags=ex.lib.IEXA_F_NO_DEVICE)\n"
            "nb = C.c_int64(0); assert m.L.iexa_debug_codegen_compile(m.h, C.byref(nb)) == 0, m.L.iexa_last_error()\n"
            "a, b = C.c_int32(), C.c_int32(); m.L.iexa_debug_cache_stats(C.byref(a), C.byref(b)); print(a.value, b.value, nb.value)\n"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, IEXA_CACHE_DIR=str(tmp_path))
    env.pop("IEXA_DUMP_DIR", None)
    run = lambda: subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split()
    first = run()
    assert first[:2] == ["1", "0"]
    files = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(files) == 1
    second = run()
    assert second[:2] == ["0", "1"] and second[2] == first[2]
    path = os.path.join(tmp_path, files[0])
    with open(path, "r+b") as f:   # corrupt the header: the entry must be ignored, recompiled and rewritten
        f.seek(8); f.write(b"\xff" * 8)
    third = run()
    assert third[:2] == ["1", "0"]
    assert run()[:2] == ["0", "1"]
    off = subprocess.run([sys.executable, "-c", code], env=dict(env, IEXA_CACHE_DIR="off"), capture_output=True, text=True, check=True).stdout.split()
    assert off[:2] == ["1", "0"]
