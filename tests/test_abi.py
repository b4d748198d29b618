"""The C-ABI library loads WITHOUT a GPU and exports every symbol include/iexa.h declares; plan
construction and host-side queries work; every evaluation entry point fails loudly (no CPU path)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import iexa_b200 as ex
from iexa_b200 import models
from conftest import ROOT, has_gpu


def _declared():
    src = open(os.path.join(ROOT, "include", "iexa.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iexa_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_all_exported():
    L = ex.lib.load()
    names = _declared()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(ex.lib.SYMBOLS) == names, "lib.py's binding list must match the header"
    assert L.iexa_version() == 100


def test_struct_layouts_match_header():
    assert ex.expr.NODE_DTYPE.itemsize == 24 and ex.expr.INDEX_DTYPE.itemsize == 64
    assert C.sizeof(ex.lib.Meta) == 11 * 8 + 6 * 4
    assert C.sizeof(ex.lib.Segment) == 24


@pytest.mark.skipif(has_gpu(), reason="exercises the GPU-less failure path")
def test_product_library_has_no_cpu_fallback():
    core = models.ode_5x5()
    with pytest.raises(ex.lib.IexaError) as e:
        ex.ExaModel(core, device=0)  # default flags: needs a device
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE)
    assert (m.meta.nvar, m.meta.ncon, m.meta.nnzj, m.meta.nnzh) == (51, 70, 185, 50)
    assert np.array_equal(m.meta.x0, core.x0_vec) and m.meta.lcon[25] == -np.inf and m.meta.ucon[25] == 42.0
    for call in (lambda: ex.obj(m, core.x0_vec), lambda: ex.cons_(m, core.x0_vec, np.zeros(70)),
                 lambda: ex.jac_coord_(m, core.x0_vec, np.zeros(185)),
                 lambda: ex.jac_structure_(m, np.zeros(185, dtype=np.int64), np.zeros(185, dtype=np.int64))):
        with pytest.raises(ex.lib.IexaError):
            call()
    h = C.c_void_p()
    rc = ex.lib.load().iexa_csr_create(C.byref(h), 3, 3, 0, None, None, 4, 0, 0)
    assert rc == 3  # IEXA_ERR_CUDA


def test_state_errors():
    L = ex.lib.load()
    h = C.c_void_p()
    assert L.iexa_plan_create(C.byref(h), 1) == 0
    m = ex.lib.Meta()
    assert L.iexa_get_meta(h, C.byref(m)) == 0 and m.nvar == 0
    x = np.zeros(1)
    assert L.iexa_cons(h, x.ctypes.data, x.ctypes.data, 0, None) == 2  # not finalized -> IEXA_ERR_STATE
    assert b"not finalized" in L.iexa_last_error()
    assert L.iexa_finalize(h, 0, 3, 2, ex.lib.IEXA_F_NO_DEVICE) == 1      # rank >= world
    assert L.iexa_plan_destroy(h) == 0
    assert L.iexa_get_meta(None, C.byref(m)) == 1


def test_round2_queries_state_and_argument_errors():
    """iexa_jac_is_csr / iexa_jac_csr_rowptr / iexa_device_bytes: argument and state errors are codes + messages, never a crash;
    the structure query works on a host-only plan, the device-memory query needs an engine"""
    from iexa_b200 import models
    L = ex.lib.load()
    h = C.c_void_p()
    assert L.iexa_plan_create(C.byref(h), 1) == 0
    flag = C.c_int32(7)
    assert L.iexa_jac_is_csr(h, C.byref(flag)) == 2 and b"not finalized" in L.iexa_last_error()
    assert L.iexa_jac_is_csr(h, None) == 1
    assert L.iexa_jac_is_csr(None, C.byref(flag)) == 1
    rp = np.zeros(4, dtype=np.int64)
    assert L.iexa_jac_csr_rowptr(h, rp.ctypes.data, 8, 0, None) == 2          # not finalized
    out6 = (C.c_int64 * 6)()
    assert L.iexa_device_bytes(h, out6) == 2
    assert L.iexa_device_bytes(h, None) == 1
    assert L.iexa_plan_destroy(h) == 0
    # a finalized host-only plan under the default policy: not CSR -> STATE with a message that names the policy
    m = ex.ExaModel(models.farmer(4), flags=ex.lib.IEXA_F_NO_DEVICE)
    assert not ex.jac_is_csr(m)
    rp = np.zeros(m.meta.ncon + 1, dtype=np.int64)
    assert L.iexa_jac_csr_rowptr(m.h, rp.ctypes.data, 8, 0, None) == 2 and b"IEXA_SLOT_ORDER_JAC_ROW_SORTED" in L.iexa_last_error()
    assert L.iexa_device_bytes(m.h, out6) == 2 and b"no device engine" in L.iexa_last_error()
    # under the row-sorted policy: host buffers are served from the plan; a bad index width falls through to the engine path
    m2 = ex.ExaModel(models.farmer(4), flags=ex.lib.IEXA_F_NO_DEVICE, slot_order=2)
    assert ex.jac_is_csr(m2)
    assert L.iexa_jac_csr_rowptr(m2.h, rp.ctypes.data, 8, 0, None) == 0 and rp[0] == 0 and rp[-1] == m2.meta.nnzj
    assert L.iexa_jac_csr_rowptr(m2.h, rp.ctypes.data, 3, 0, None) != 0
    assert L.iexa_jac_csr_rowptr(m2.h, None, 8, 0, None) != 0
