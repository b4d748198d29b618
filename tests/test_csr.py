"""COO -> CSR value permutation against scipy on the KKT-style pattern of a
transcribed model: duplicates summed, pattern sorted, values bit-reproducible."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

import iexa_b200 as ex
from iexa_b200 import models

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("idx_dtype", [np.int32, np.int64])
def test_csr_matches_scipy(idx_dtype):
    import torch
    core = models.quadrotor(50, "oc")
    m = ex.ExaModel(core, device=0)
    nh = m.meta.nnzh
    r = np.zeros(nh, dtype=idx_dtype); c = np.zeros(nh, dtype=idx_dtype)
    ex.hess_structure_(m, r, c)
    L = ex.lib.load()
    h = C.c_void_p()
    n = m.meta.nvar
    assert L.iexa_csr_create(C.byref(h), n, n, nh, r.ctypes.data, c.ctypes.data, r.dtype.itemsize, 0, 0) == 0, L.iexa_last_error()
    ref = sp.coo_matrix((np.ones(nh), (r - 1, c - 1)), shape=(n, n)).tocsr()
    ref.sum_duplicates(); ref.sort_indices()
    nnz = L.iexa_csr_nnz(h)
    assert nnz == ref.nnz
    rowptr = np.zeros(n + 1, dtype=np.int32); colind = np.zeros(nnz, dtype=np.int32)
    assert L.iexa_csr_pattern(h, rowptr.ctypes.data, colind.ctypes.data, 0) == 0
    assert (rowptr == ref.indptr).all() and (colind == ref.indices).all()
    vals = np.random.default_rng(0).uniform(-1, 1, nh)
    refv = sp.coo_matrix((vals, (r - 1, c - 1)), shape=(n, n)).tocsr()
    refv.sum_duplicates(); refv.sort_indices()
    out = np.zeros(nnz)
    assert L.iexa_csr_apply(h, vals.ctypes.data, out.ctypes.data, 0, None) == 0
    assert np.allclose(out, refv.data, rtol=1e-15, atol=1e-15)
    # device buffers, twice: bit-reproducible (no atomics)
    vd = torch.from_numpy(vals).cuda(); o1 = torch.zeros(nnz, dtype=torch.float64, device="cuda"); o2 = torch.zeros_like(o1)
    st = torch.cuda.current_stream().cuda_stream
    assert L.iexa_csr_apply(h, vd.data_ptr(), o1.data_ptr(), 1, st) == 0
    assert L.iexa_csr_apply(h, vd.data_ptr(), o2.data_ptr(), 1, st) == 0
    torch.cuda.synchronize()
    assert torch.equal(o1, o2) and np.array_equal(o1.cpu().numpy(), out)
    L.iexa_csr_destroy(h)
    # locality-keyed work order (iexa_coo_locality): same pattern, bit-identical values
    rd, cd = torch.from_numpy(r.astype(np.int32)).cuda(), torch.from_numpy(c.astype(np.int32)).cuda()
    keys = torch.zeros(nh, dtype=torch.int32, device="cuda")
    assert L.iexa_coo_locality(m.h, 1, keys.data_ptr(), 1, st) == 0, L.iexa_last_error()
    torch.cuda.synchronize()
    assert int(keys.min()) >= 0 and int(keys.max()) < (1 << 20)
    hk = C.c_void_p()
    assert L.iexa_csr_create_keyed(C.byref(hk), n, n, nh, rd.data_ptr(), cd.data_ptr(), 4, keys.data_ptr(), 1, 0) == 0, L.iexa_last_error()
    assert L.iexa_csr_nnz(hk) == nnz
    rp2 = np.zeros(n + 1, dtype=np.int32); ci2 = np.zeros(nnz, dtype=np.int32)
    assert L.iexa_csr_pattern(hk, rp2.ctypes.data, ci2.ctypes.data, 0) == 0
    assert (rp2 == rowptr).all() and (ci2 == colind).all()
    o3 = torch.zeros(nnz, dtype=torch.float64, device="cuda")
    assert L.iexa_csr_apply(hk, vd.data_ptr(), o3.data_ptr(), 1, st) == 0
    torch.cuda.synchronize()
    assert torch.equal(o3, o1)
    L.iexa_csr_destroy(hk)


def test_csr_empty_pattern():
    L = ex.lib.load()
    h = C.c_void_p()
    assert L.iexa_csr_create(C.byref(h), 4, 4, 0, None, None, 4, 0, 0) == 0
    assert L.iexa_csr_nnz(h) == 0
    rowptr = np.ones(5, dtype=np.int32); colind = np.zeros(1, dtype=np.int32)
    assert L.iexa_csr_pattern(h, rowptr.ctypes.data, colind.ctypes.data, 0) == 0
    assert (rowptr == 0).all()
    L.iexa_csr_destroy(h)
