"""scratch timing of the COO->CSR value permutation at config 3 (not a test): plain work order (CSR entries by the COO
position of their first source) and locality-keyed work order (iexa_coo_locality)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import iexa_b200 as ex
from iexa_b200 import models

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
core = models.quadrotor(N, "oc")
m = ex.ExaModel(core, device=0)
L = ex.lib.load()
st = torch.cuda.current_stream().cuda_stream
for name, which, nnz, fn in (("hess", 1, m.meta.nnzh, ex.hess_structure_), ("jac", 0, m.meta.nnzj, ex.jac_structure_)):
    r = torch.zeros(nnz, dtype=torch.int32, device="cuda"); c = torch.zeros_like(r)
    fn(m, r, c)
    keys = torch.zeros(nnz, dtype=torch.int32, device="cuda")
    assert L.iexa_coo_locality(m.h, which, keys.data_ptr(), 1, st) == 0, L.iexa_last_error()
    torch.cuda.synchronize()
    nrows = m.meta.nvar if name == "hess" else m.meta.ncon
    vals = torch.rand(nnz, dtype=torch.float64, device="cuda")
    ref = None
    for label, kp in (("plain", None), ("keyed", keys.data_ptr())):
        h = C.c_void_p()
        t0 = time.time()
        rc = L.iexa_csr_create_keyed(C.byref(h), nrows, m.meta.nvar, nnz, r.data_ptr(), c.data_ptr(), 4, kp, 1, 0)
        assert rc == 0, L.iexa_last_error()
        t1 = time.time()
        cn = L.iexa_csr_nnz(h)
        out = torch.zeros(cn, dtype=torch.float64, device="cuda")
        for _ in range(3):
            L.iexa_csr_apply(h, vals.data_ptr(), out.data_ptr(), 1, st)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            L.iexa_csr_apply(h, vals.data_ptr(), out.data_ptr(), 1, st)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        # algorithmic bytes: every COO value and every CSR value once, plus the 4-byte source / destination maps
        byt = 8 * nnz + 8 * cn + 4 * cn + (0 if (cn == nnz and kp is None) else 4 * nnz + 4 * cn)
        print(f"{name} {label}: nnz={nnz} csr_nnz={cn} setup {t1-t0:.2f}s apply {ms:.3f} ms  {byt/ms/1e6:.0f} GB/s ({byt/ms/1e6/6552:.2f} of 6552)")
        if ref is None:
            ref = out.clone()
        else:
            assert torch.equal(ref, out), "keyed and plain work orders disagree"
        L.iexa_csr_destroy(h)
        del out
    del r, c, vals, keys, ref
    torch.cuda.empty_cache()
