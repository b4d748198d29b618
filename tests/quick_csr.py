"""scratch timing of the COO->CSR value permutation at config 3 (not a test)."""
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
for name, nnz, fn in (("hess", m.meta.nnzh, ex.hess_structure_), ("jac", m.meta.nnzj, ex.jac_structure_)):
    r = torch.zeros(nnz, dtype=torch.int32, device="cuda"); c = torch.zeros_like(r)
    fn(m, r, c)
    torch.cuda.synchronize()
    nrows = m.meta.nvar if name == "hess" else m.meta.ncon
    h = C.c_void_p()
    t0 = time.time()
    rc = L.iexa_csr_create(C.byref(h), nrows, m.meta.nvar, nnz, r.data_ptr(), c.data_ptr(), 4, 1, 0)
    assert rc == 0, L.iexa_last_error()
    t1 = time.time()
    cn = L.iexa_csr_nnz(h)
    vals = torch.rand(nnz, dtype=torch.float64, device="cuda"); out = torch.zeros(cn, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        L.iexa_csr_apply(h, vals.data_ptr(), out.data_ptr(), 1, st)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        L.iexa_csr_apply(h, vals.data_ptr(), out.data_ptr(), 1, st)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    byt = 8 * nnz + 4 * nnz + 8 * cn + 8 * cn
    print(f"{name}: nnz={nnz} csr_nnz={cn} setup {t1-t0:.2f}s apply {ms:.3f} ms  {byt/ms/1e6:.0f} GB/s ({byt/ms/1e6/6552:.2f} of 6552)")
    # spot check against torch
    ref = torch.zeros(cn, dtype=torch.float64, device="cuda")
    L.iexa_csr_destroy(h)
    del r, c, vals, out
    torch.cuda.empty_cache()
