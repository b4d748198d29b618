"""scratch timing of the callbacks on any config model (not a test).

usage: python tests/quick_bench.py <model> <size> [flags]
  model: quad | quadfd | pandemic | pandemic128 | opf | opf30 | opf118 | farmer
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import iexa_b200 as ex
from iexa_b200 import models, opf
from iexa_b200.transform import exa_core

name = sys.argv[1] if len(sys.argv) > 1 else "quad"
if name.isdigit():  # backwards compatible: quick_bench.py <N> [flags]
    name, N, flags = "quad", int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 0
else:
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t0 = time.time()
core = {"quad": lambda: models.quadrotor(N, "oc"), "quadfd": lambda: models.quadrotor(N, "fd"),
        "pandemic": lambda: models.pandemic(N, 4), "pandemic128": lambda: models.pandemic(N, 128),
        "opf": lambda: exa_core(opf.opf(None, num_supports=N))[0],
        "opf30": lambda: exa_core(opf.opf(opf.synthetic_grid(30), num_supports=N))[0],
        "opf118": lambda: exa_core(opf.opf(opf.synthetic_grid(118), num_supports=N))[0],
        "farmer": lambda: models.farmer(N)}[name]()
t1 = time.time()
WORLD = int(os.environ.get("IEXA_QB_WORLD", "1"))   # time ONE shard of a world-W sharding on this GPU (rank W/2)
m = ex.ExaModel(core, device=0, flags=flags, rank=WORLD // 2, world=WORLD)
t2 = time.time()
print(f"{name} N={N} build core {t1-t0:.2f}s plan+finalize {t2-t1:.2f}s nvar={m.meta.nvar} ncon={m.meta.ncon} "
      f"nnzj={m.meta.nnzj} nnzh={m.meta.nnzh} gens={m.cmeta.nobj_gen + m.cmeta.ncon_gen} spec={m.cmeta.n_kernels_specialised} "
      f"note={m.L.iexa_engine_note(m.h).decode()[:80]!r}")
rng = np.random.default_rng(0)
x0 = np.where(np.isfinite(core.x0_vec), core.x0_vec, 0.0)
x = torch.from_numpy(x0 + 0.1 * rng.uniform(-1, 1, core.nvar)).cuda()
y = torch.from_numpy(rng.uniform(-1, 1, max(m.loc_ncon, 1))).cuda()
c = torch.zeros(max(m.loc_ncon, 1), dtype=torch.float64, device="cuda")
jv = torch.zeros(max(m.loc_nnzj, 1), dtype=torch.float64, device="cuda")
hv = torch.zeros(max(m.loc_nnzh, 1), dtype=torch.float64, device="cuda")
g = torch.zeros(m.meta.nvar, dtype=torch.float64, device="cuda")


REPS = int(os.environ.get("IEXA_QB_REPS", "50"))   # profiling runs (ncu --set full): IEXA_QB_REPS=1


def timeit(fn, n=None):
    n = REPS if n is None else n
    for _ in range(3 if REPS > 2 else 1):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


tb = time.time(); B = [ex.algorithmic_bytes(m, w) for w in range(5)]; print("bytes", B, f"({time.time()-tb:.1f}s)")
tot = 0.0
from iexa_b200.model import bind  # pre-bound calls: one foreign call each, like a Julia ccall (the un-bound Python wrappers cost ~10 us)
for nm, fn, w in (("cons", bind(m, "cons", x, c), 2), ("jac", bind(m, "jac_coord", x, jv), 3),
                  ("hess", bind(m, "hess_coord", x, hv, y, 1.0), 4), ("grad", bind(m, "grad", x, g), 1),
                  ("obj", lambda: ex.obj(m, x), 0)):
    ms = timeit(fn)
    if w >= 2:
        tot += ms
    print(f"{nm}: {ms:.4f} ms  {B[w]/ms/1e6:.1f} GB/s  frac_of_6552={B[w]/ms/1e6/6552:.3f}")
_fc, _fj, _fh = bind(m, "cons", x, c), bind(m, "jac_coord", x, jv), bind(m, "hess_coord", x, hv, y, 1.0)
ms3 = timeit(lambda: (_fc(), _fj(), _fh()))
print(f"step (cons, jac, hess back to back): {ms3:.4f} ms -> {1e3/ms3:.0f} evals/s (sum of the three timed alone: {tot:.4f} ms)")
if os.environ.get("IEXA_EVAL3", "1") != "0":   # the fused cons + jac + hess kernel (compiled on first use)
    f3 = bind(m, "eval3", x, (c, jv, hv), y, 1.0)
    tp = time.time(); f3(); torch.cuda.synchronize(); tb3 = time.time() - tp
    ms = timeit(f3)
    print(f"eval3 (one fused launch): {ms:.4f} ms -> {1e3/ms:.0f} evals/s, {sum(B[2:])/ms/1e6:.0f} GB/s ({sum(B[2:])/ms/1e6/6552:.3f} of 6552)  [three launches: {tot:.4f} ms; module build {tb3:.2f}s]")
if os.environ.get("IEXA_PROD", "1") != "0":  # matrix-free products (fused kernels, second NVRTC module compiled on first use)
    v = torch.from_numpy(rng.uniform(-1, 1, core.nvar)).cuda()
    w = torch.from_numpy(rng.uniform(-1, 1, max(m.loc_ncon, 1))).cuda()
    Jv = torch.zeros(max(m.loc_ncon, 1), dtype=torch.float64, device="cuda")
    Jtw = torch.zeros(m.meta.nvar, dtype=torch.float64, device="cuda"); Hv = torch.zeros_like(Jtw)
    tp = time.time(); ex.jprod_(m, x, v, Jv); torch.cuda.synchronize(); print(f"product module build {time.time()-tp:.2f}s note={m.L.iexa_engine_note(m.h).decode()[:80]!r}")
    Bp = [ex.algorithmic_bytes(m, w_) for w_ in (5, 6, 7)]
    for nm, fn, b in (("jprod", bind(m, "jprod", x, Jv, v=v), Bp[0]), ("jtprod", bind(m, "jtprod", x, Jtw, v=w), Bp[1]),
                      ("hprod", bind(m, "hprod", x, Hv, y, 1.0, v=v), Bp[2])):
        ms = timeit(fn)
        print(f"{nm}: {ms:.4f} ms  {b/ms/1e6:.1f} GB/s  frac_of_6552={b/ms/1e6/6552:.3f}  bytes={b}  launches={ex.launches_per_call(m, {'jprod': 5, 'jtprod': 6, 'hprod': 7}[nm])}")
if os.environ.get("IEXA_GRAPH"):  # the three callbacks captured into one CUDA graph and replayed
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fc, fj, fh = bind(m, "cons", x, c), bind(m, "jac_coord", x, jv), bind(m, "hess_coord", x, hv, y, 1.0)
        fc(); fj(); fh()
    st.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        fc(); fj(); fh()
    print(f"cons+jac+hess as ONE CUDA graph replay: {timeit(gr.replay):.4f} ms")
print(f"cons+jac+hess: {tot:.4f} ms -> {1e3/tot:.0f} evals/s, {sum(B[2:])/tot/1e6:.0f} GB/s ({sum(B[2:])/tot/1e6/6552:.3f} of 6552)")
