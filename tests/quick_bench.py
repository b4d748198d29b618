"""scratch timing of the three callbacks at large N (not a test)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iexa_b200 as ex
from iexa_b200 import models

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t0 = time.time(); core = models.quadrotor(N, "oc"); t1 = time.time()
m = ex.ExaModel(core, device=0, flags=flags); t2 = time.time()
print(f"N={N} build core {t1-t0:.2f}s plan+finalize {t2-t1:.2f}s nvar={m.meta.nvar} ncon={m.meta.ncon} nnzj={m.meta.nnzj} nnzh={m.meta.nnzh} spec={m.cmeta.n_kernels_specialised}")
rng = np.random.default_rng(0)
x = torch.from_numpy(core.x0_vec + 0.1 * rng.uniform(-1, 1, core.nvar)).cuda()
y = torch.from_numpy(rng.uniform(-1, 1, core.ncon)).cuda()
c = torch.zeros(m.meta.ncon, dtype=torch.float64, device="cuda")
jv = torch.zeros(m.meta.nnzj, dtype=torch.float64, device="cuda")
hv = torch.zeros(m.meta.nnzh, dtype=torch.float64, device="cuda")
g = torch.zeros(m.meta.nvar, dtype=torch.float64, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
tb = time.time(); B = [ex.algorithmic_bytes(m, w) for w in range(5)]; print("bytes", B, f"({time.time()-tb:.1f}s)")
for name, fn, w in (("cons", lambda: ex.cons_(m, x, c), 2), ("jac", lambda: ex.jac_coord_(m, x, jv), 3),
                    ("hess", lambda: ex.hess_coord_(m, x, y, hv), 4), ("grad", lambda: ex.grad_(m, x, g), 1),
                    ("obj", lambda: ex.obj(m, x), 0)):
    ms = timeit(fn)
    print(f"{name}: {ms:.3f} ms  {B[w]/ms/1e6:.1f} GB/s  frac_of_6552={B[w]/ms/1e6/6552:.3f}")
