"""host-side cost of one callback (scratch): tiny model, so the GPU is never the bottleneck"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iexa_b200 as ex
from iexa_b200 import models
from iexa_b200.model import bind
core = models.farmer(1000)
m = ex.ExaModel(core, device=0)
x = torch.from_numpy(core.x0_vec + 1.0).cuda()
c = torch.zeros(m.meta.ncon, dtype=torch.float64, device="cuda")
f = bind(m, "cons", x, c)
L = m.L
def loop(fn, n):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter() - t
    torch.cuda.synchronize(); t2 = time.perf_counter() - t
    return 1e6 * t1 / n, 1e6 * t2 / n
for _ in range(3): loop(f, 200)
print("iexa_cons via BoundCall: issue %.2f us/call, issue+drain %.2f us/call" % loop(f, 5000))
print("ctypes no-op (iexa_last_error): %.2f us/call" % loop(lambda: L.iexa_last_error(), 5000)[0])
a = torch.zeros(1024, device="cuda")
print("torch tiny kernel (a.add_(1)): issue %.2f us/call, issue+drain %.2f" % loop(lambda: a.add_(1), 5000))
