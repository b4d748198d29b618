"""where the time of building a plan goes (scratch): per C-ABI entry point, on the GPU box"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iexa_b200 as ex
from iexa_b200 import models, lib as _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
t0 = time.perf_counter(); core = models.quadrotor(N, "oc"); t_model = time.perf_counter() - t0
L = _lib.load()
acc = {}
class P:
    def __init__(s, L): s.L = L
    def __getattr__(s, n):
        f = getattr(s.L, n)
        if not n.startswith("iexa_"): return f
        def g(*a):
            t = time.perf_counter(); r = f(*a); acc[n] = acc.get(n, 0) + time.perf_counter() - t; return r
        return g
t = time.perf_counter()
m = ex.ExaModel(core, device=0, library=P(L))
tot = time.perf_counter() - t
print(f"N={N}: python model {t_model:.2f}s, ExaModel() {tot:.2f}s")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:8]: print(f"  {k:24s}{v:.3f}s")
t = time.perf_counter(); m2 = ex.ExaModel(core, device=0, library=P(L)); print(f"second build (cubin cached): {time.perf_counter()-t:.2f}s")
