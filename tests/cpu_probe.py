import os, sys, time, subprocess
sys.path.insert(0, '.')
print('nproc', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for f in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us'):
    try: print(f, open(f).read().strip())
    except Exception as e: pass
code = '''
import sys, time, os; sys.path.insert(0,'.')
import numpy as np
from iexa_b200 import models
from oracle.oracle import OracleModel
core=models.quadrotor(20000,'oc'); om=OracleModel(core)
rng=np.random.default_rng(0); x=core.x0_vec+0.1*rng.uniform(-1,1,core.nvar); y=rng.uniform(-1,1,core.ncon)
om.cons(x); om.jac_coord(x); om.hess_coord(x,y,1.0)
t0=time.perf_counter()
for _ in range(5): om.cons(x); om.jac_coord(x); om.hess_coord(x,y,1.0)
print('threads', os.environ['OMP_NUM_THREADS'], 's/eval', (time.perf_counter()-t0)/5)
'''
for t in (1, 2, 4, 8, 16):
    subprocess.run([sys.executable, '-c', code], env=dict(os.environ, OMP_NUM_THREADS=str(t)))
