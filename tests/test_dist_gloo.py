"""world_size-2 (and 3) tests of the sharded path on CPU with the gloo backend: sharding layout,
the shared-variable gradient slice, the objective / gradient all-reduce and the assembly of the
global vectors — with the per-rank arithmetic supplied by the test-only host executor
(tests/hostcheck) instead of the CUDA engine, and checked against the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class HostEvaluator:
    """hostcheck_eval_local behind the evaluator interface of ShardedExaModel (test only)"""

    def __init__(self, L, model):
        self.L, self.m = L, model

    def _run(self, which, x, y, s, out):
        rc = self.L.hostcheck_eval_local(self.m.h, which, x.ctypes.data, None if y is None else y.ctypes.data, s, out.ctypes.data)
        assert rc == 0
        # every column position / theta entry the rank touched is inside what the CUDA engine keeps resident on that rank
        assert self.L.hostcheck_residency_misses() == 0, "a rank read a column position or a theta entry outside its resident slices"
        return out

    def obj(self, x):
        return float(self._run(0, x, None, 1.0, np.zeros(1))[0])

    def grad_(self, x, g): return self._run(1, x, None, 1.0, g)
    def cons_(self, x, c): return self._run(2, x, None, 1.0, c)
    def jac_coord_(self, x, v): return self._run(3, x, None, 1.0, v)
    def hess_coord_(self, x, y, v, w): return self._run(4, x, y, w, v)


def _case_core(case):
    """named model, or ``fuzz<seed>``: a random model of tests/test_fuzz.py (arbitrary integer columns, product iterators)"""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from iexa_b200 import models
    if case.startswith("fuzz"):
        import test_fuzz
        return test_fuzz.random_model(int(case[4:]), K1=130, K2=4)[0]
    if case in ("probe_shift", "probe_product"):
        return _probe(case)
    return {"ode_5x5": models.ode_5x5, "quadrotor": lambda: models.quadrotor(13, "oc"), "farmer": lambda: models.farmer(23)}[case]()


def _probe(case):
    """objectives whose gradient entries are written by several ranks WITHOUT a constant index (the shared set must come
    from index ranges, not from constant indices): (A) (y[i+1] - y[i])^2 over i = 1..T-1 — the entry at every shard
    boundary is written by both neighbours; (B) w[s] * z[t]^2 over the product iterator (t, s) — every z[t] is written
    by every rank that holds some scenario s."""
    import numpy as np
    from iexa_b200.core import ExaCore, Itr
    from iexa_b200.expr import DataSource, abs2
    core = ExaCore(minimize=True)
    ds = DataSource()
    T = 17
    if case == "probe_shift":
        y = core.add_var(T, start=0.3)
        it = Itr(T - 1, {"group_idx1": np.arange(1, T)}, {"c": np.linspace(1, 2, T - 1)})
        i = ds.group_idx1.idx()
        core.add_obj(ds.c * abs2(y[i + 1] - y[i]), it)
        core.add_con(y[ds.group_idx1], Itr(T, {"group_idx1": np.arange(1, T + 1)}, {}), -1.0, 1.0)
    else:
        S = 5
        z = core.add_var(T, start=0.4)
        it_t = Itr(T, {"group_idx1": np.arange(1, T + 1)}, {})
        it_s = Itr(S, {"group_idx2": np.arange(1, S + 1)}, {"w": np.linspace(0.5, 1.5, S)})
        core.add_obj(ds.w * abs2(z[ds.group_idx1]), Itr.product([it_t, it_s]))
        core.add_con(z[ds.group_idx1], it_t, -1.0, 1.0)
    return core


def _worker(rank, world, port, case, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import iexa_b200 as ex
    from iexa_b200 import models
    from iexa_b200.dist import ShardedExaModel
    from conftest import eval_point
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        L = ex.lib.load(os.path.join(ROOT, "tests", "hostcheck", "libiexa_hostcheck.so"))
        core = _case_core(case)
        sm = ShardedExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
        sm._ev = HostEvaluator(L, sm.model)
        x, y = eval_point(core, seed=4)
        f = sm.obj(x)
        g = sm.grad_(x, np.zeros(core.nvar))
        gfull = sm.grad_full_(x, np.zeros(core.nvar))
        c = sm.cons_(x, np.zeros(max(sm.model.loc_ncon, 1)))
        jv = sm.jac_coord_(x, np.zeros(max(sm.model.loc_nnzj, 1)))
        yl = sm.scatter_local(0, y)
        hv = sm.hess_coord_(x, yl, np.zeros(max(sm.model.loc_nnzh, 1)), 0.7)
        cg = sm.gather_global(0, c).numpy()
        jg = sm.gather_global(1, jv).numpy()
        hg = sm.gather_global(2, hv).numpy()
        # x outside iexa_x_ranges is never read by this rank: poison it and evaluate again
        xp = np.full_like(x, np.nan)
        cover = 0
        for s0, ln in sm.x_ranges():
            xp[s0:s0 + ln] = x[s0:s0 + ln]; cover += ln
        assert world == 1 or case in ("ode_5x5", "probe_product") or case.startswith("fuzz") or cover < core.nvar, "x ranges are not a proper subset"
        c2 = sm.cons_(xp, np.zeros(max(sm.model.loc_ncon, 1)))
        jv2 = sm.jac_coord_(xp, np.zeros(max(sm.model.loc_nnzj, 1)))
        hv2 = sm.hess_coord_(xp, yl, np.zeros(max(sm.model.loc_nnzh, 1)), 0.7)
        assert np.array_equal(c2, c) and np.array_equal(jv2, jv) and np.array_equal(hv2, hv), "a rank read x outside its ranges"
        assert sm._ev.obj(xp) == sm._ev.obj(x)
        # distributed iterate: every rank starts with current values on the ranges it OWNS only; after exchange_x the
        # ranges it reads are current and the evaluation is unchanged
        owned, recv, send = sm.x_partition()
        xo = np.full_like(x, np.nan)
        for lo, hi in owned:
            xo[lo:hi] = x[lo:hi]
        n_owned = torch.tensor([sum(hi - lo for lo, hi in owned)]); dist.all_reduce(n_owned)
        assert int(n_owned) <= core.nvar                         # ownership is disjoint
        sm.exchange_x(xo)
        c3 = sm.cons_(xo, np.zeros(max(sm.model.loc_ncon, 1)))
        jv3 = sm.jac_coord_(xo, np.zeros(max(sm.model.loc_nnzj, 1)))
        assert np.array_equal(c3, c) and np.array_equal(jv3, jv), "halo exchange left a read range stale"
        # map_dual on sharded buffers: the multipliers of ONE constraint (rows of one generator) from the local slices
        gcon = core.cons[min(1, len(core.cons) - 1)]
        d = sm.gather_rows(yl, gcon.row_offset, gcon.itr.K).numpy()
        assert np.array_equal(d, y[gcon.row_offset:gcon.row_offset + gcon.itr.K]), "gather_rows"
        # owned gradient entries: sum over ranks of the per-rank g must double-count ONLY the shared slice
        gsum = torch.from_numpy(g.copy()); dist.all_reduce(gsum)
        shared = np.arange(core.nvar) if sm.shared_all else sm.shared_idx
        # ... and PER RANK: a non-shared entry is either complete on this rank (it is the only contributor) or exactly zero —
        # never a partial sum (ShardedExaModel.grad_'s contract; a rank-sum check alone is blind to that)
        from oracle.oracle import OracleModel
        ref = OracleModel(core).grad(x)
        excl = np.ones(core.nvar, dtype=bool); excl[shared] = False
        tol = 1e-13 + 1e-12 * np.abs(ref)
        whole, zero = np.abs(g - ref) <= tol, g == 0.0
        assert (whole | zero)[excl].all(), f"rank {rank}: partial sums outside the shared set at {np.flatnonzero(excl & ~(whole | zero))[:8]}"
        assert (np.abs(g - ref) <= tol)[shared].all(), f"rank {rank}: shared entries incomplete after grad_"
        if rank == 0:
            q.put(dict(f=f, gfull=gfull, c=cg, j=jg, h=hg, gsum=gsum.numpy(), shared=shared, g=g,
                       loc=(sm.model.loc_ncon, sm.model.loc_nnzj, sm.model.loc_nnzh)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case,world", [("ode_5x5", 2), ("quadrotor", 2), ("quadrotor", 3), ("farmer", 2), ("fuzz7", 3), ("fuzz11", 2),
                                        ("probe_shift", 2), ("probe_shift", 3), ("probe_product", 2)])
def test_sharded_evaluation_matches_oracle(case, world, hostcheck_lib):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from iexa_b200 import models
    from oracle.oracle import OracleModel
    from conftest import assert_close, eval_point
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue as _queue
    res = None
    for _ in range(180):
        try:
            res = q.get(timeout=1)
            break
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0, "a rank failed"
    assert res is not None
    core = _case_core(case)
    om = OracleModel(core)
    x, y = eval_point(core, seed=4)
    assert abs(res["f"] - om.obj(x)) <= 1e-12 * max(1.0, abs(om.obj(x)))
    assert np.allclose(res["gfull"], om.grad(x), rtol=1e-12, atol=1e-13)
    assert_close(res["c"], om.cons(x), "cons")
    assert_close(res["j"], om.jac_coord(x), "jac")
    assert_close(res["h"], om.hess_coord(x, y, 0.7), "hess")
    # rank 0 owns fewer rows than the whole model
    assert res["loc"][0] < om.ncon
    # after grad_, shared entries are complete on every rank; everything else is owned by exactly one rank
    ref = om.grad(x)
    sh = res["shared"]
    assert np.allclose(res["g"][sh], ref[sh], rtol=1e-12, atol=1e-13)
    expect = ref.copy(); expect[sh] *= world
    assert np.allclose(res["gsum"], expect, rtol=1e-12, atol=1e-13)
    if case in ("ode_5x5", "probe_shift", "probe_product"):
        assert len(sh) >= 1  # z is shared by every support; shard-boundary / product-iterator entries
    if case == "farmer":
        assert len(sh) == 0  # the first-stage x enter the objective through length-1 generators only: all on rank 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_a_rank_keeps_one_nth_of_columns_and_theta(world, hostcheck_lib):
    """SURVEY 8(e) / VERDICT item 3: inputs are sharded, not just outputs.  For the quadrotor OC model (three K-long parameter
    functions in theta, support / coefficient / weight columns) the slices a rank keeps resident — Plan::column_read_ranges,
    Plan::theta_read_ranges, what engine.cu uploads — are ~1/world of the model's, every access of the rank's evaluation falls
    inside them (hostcheck_eval_local counts misses), and together the ranks cover everything."""
    sys.path.insert(0, ROOT)
    import ctypes as C
    import iexa_b200 as ex
    from iexa_b200 import models
    from conftest import eval_point
    L = hostcheck_lib
    core = models.quadrotor(400, "oc")
    x, y = eval_point(core, seed=2)
    tot = np.zeros(4, dtype=np.int64)
    for r in range(world):
        m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L, rank=r, world=world)
        b = np.zeros(4, dtype=np.int64)
        assert L.hostcheck_residency_bytes(m.h, b.ctypes.data) == 0
        assert b[1] > 0 and b[3] == 8 * core.npar > 0
        # a slice plus a halo of a few supports per column / theta block
        assert b[0] <= b[1] / world + 64 * 40, (r, b)
        assert b[2] <= b[3] / world + 64 * 8, (r, b)
        tot[:] += b
        for which, n in ((2, m.loc_ncon), (3, m.loc_nnzj), (4, m.loc_nnzh), (1, m.meta.nvar), (0, 1)):
            out = np.zeros(max(n, 1))
            yl = np.zeros(max(m.loc_ncon, 1))
            assert L.hostcheck_eval_local(m.h, which, x.ctypes.data, yl.ctypes.data, 0.7, out.ctypes.data) == 0
            assert L.hostcheck_residency_misses() == 0
    assert tot[0] >= b[1] and tot[2] >= b[3]      # the ranks' slices cover the model
