"""Multi-GPU parity check of the sharded path on real GPUs (NCCL over NVLink), one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tests/dist_gpu_check.py

Every rank evaluates its support blocks with the CUDA engine; obj and the shared-variable slice of grad! go through an
NCCL all-reduce; the global c / Jacobian / Hessian values are assembled from the ranks' slices and compared with the
oracle on rank 0 (tolerance 1e-12 relative / 1e-14 absolute).  The x halo exchange and the objective / shared-gradient
all-reduce are run BOTH ways: NCCL (torch.distributed) and the NVLink peer-memory kernels of csrc/halo.cu.
tests/test_dist_gpu.py runs this script under torch.distributed.run as a `-m gpu` test when >= 2 GPUs are visible."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import iexa_b200 as ex
    from iexa_b200 import models
    from iexa_b200.dist import ShardedExaModel
    from conftest import assert_close, eval_point
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from iexa_b200 import opf
    from iexa_b200.transform import exa_core
    cases = {"ode_5x5": models.ode_5x5, "quadrotor_oc_1000": lambda: models.quadrotor(1000, "oc"),
             "pandemic_200x8": lambda: models.pandemic(200, 8), "farmer_5000": lambda: models.farmer(5000),
             "opf_case3_500": lambda: exa_core(opf.opf(None, num_supports=500))[0]}
    ok = True
    for name, build in cases.items():
        core = build()
        sm = ShardedExaModel(core, device=local)
        assert sm.model.cmeta.n_kernels_specialised > 0
        x, y = eval_point(core, seed=4)
        x = np.where(np.isfinite(x), x, 0.0)
        xd, yd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
        f = sm.obj(xd)
        g = sm.grad_(xd, torch.zeros(core.nvar, dtype=torch.float64, device=dev))
        gfull = sm.grad_full_(xd, torch.zeros(core.nvar, dtype=torch.float64, device=dev))
        c = sm.cons_(xd, torch.zeros(max(sm.model.loc_ncon, 1), dtype=torch.float64, device=dev))
        jv = sm.jac_coord_(xd, torch.zeros(max(sm.model.loc_nnzj, 1), dtype=torch.float64, device=dev))
        yl = sm.scatter_local(0, yd)
        hv = sm.hess_coord_(xd, yl, torch.zeros(max(sm.model.loc_nnzh, 1), dtype=torch.float64, device=dev), 0.7)
        cg, jg, hg = sm.gather_global(0, c), sm.gather_global(1, jv), sm.gather_global(2, hv)
        # x outside iexa_x_ranges is never read by this rank: poison it and evaluate again (bit-identical results)
        xp = torch.full_like(xd, float("nan"))
        cover = 0
        for s0, ln in sm.x_ranges():
            xp[s0:s0 + ln] = xd[s0:s0 + ln]; cover += ln
        c2 = sm.cons_(xp, torch.zeros_like(c)); jv2 = sm.jac_coord_(xp, torch.zeros_like(jv))
        hv2 = sm.hess_coord_(xp, yl, torch.zeros_like(hv), 0.7)
        assert torch.equal(c2, c) and torch.equal(jv2, jv) and torch.equal(hv2, hv), f"rank {rank} read x outside its ranges"
        frac = cover / core.nvar
        # distributed iterate: current values on the OWNED ranges only, then the halo exchange over NCCL point-to-point
        owned, recv, send = sm.x_partition()
        xo = torch.full_like(xd, float("nan"))
        for lo, hi in owned:
            xo[lo:hi] = xd[lo:hi]
        sm.exchange_x(xo)
        c3 = sm.cons_(xo, torch.zeros_like(c)); hv3 = sm.hess_coord_(xo, yl, torch.zeros_like(hv), 0.7)
        assert torch.equal(c3, c) and torch.equal(hv3, hv), f"rank {rank}: halo exchange left a read range stale"
        halo_t = torch.tensor([sum(hi - lo for v in recv.values() for lo, hi in v)], device=dev)
        dist.all_reduce(halo_t)
        halo = int(halo_t.item())
        # the same exchange as ONE kernel per rank over NVLink peer memory (csrc/halo.cu), three epochs in a row with a
        # changing iterate: every epoch must deliver the owners' CURRENT values
        xq = torch.full_like(xd, float("nan"))
        peer = sm.enable_peer_halo(xq)
        assert peer, "CUDA IPC peer mapping failed"
        for epoch in range(3):
            xe = xd * (1.0 + 0.25 * epoch)
            xq.fill_(float("nan"))
            for lo, hi in owned:
                xq[lo:hi] = xe[lo:hi]
            sm.exchange_x(xq)
            ce = sm.cons_(xe, torch.zeros_like(c)); cq = sm.cons_(xq, torch.zeros_like(c))
            je = sm.jac_coord_(xe, torch.zeros_like(jv)); jq = sm.jac_coord_(xq, torch.zeros_like(jv))
            assert torch.equal(ce, cq) and torch.equal(je, jq), f"rank {rank}: peer halo exchange left a read range stale (epoch {epoch})"
        assert sm.peer_status() == 0, f"rank {rank}: a bounded wait expired: {sm.peer_status()}"
        # obj + shared gradient slice in ONE peer-memory all-reduce: deterministic and bit-identical on all ranks
        m = sm.model
        f_dev = torch.zeros(1, dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        import ctypes as C
        ex.lib.check(m.L, m.L.iexa_obj_device(m.h, C.c_void_p(xd.data_ptr()), C.c_void_p(f_dev.data_ptr()), C.c_void_p(st)))
        g2 = torch.zeros(core.nvar, dtype=torch.float64, device=dev)
        ex.grad_(m, xd, g2)
        sm.allreduce_obj_grad_(f_dev, g2)
        fa = [torch.zeros_like(f_dev) for _ in range(world)]
        dist.all_gather(fa, f_dev)
        assert all(torch.equal(fa[0], t) for t in fa), "peer all-reduce: objective differs between ranks"
        f_peer = float(f_dev.item())
        shi = torch.from_numpy(sm.shared_idx).to(dev)
        assert sm.shared_all or torch.allclose(g2[shi], g[shi], rtol=1e-13, atol=1e-14), "peer all-reduce: shared gradient slice"
        # matrix-free products on shards: J v rows are owned; J' w and H v are per-rank partial sums (all-reduce)
        rng = np.random.default_rng(9)
        v = rng.uniform(-1, 1, core.nvar); w = rng.uniform(-1, 1, core.ncon)
        vd = torch.from_numpy(v).to(dev); wl = sm.scatter_local(0, torch.from_numpy(w).to(dev))
        Jv = ex.jprod_(m, xd, vd, torch.zeros(max(m.loc_ncon, 1), dtype=torch.float64, device=dev))
        Jvg = sm.gather_global(0, Jv)
        Jtw = ex.jtprod_(m, xd, wl, torch.full((core.nvar,), 7.0, dtype=torch.float64, device=dev)); dist.all_reduce(Jtw)
        Hv = ex.hprod_(m, xd, yl, vd, torch.full((core.nvar,), 7.0, dtype=torch.float64, device=dev), 0.7); dist.all_reduce(Hv)
        sm.close_peer_halo()
        if rank == 0:
            from oracle.oracle import OracleModel
            om = OracleModel(core)
            try:
                assert_close(f, om.obj(x), "obj (all-reduce)")
                assert_close(f_peer, om.obj(x), "obj (peer-memory all-reduce)")
                assert_close(Jvg.cpu().numpy(), om.jprod(x, v), "jprod")
                assert_close(Jtw.cpu().numpy(), om.jtprod(x, w), "jtprod")
                assert_close(Hv.cpu().numpy(), om.hprod(x, y, v, 0.7), "hprod")
                ref = om.grad(x)
                assert np.allclose(gfull.cpu().numpy(), ref, rtol=1e-12, atol=1e-13), "grad (full all-reduce)"
                sh = sm.shared_idx
                assert np.allclose(g.cpu().numpy()[sh], ref[sh], rtol=1e-12, atol=1e-13), "grad (shared slice)"
                assert_close(cg.cpu().numpy(), om.cons(x), "cons")
                assert_close(jg.cpu().numpy(), om.jac_coord(x), "jac_coord")
                assert_close(hg.cpu().numpy(), om.hess_coord(x, y, 0.7), "hess_coord")
                print(f"{name}: world={world} OK  (rank0 owns {sm.model.loc_ncon}/{om.ncon} rows, "
                      f"{len(sh)} shared gradient entries, reads {100 * frac:.1f}% of x, {halo} entries cross ranks in the halo exchange)", flush=True)
            except AssertionError as e:
                ok = False
                print(f"{name}: world={world} FAILED: {e}", flush=True)
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
