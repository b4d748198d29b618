"""Multi-GPU parity check of the sharded path on real GPUs (NCCL over NVLink), one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tests/dist_gpu_check.py

Every rank evaluates its support blocks with the CUDA engine; obj and the shared-variable slice of grad! go through an
NCCL all-reduce; the global c / Jacobian / Hessian values are assembled from the ranks' slices and compared with the
oracle on rank 0 (tolerance 1e-12 relative / 1e-14 absolute).  Not a pytest file: the driver's `-m gpu` run has one GPU
(the single-GPU sharding test lives in tests/test_gpu_parity.py, the host logic in tests/test_dist_gloo.py)."""
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
    cases = {"ode_5x5": models.ode_5x5, "quadrotor_oc_1000": lambda: models.quadrotor(1000, "oc"),
             "pandemic_200x8": lambda: models.pandemic(200, 8), "farmer_5000": lambda: models.farmer(5000)}
    ok = True
    for name, build in cases.items():
        core = build()
        sm = ShardedExaModel(core, device=local)
        assert sm.model.cmeta.n_kernels_specialised > 0
        x, y = eval_point(core, seed=4)
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
        if rank == 0:
            from oracle.oracle import OracleModel
            om = OracleModel(core)
            try:
                assert_close(f, om.obj(x), "obj (all-reduce)")
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
