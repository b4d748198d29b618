#!/usr/bin/env python
"""bench.py — cons! + jac_coord! + hess_coord! throughput on the BASELINE.json workload.

One "step" = one eval = one call of each of the three callbacks at the same (x, y) on the
ESCAPE34/quadrotor.jl transcription (OrthogonalCollocation(3), piecewise-constant controls) with
10^6 public time supports (BASELINE.json configs[2]; the configuration the north-star evals/s and
roofline target is quoted on; it fits one B200).  With --gpus N the SAME model is sharded by
contiguous support blocks over N ranks (strong scaling, no data-path collective: every rank owns
its rows / Jacobian slots / Hessian slots).

`--workload pandemic` runs the same contract on BASELINE.json configs[1] (ESCAPE34/pandemic.jl, 10^5 time supports x 4
scenarios; small: a few waves of blocks per kernel).  Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference's
evaluator (oracle/, all host threads) on a bounded sample of the same workload — ExaModels.jl
itself cannot be installed here (no Julia, no network).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the north-star evals/s / roofline / 8-GPU targets are quoted on (default)
    "quadrotor": dict(name="ESCAPE34/quadrotor.jl OC(3), 10^6 time supports: cons!+jac_coord!+hess_coord!", supports=1_000_000,
                      cpu_sample=20_000, build=lambda n: __import__("iexa_b200").models.quadrotor(n, "oc")),
    # BASELINE.json configs[1]: SEIR optimal control, 10^5 time supports x 4 scenarios (small: a few waves per kernel)
    "pandemic": dict(name="ESCAPE34/pandemic.jl, 10^5 time supports x 4 scenarios: cons!+jac_coord!+hess_coord!", supports=100_000,
                     cpu_sample=20_000, build=lambda n: __import__("iexa_b200").models.pandemic(n, 4)),
}
WORKLOAD = WORKLOADS["quadrotor"]["name"]
METRIC = "cons+jac+hess evals/s"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="quadrotor", choices=sorted(WORKLOADS), help="BASELINE.json configuration (default: configs[2])")
    ap.add_argument("--supports", type=int, default=None, help="public time supports (default: the named size of the workload)")
    ap.add_argument("--cpu-sample", type=int, default=None, help="supports of the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--interp", action="store_true", help="AOT tape-interpreter kernels only (no NVRTC)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    args.supports = w["supports"] if args.supports is None else args.supports
    args.cpu_sample = w["cpu_sample"] if args.cpu_sample is None else args.cpu_sample
    args.build, args.workload_name = w["build"], w["name"]
    return args


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [s.strip() for s in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_eval_rate(build, full_rows: int, sample_supports: int, threads: int, reps: int = 2):
    """evals/s of the oracle (CPU restatement of the reference evaluator) on a bounded sample,
    scaled linearly to the full support count."""
    import iexa_b200 as ex  # noqa: F401  (models only; the oracle does the arithmetic)
    from iexa_b200 import models
    from oracle import oracle as orc
    from oracle.oracle import OracleModel
    orc.set_threads(threads)
    core = build(sample_supports)
    om = OracleModel(core)
    rng = np.random.default_rng(0)
    x = np.where(np.isfinite(core.x0_vec), core.x0_vec, 0.0) + 0.1 * rng.uniform(-1, 1, core.nvar)
    y = rng.uniform(-1, 1, core.ncon)
    om.cons(x); om.jac_coord(x); om.hess_coord(x, y, 1.0)  # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        om.cons(x); om.jac_coord(x); om.hess_coord(x, y, 1.0)
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    scale = full_rows / core.ncon          # rows (and slots) are proportional to the number of supports
    return 1.0 / (t * scale), t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # each "step" is one eval of the bounded sample; K steps after W warm-ups
    from iexa_b200 import models
    from oracle import oracle as orc
    from oracle.oracle import OracleModel
    orc.set_threads(threads)
    ns = args.cpu_sample
    core = args.build(ns)
    om = OracleModel(core)
    rng = np.random.default_rng(0)
    x = np.where(np.isfinite(core.x0_vec), core.x0_vec, 0.0) + 0.1 * rng.uniform(-1, 1, core.nvar)
    y = rng.uniform(-1, 1, core.ncon)
    for _ in range(args.warmup):
        om.cons(x); om.jac_coord(x); om.hess_coord(x, y, 1.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        om.cons(x); om.jac_coord(x); om.hess_coord(x, y, 1.0)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    full = args.build(args.supports) if args.supports <= 200_000 else None
    scale = (full.ncon / core.ncon) if full is not None else (2 * args.supports - 1) / (2 * ns - 1)   # quadrotor: T = 2N - 1 supports
    v = 1.0 / (dt * scale)
    sample = (f"oracle (C restatement of ExaModels' per-support recursive AD, OpenMP over supports) on "
              f"{ns} of {args.supports} public supports, time scaled linearly by {scale:.1f}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * scale * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload_name, "supports": args.supports},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "ExaModels.jl (the reference's evaluator) is not installable offline: no Julia, no network",
    }))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import iexa_b200 as ex
    from iexa_b200 import models

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    core = args.build(args.supports)
    flags = ex.lib.IEXA_F_NO_SPECIALISE if args.interp else ex.lib.IEXA_F_DEFAULT
    m = ex.ExaModel(core, device=local_rank, rank=rank, world=world, flags=flags)
    rng = np.random.default_rng(0)
    x_h = np.where(np.isfinite(core.x0_vec), core.x0_vec, 0.0) + 0.1 * rng.uniform(-1, 1, core.nvar)
    y_full = rng.uniform(-1, 1, core.ncon)
    # this rank's multipliers, in its local row layout
    segs = (ex.lib.Segment * 4096)()
    nseg = m.L.iexa_segments(m.h, 0, segs, 4096)
    y_h = np.zeros(max(m.loc_ncon, 1))
    for s in segs[:nseg]:
        y_h[s.local_start:s.local_start + s.length] = y_full[s.global_start:s.global_start + s.length]
    x = torch.from_numpy(x_h).to(dev)
    y = torch.from_numpy(y_h).to(dev)
    c = torch.empty(max(m.loc_ncon, 1), dtype=torch.float64, device=dev)
    jv = torch.empty(max(m.loc_nnzj, 1), dtype=torch.float64, device=dev)
    hv = torch.empty(max(m.loc_nnzh, 1), dtype=torch.float64, device=dev)

    # buffers are bound once (raw pointers + stream), like a Julia ccall on CuArray pointers: every
    # callback below is exactly one call into the C ABI
    from iexa_b200.model import bind
    f_cons = bind(m, "cons", x, c)
    f_jac = bind(m, "jac_coord", x, jv)
    f_hess = bind(m, "hess_coord", x, hv, y, 1.0)

    def step():
        f_cons(); f_jac(); f_hess()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # (1) the timed region: EXACTLY K steps between two events on the launching (current torch) stream —
    #     nothing else is enqueued between the callbacks, as in a solver iteration
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        f_cons(); f_jac(); f_hess()
    e1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    total_ms = e0.elapsed_time(e1)
    # (2) per-callback breakdown for the roofline: a second live pass with an event around every launch
    #     (an event between two kernels keeps the next one from being launched ahead — PDL — so this pass
    #     is a little slower than (1); its per-kernel durations are what the roofline is quoted on)
    nb = max(3, min(args.steps, 100))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(nb)]
    for i in range(nb):
        ev[i][0].record(); f_cons()
        ev[i][1].record(); f_jac()
        ev[i][2].record(); f_hess()
        ev[i][3].record()
    barrier()
    per = np.array([[e[j].elapsed_time(e[j + 1]) for j in range(3)] for e in ev]).mean(axis=0)  # ms
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = float(tmax.item()) / args.steps
    value = 1e3 / ms_per_step

    # ---- e2e: the same step through the C ABI with HOST buffers (pinned), copies inside the call
    e2e = None
    if not args.no_e2e:
        xp = torch.from_numpy(x_h).pin_memory()
        yp = torch.from_numpy(y_h).pin_memory()
        cp = torch.empty(max(m.loc_ncon, 1), dtype=torch.float64).pin_memory()
        jp = torch.empty(max(m.loc_nnzj, 1), dtype=torch.float64).pin_memory()
        hp = torch.empty(max(m.loc_nnzh, 1), dtype=torch.float64).pin_memory()

        # one eval = cons! at a NEW x, then jac_coord! and hess_coord! at that same x: what Ipopt's new_x flag says
        # (IEXA_MEM_HOST_SAME_X: the engine keeps the device copy of x between the three calls)
        h_cons = bind(m, "cons", xp, cp)
        h_jac = bind(m, "jac_coord", xp, jp, new_x=False)
        h_hess = bind(m, "hess_coord", xp, hp, yp, 1.0, new_x=False)

        def step_host():
            h_cons(); h_jac(); h_hess()

        n_e2e = max(2, min(args.steps, 5))
        step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            step_host()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = 8 * (m.meta.nvar + m.loc_ncon)
        d2h = 8 * (m.loc_ncon + m.loc_nnzj + m.loc_nnzh)
        e2e = {"value": 1.0 / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": n_e2e, "note": "host (pinned) buffers through iexa_cons/iexa_jac_coord/iexa_hess_coord; PCIe copies inside the timed region; x uploaded once per eval (new_x), y once, c + Jacobian + Hessian values downloaded"}

    # ---- NON-TARGET cost, reported separately: getting a new iterate to the ranks when the solver lives
    #      on GPU 0.  (a) NCCL broadcast of the whole x; (b) only the ranges each rank READS (iexa_x_ranges: its own
    #      supports of every variable block + shared variables + halos), packed, sent point to point, unpacked.
    xdist = None
    if world > 1:
        n = m.L.iexa_x_ranges(m.h, None, 0)
        segs = (ex.lib.Segment * max(n, 1))()
        m.L.iexa_x_ranges(m.h, segs, n)
        mine = [(s.global_start, s.length) for s in segs[:n]]
        allr = [None] * world
        dist.all_gather_object(allr, mine)

        def timed(fn, reps=11):
            for _ in range(4):          # first calls set up NCCL point-to-point channels
                fn()
            ts = []
            for _ in range(reps):
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                barrier()
                t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ts.append(float(t.item()))
            return float(np.median(ts))

        def scatter_ranges():
            if rank == 0:
                ops = []
                for r in range(1, world):
                    buf = torch.cat([x[s0:s0 + ln] for s0, ln in allr[r]])
                    ops.append(dist.P2POp(dist.isend, buf, r))
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            else:
                tot = sum(ln for _, ln in mine)
                buf = torch.empty(tot, dtype=torch.float64, device=dev)
                for w in dist.batch_isend_irecv([dist.P2POp(dist.irecv, buf, 0)]):
                    w.wait()
                o = 0
                for s0, ln in mine:
                    x[s0:s0 + ln] = buf[o:o + ln]; o += ln

        from iexa_b200.dist import ShardedExaModel
        sm = ShardedExaModel.wrap(m)
        _, recv, _ = sm.x_partition()
        halo = torch.tensor([sum(hi - lo for v in recv.values() for lo, hi in v)], device=dev)
        dist.all_reduce(halo)
        xdist = {"broadcast_whole_x_ms": timed(lambda: dist.broadcast(x, src=0)),
                 "scatter_read_ranges_ms": timed(scatter_ranges),
                 "halo_exchange_ms": timed(lambda: sm.exchange_x(x)), "halo_entries_all_ranks": int(halo.item()),
                 "x_bytes": int(8 * m.meta.nvar), "read_fraction_per_rank": sum(ln for _, ln in mine) / m.meta.nvar,
                 "note": "non-target: a solver that lives on GPU 0 broadcasts x or scatters the ranges each rank reads; a distributed solver (each rank owns its part of x, ShardedExaModel.exchange_x) exchanges the shared slice and the shard-boundary halos only"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (largest share of the step), measured live above
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    names = ["cons", "jac_coord", "hess_coord"]
    which = [ex.lib.CB_CONS, ex.lib.CB_JAC, ex.lib.CB_HESS]
    bytes_cb = [ex.algorithmic_bytes(m, w) for w in which]
    dom = int(np.argmax(per))
    achieved = bytes_cb[dom] / (per[dom] * 1e-3) / 1e9
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get(names[dom])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": f"iexa_cb_{names[dom].split('_')[0]}", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                "algorithmic_bytes_per_launch": int(bytes_cb[dom]), "avg_launch_ms": float(per[dom]),
                "per_callback": {n: {"ms": float(t), "GB/s": b / (t * 1e-3) / 1e9, "frac": b / (t * 1e-3) / 1e9 / peak,
                                     "bytes": int(b)} for n, t, b in zip(names, per, bytes_cb)},
                "all_three": {"bytes": int(sum(bytes_cb)), "GB/s": sum(bytes_cb) / (ms_per_step * 1e-3) / 1e9,
                              "frac": sum(bytes_cb) / (ms_per_step * 1e-3) / 1e9 / peak,
                              "note": "whole step of the timed region (two events around K steps)"},
                "timing": f"per-kernel durations: CUDA events around every launch in a second live pass of {nb} steps "
                          "right after the timed region (events between the callbacks would serialise the "
                          "programmatic dependent launches the timed region uses)"}

    cpu = None
    if not args.no_cpu and world == 1:
        v1, t1 = cpu_eval_rate(args.build, int(m.meta.ncon), args.cpu_sample, 1)
        cpu = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"oracle (C restatement of ExaModels' sequential per-support AD) on {args.cpu_sample} of "
                         f"{args.supports} public supports ({t1:.2f} s/eval measured), scaled linearly; ExaModels' "
                         f"CPU backend is sequential (Julia threads = 1)"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload_name, "supports": args.supports, "nvar": int(m.meta.nvar), "ncon": int(m.meta.ncon),
                   "nnzj": int(m.meta.nnzj), "nnzh": int(m.meta.nnzh), "sharding": f"contiguous support blocks x{world}",
                   "l2": "inputs larger than L2 (x = %.0f MB, outputs %.1f GB per eval)" % (m.meta.nvar * 8 / 1e6, (m.loc_nnzj + m.loc_nnzh + m.loc_ncon) * 8 / 1e9),
                   "kernels": "interpreter" if args.interp else f"nvrtc-specialised ({m.cmeta.n_kernels_specialised})"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "x_distribution": xdist,
        "gpu_launches": int(args.steps * sum(ex.launches_per_call(m, w) for w in which)),
        "clocks": clocks, "wall_s_timed_region": t_wall,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
