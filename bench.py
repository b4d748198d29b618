#!/usr/bin/env python
"""bench.py — cons! + jac_coord! + hess_coord! throughput on the BASELINE.json workload.

One "step" = one eval = one call of each of the three callbacks at the same (x, y) on the
ESCAPE34/quadrotor.jl transcription (OrthogonalCollocation(3), piecewise-constant controls) with
10^6 public time supports (BASELINE.json configs[2]; the configuration the north-star evals/s and
roofline target is quoted on; it fits one B200).  With --gpus N the SAME model is sharded by
contiguous support blocks over N ranks (strong scaling, no data-path collective: every rank owns
its rows / Jacobian slots / Hessian slots).

The ONE JSON line (rank 0) also carries, next to the headline:
  * ``products``   — per-call ms and roofline of the fused jprod! / jtprod! / hprod! kernels;
  * ``iteration``  — obj + grad! + cons! + jac_coord! + hess_coord! + 2 x COO->CSR, device-resident; for N > 1 with the
                     x halo exchange and the NCCL all-reduces (objective, shared gradient slice) INSIDE the timed step
                     (the analogue of the reference's ``ad_time``, ESCAPE34/utils.jl:7,23);
  * ``workloads``  — BASELINE configs 2 / 4 / 5 at their named sizes (pandemic 10^5 x 4 and 100 x 128, OPF case3 x 10^5 and
                     118-bus x 10^4, farmer 10^5): per-callback ms, algorithmic bytes, roofline fraction, and a parity
                     flag computed in this run against the oracle; with --gpus N the same block for the sharded models.

`--impl reference` times the CPU restatement of the reference's evaluator (oracle/, all host threads) on the FULL named
configuration — ExaModels.jl itself cannot be installed here (no Julia, no network).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cons+jac+hess evals/s"
UNIT = "evals/s"
RTOL, ATOL = 1e-12, 1e-14          # north-star tolerance (BASELINE.json)


def _models():
    from iexa_b200 import models
    return models


def _opf(nbus, K):
    from iexa_b200 import opf
    from iexa_b200.transform import exa_core
    return exa_core(opf.opf(None if nbus == 3 else opf.synthetic_grid(nbus), num_supports=K))[0]


HEADLINES = {
    # BASELINE.json configs[2]: the configuration the north-star evals/s / roofline / 8-GPU targets are quoted on (default)
    "quadrotor": dict(name="ESCAPE34/quadrotor.jl OC(3), 10^6 time supports: cons!+jac_coord!+hess_coord!", supports=1_000_000,
                      build=lambda n: _models().quadrotor(n, "oc")),
    # BASELINE.json configs[1]: SEIR optimal control, 10^5 time supports x 4 scenarios (small: a few waves per kernel)
    "pandemic": dict(name="ESCAPE34/pandemic.jl, 10^5 time supports x 4 scenarios: cons!+jac_coord!+hess_coord!", supports=100_000,
                     build=lambda n: _models().pandemic(n, 4)),
}
# the other BASELINE configurations at their named sizes: the `workloads` block of the JSON line
WORKLOADS = {
    "pandemic_1e5x4": dict(config="configs[1] ESCAPE34/pandemic.jl:4-34, 10^5 time supports x 4 scenarios", build=lambda: _models().pandemic(100_000, 4)),
    "pandemic_100x128": dict(config="configs[1] ESCAPE34/pandemic.jl at the largest case of the reference's study grid (run_cases_gpu.jl:100): 100 time supports x 128 scenarios",
                             build=lambda: _models().pandemic(100, 128)),
    "opf_case3_1e5": dict(config="configs[3] ESCAPE34/opf.jl:36-286, embedded 3-bus case x 10^5 scenarios", build=lambda: _opf(3, 100_000)),
    "opf_118bus_1e4": dict(config="configs[3] ESCAPE34/opf.jl:36-286, synthetic 118-bus grid (2832 generators -> shape-class kernels) x 10^4 scenarios",
                           build=lambda: _opf(118, 10_000)),
    "farmer_1e5": dict(config="configs[4] examples/2stage_example.jl:20-37, 10^5 scenarios (LP: empty Hessian)", build=lambda: _models().farmer(100_000)),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="quadrotor", choices=sorted(HEADLINES), help="headline configuration (default: BASELINE configs[2])")
    ap.add_argument("--supports", type=int, default=None, help="public time supports (default: the named size of the workload)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the `workloads` block (configs 2 / 4 / 5)")
    ap.add_argument("--no-iteration", action="store_true")
    ap.add_argument("--no-products", action="store_true")
    ap.add_argument("--only-workloads", default=None, help="comma-separated subset of the workloads block")
    ap.add_argument("--interp", action="store_true", help="AOT tape-interpreter kernels only (no NVRTC)")
    args = ap.parse_args()
    w = HEADLINES[args.workload]
    args.supports = w["supports"] if args.supports is None else args.supports
    args.build, args.workload_name = w["build"], w["name"]
    return args


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def eval_point(core, seed=0):
    """x = x0 + 0.1 U(-1,1), y ~ U(-1,1)  """
    rng = np.random.default_rng(seed)
    x = np.where(np.isfinite(core.x0_vec), core.x0_vec, 0.0) + 0.1 * rng.uniform(-1, 1, core.nvar)
    y = rng.uniform(-1, 1, core.ncon)
    return x, y


def workload_config(name, supports, nvar, ncon, nnzj, nnzh):
    """identical in the b200 arm and the reference arm: what is computed, not how"""
    out_gb = 8 * (ncon + nnzj + nnzh) / 1e9
    return {"workload": name, "supports": int(supports), "nvar": int(nvar), "ncon": int(ncon), "nnzj": int(nnzj), "nnzh": int(nnzh),
            "l2": "inputs larger than L2 (x = %.0f MB, outputs %.1f GB per eval): no flush between iterations" % (nvar * 8 / 1e6, out_gb)
                  if 8 * nvar > 126e6 else "working set near the 126 MB L2: a 512 MB buffer is rewritten between timed iterations (L2 flush)"}


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region.  A 20-step timed region lasts 10 ms — shorter than nvidia-smi's fastest
    sampling period (and a longer warm-up to wait for it runs the GPU into its power cap) — so the clocks are polled in-process
    through NVML every millisecond by a thread (the launching thread sits in cudaStreamSynchronize without the GIL);
    `nvidia-smi -lms` (B200_PROFILING.md recipe) is the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.rows, self.samples = [], []
        self.proc, self.nvml, self.h = None, None, None
        self.gpu = gpu_index
        self._stop = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices: honour CUDA_VISIBLE_DEVICES
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                try:
                    rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                util = int(n.nvmlDeviceGetUtilizationRates(self.h).gpu)
                self.samples.append((time.perf_counter(), mhz, rs, util))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        return time.perf_counter()

    def stop(self, t0=None, t1=None):
        if self.nvml is not None:
            self._stop = True
            self.t.join(timeout=1)
            sel = [s for s in self.samples if (t0 is None or s[0] >= t0) and (t1 is None or s[0] <= t1)] or self.samples
            reasons = set()
            for s in sel:
                for bit, name in self.REASONS.items():
                    if s[2] & bit:
                        reasons.add(name)
            return {"sm_mhz": float(np.median([s[1] for s in sel])) if sel else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(reasons), "samples": len(sel), "samples_total": len(self.samples),
                    "how": "NVML polled in-process every ms; samples between the start of the timed region and the end of the per-kernel pass"}
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
                "reasons": sorted(reasons), "samples": len(sm), "how": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference's evaluator) on the FULL configuration
# ---------------------------------------------------------------------------------------------------------
def oracle_eval_seconds(om, x, y, threads, reps, warm=1):
    from oracle import oracle as orc
    orc.set_threads(threads)
    for _ in range(warm):
        om.cons(x); om.jac_coord(x); om.hess_coord(x, y, 1.0)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        om.cons(x); om.jac_coord(x); om.hess_coord(x, y, 1.0)
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args):
    """the reference arm: EVERY step is one full-size eval (cons + jac_coord + hess_coord at the named 10^6 supports) of the
    oracle with all host threads; nothing is sampled or scaled"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from oracle.oracle import OracleModel
    threads = host_threads()
    orc.set_threads(threads)
    core = args.build(args.supports)
    om = OracleModel(core)
    x, y = eval_point(core)
    for _ in range(args.warmup):
        om.cons(x); om.jac_coord(x); om.hess_coord(x, y, 1.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        om.cons(x); om.jac_coord(x); om.hess_coord(x, y, 1.0)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    v = 1.0 / dt
    sample = (f"oracle (C restatement of ExaModels' per-support recursive AD, OpenMP over supports, {threads} threads) on the FULL "
              f"configuration: every step evaluates all {args.supports} public supports, no sampling, no scaling")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload_name, args.supports, core.nvar, om.ncon, om.nnzj, om.nnzh),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "ExaModels.jl (the reference's evaluator) is not installable offline: no Julia, no network; the port is a tree-walking "
                "restatement, not Julia-compiled code",
    }))


# ---------------------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ---------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(gpu_index: int):
    """one process per GPU: run this rank — and therefore allocate its pinned host buffers — on the NUMA node its GPU hangs off
    (what `numactl --cpunodebind --membind` does in a deployment): the host-buffer path (`e2e`) otherwise pulls half of the
    ranks' PCIe traffic across the socket interconnect.  Best effort; returns the node or None."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]            # sysfs uses a 4-digit domain
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


class Ctx:
    """torch / distributed plumbing of one rank"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa = bind_to_gpu_numa_node(self.local_rank) if self.world > 1 else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self._flush = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor(np.atleast_1d(np.asarray(v, dtype=np.float64)), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def sum_over_ranks(self, v):
        t = self.torch.tensor(np.atleast_1d(np.asarray(v, dtype=np.float64)), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def flush_l2(self):
        """rewrite a buffer larger than the 126 MB L2"""
        if self._flush is None:
            self._flush = self.torch.empty(512 * 1024 * 1024 // 8, dtype=self.torch.float64, device=self.dev)
        self._flush.fill_(1.0)

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)


def local_slices(ex, m, which):
    n = m.L.iexa_segments(m.h, which, None, 0)
    segs = (ex.lib.Segment * max(n, 1))()
    m.L.iexa_segments(m.h, which, segs, n)
    return [(s.global_start, s.local_start, s.length) for s in segs[:n]]


def to_local(ex, m, which, glob, n_local):
    out = np.zeros(max(n_local, 1))
    for gs, ls, ln in local_slices(ex, m, which):
        out[ls:ls + ln] = glob[gs:gs + ln]
    return out


def close(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all((np.abs(a - b) <= ATOL + RTOL * np.maximum(np.abs(a), np.abs(b))) | (np.isnan(a) & np.isnan(b))))


def parity_vs_oracle(ctx, ex, m, core, x_h, y_full, bufs):
    """every rank checks ITS slices of c / Jacobian / Hessian values against the oracle's global vectors (computed by rank 0
    with all host threads and broadcast); returns the flags AND-ed over the ranks"""
    torch, dist = ctx.torch, ctx.dist
    from oracle.oracle import OracleModel
    refs = None
    if ctx.rank == 0:
        from oracle import oracle as orc
        orc.set_threads(host_threads())
        om = OracleModel(core)
        refs = [om.cons(x_h), om.jac_coord(x_h), om.hess_coord(x_h, y_full, 1.0)]
        dims_ok = (om.ncon, om.nnzj, om.nnzh) == (m.meta.ncon, m.meta.nnzj, m.meta.nnzh)
    if ctx.world > 1:
        sizes = [m.meta.ncon, m.meta.nnzj, m.meta.nnzh]
        got = []
        for i, n in enumerate(sizes):
            t = torch.from_numpy(refs[i]).to(ctx.dev) if ctx.rank == 0 else torch.empty(max(n, 0), dtype=torch.float64, device=ctx.dev)
            if n:
                dist.broadcast(t, src=0)
            got.append(t.cpu().numpy())
            del t
        refs = got
        dims_ok = True
    flags = {}
    for name, which, buf, nloc in (("cons", 0, bufs[0], m.loc_ncon), ("jac_coord", 1, bufs[1], m.loc_nnzj), ("hess_coord", 2, bufs[2], m.loc_nnzh)):
        loc = buf.cpu().numpy()[:nloc]
        ref = to_local(ex, m, which, refs[which], nloc)[:nloc]
        flags[name] = close(loc, ref)
    flags["dims"] = bool(dims_ok)
    ok = np.array([float(all(flags.values()))] + [float(flags[k]) for k in ("cons", "jac_coord", "hess_coord", "dims")])
    if ctx.world > 1:
        t = torch.tensor(ok, dtype=torch.float64, device=ctx.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = t.cpu().numpy()
    return {"all": bool(ok[0]), "cons": bool(ok[1]), "jac_coord": bool(ok[2]), "hess_coord": bool(ok[3]), "dims": bool(ok[4]),
            "tolerance": f"|a-b| <= {ATOL:g} + {RTOL:g}*max(|a|,|b|) vs the oracle on the same x, y (sigma = 1), every local slot of every rank"}


def per_callback_ms(ctx, calls, reps, flush):
    """mean duration of every call in `calls` (CUDA events on the launching stream around each launch), max over ranks"""
    torch = ctx.torch
    for _ in range(3):
        for f in calls:
            f()
    ctx.barrier()
    acc = np.zeros(len(calls))
    for _ in range(reps):
        for j, f in enumerate(calls):
            if flush:
                ctx.flush_l2()
            a, b = ctx.ev(), ctx.ev()
            a.record(); f(); b.record()
            b.synchronize()
            acc[j] += a.elapsed_time(b)
    return ctx.max_over_ranks(acc / reps)


def measure_workload(ctx, ex, name, spec, peak, reps=20):
    """one entry of the `workloads` block"""
    torch = ctx.torch
    from iexa_b200.model import bind
    t0 = time.perf_counter()
    core = spec["build"]()
    t_core = time.perf_counter() - t0
    t0 = time.perf_counter()
    m = ex.ExaModel(core, device=ctx.local_rank, rank=ctx.rank, world=ctx.world)
    t_plan = time.perf_counter() - t0
    x_h, y_full = eval_point(core, seed=11)
    y_h = to_local(ex, m, 0, y_full, m.loc_ncon)
    x, y = torch.from_numpy(x_h).to(ctx.dev), torch.from_numpy(y_h).to(ctx.dev)
    z = lambda n: torch.zeros(max(int(n), 1), dtype=torch.float64, device=ctx.dev)
    c, jv, hv, g = z(m.loc_ncon), z(m.loc_nnzj), z(m.loc_nnzh), z(m.meta.nvar)
    names = ["cons", "jac_coord", "hess_coord", "grad"]
    which = [ex.lib.CB_CONS, ex.lib.CB_JAC, ex.lib.CB_HESS, ex.lib.CB_GRAD]
    calls = [bind(m, "cons", x, c), bind(m, "jac_coord", x, jv), bind(m, "hess_coord", x, hv, y, 1.0), bind(m, "grad", x, g)]
    for f in calls:
        f()
    ctx.barrier()
    parity = parity_vs_oracle(ctx, ex, m, core, x_h, y_full, (c, jv, hv))
    # L2: these working sets are near or below the 126 MB L2 — rewrite a 512 MB buffer before every timed launch
    ms = per_callback_ms(ctx, calls, reps, flush=True)
    # back to back, no flush, one event pair around the three callbacks (what a solver iteration sees)
    ctx.barrier()
    a, b = ctx.ev(), ctx.ev()
    a.record()
    for _ in range(reps):
        calls[0](); calls[1](); calls[2]()
    b.record(); ctx.barrier()
    warm_ms = float(ctx.max_over_ranks(a.elapsed_time(b) / reps)[0])
    byt = np.array([float(ex.algorithmic_bytes(m, w)) for w in which])
    byt = ctx.sum_over_ranks(byt)                 # whole job: the ranks' bytes add up
    per = {}
    for n, t, bts in zip(names, ms, byt):
        gbs = bts / (t * 1e-3) / 1e9 if t > 0 else 0.0
        per[n] = {"ms": float(t), "bytes": int(bts), "GB/s": gbs, "frac": gbs / (peak * ctx.world)}
    tot_ms, tot_b = float(ms[:3].sum()), float(byt[:3].sum())
    out = {"config": spec["config"], "nvar": int(m.meta.nvar), "ncon": int(m.meta.ncon), "nnzj": int(m.meta.nnzj), "nnzh": int(m.meta.nnzh),
           "generators": int(m.cmeta.nobj_gen + m.cmeta.ncon_gen), "kernels": f"nvrtc-specialised ({m.cmeta.n_kernels_specialised})" if m.cmeta.n_kernels_specialised else "interpreter",
           "engine_note": m.L.iexa_engine_note(m.h).decode()[:120], "n_gpus": ctx.world,
           "per_callback": per,
           "cons+jac+hess": {"ms": tot_ms, "evals/s": 1e3 / tot_ms, "bytes": int(tot_b), "GB/s": tot_b / (tot_ms * 1e-3) / 1e9,
                             "frac": tot_b / (tot_ms * 1e-3) / 1e9 / (peak * ctx.world),
                             "timing": f"sum of the three per-launch means ({reps} launches each, CUDA events, L2 flushed before every launch, max over ranks)"},
           "cons+jac+hess_back_to_back": {"ms": warm_ms, "evals/s": 1e3 / warm_ms,
                                          "note": "no flush between launches: part of the working set stays in the 126 MB L2 (what a solver iteration sees; not a roofline number)"},
           "parity": parity, "build_s": {"lowering": t_core, "plan+upload+nvrtc": t_plan}}
    del m, x, y, c, jv, hv, g, calls
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import iexa_b200 as ex
    from iexa_b200.model import bind
    ctx = Ctx()
    torch, dist = ctx.torch, ctx.dist
    world, rank, local_rank, dev = ctx.world, ctx.rank, ctx.local_rank, ctx.dev
    barrier = ctx.barrier

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    t0 = time.perf_counter()
    core = args.build(args.supports)
    t_core = time.perf_counter() - t0
    flags = ex.lib.IEXA_F_NO_SPECIALISE if args.interp else ex.lib.IEXA_F_DEFAULT
    t0 = time.perf_counter()
    m = ex.ExaModel(core, device=local_rank, rank=rank, world=world, flags=flags)
    t_plan = time.perf_counter() - t0
    cstat = (C.c_int32(), C.c_int32())
    m.L.iexa_debug_cache_stats(C.byref(cstat[0]), C.byref(cstat[1]))
    # the same build again: the compiled image now comes from the cache (what every later solve of a parameter study, or a
    # run with a warm on-disk cache, pays): host plan + upload only
    t0 = time.perf_counter()
    m2 = ex.ExaModel(core, device=local_rank, rank=rank, world=world, flags=flags)
    t_rebuild = time.perf_counter() - t0
    del m2
    x_h, y_full = eval_point(core)
    y_h = to_local(ex, m, 0, y_full, m.loc_ncon)     # this rank's multipliers, in its local row layout
    x = torch.from_numpy(x_h).to(dev)
    y = torch.from_numpy(y_h).to(dev)
    c = torch.empty(max(m.loc_ncon, 1), dtype=torch.float64, device=dev)
    jv = torch.empty(max(m.loc_nnzj, 1), dtype=torch.float64, device=dev)
    hv = torch.empty(max(m.loc_nnzh, 1), dtype=torch.float64, device=dev)

    # ---- device memory per rank (inputs sharded, not just outputs); EVERY rank takes part in the max
    db = ex.device_bytes(m)
    nxr = m.L.iexa_x_ranges(m.h, None, 0)
    xsegs = (ex.lib.Segment * max(nxr, 1))()
    m.L.iexa_x_ranges(m.h, xsegs, nxr)
    xr = 8 * sum(sg.length for sg in xsegs[:nxr]) if world > 1 else 8 * m.meta.nvar
    caller_local = 8 * (m.loc_ncon * 2 + m.loc_nnzj + m.loc_nnzh)           # y, c, Jacobian values, Hessian values of this rank
    mine = [db["columns"] + db["theta"] + db["programs_tables"], caller_local, xr, 8 * m.meta.nvar]
    mx = [float(v) for v in ctx.max_over_ranks([float(v) for v in mine])]
    whole = 8 * (core.ncon * 2 + m.meta.nnzj + m.meta.nnzh) + db["columns_unsharded"] + db["theta_unsharded"]
    device_memory = {
        "engine_bytes_per_rank_max": int(mx[0]), "engine_columns": db["columns"], "engine_columns_unsharded": db["columns_unsharded"],
        "engine_theta": db["theta"], "engine_theta_unsharded": db["theta_unsharded"], "engine_programs_tables": db["programs_tables"],
        "caller_y_c_jac_hess_bytes_per_rank_max": int(mx[1]), "x_bytes_touched_per_rank_max": int(mx[2]), "x_bytes_addressed": int(mx[3]),
        "per_rank_over_unsharded": (mx[0] + mx[1] + mx[2]) / float(whole + 8 * m.meta.nvar),
        "note": "engine = iterator columns (the slice this rank's supports visit) + theta (full-length virtual range, only this rank's 2 MB granules "
                "backed) + programs / tables; caller = y, c and the value arrays of the rank's rows; x is addressed full-length by contract "
                "(global indices) — a rank touches x_bytes_touched of it (iexa_x_ranges); per_rank_over_unsharded counts the touched part"}

    # buffers are bound once (raw pointers + stream), like a Julia ccall on CuArray pointers: every
    # callback below is exactly one call into the C ABI
    f_cons = bind(m, "cons", x, c)
    f_jac = bind(m, "jac_coord", x, jv)
    f_hess = bind(m, "hess_coord", x, hv, y, 1.0)

    def step():
        f_cons(); f_jac(); f_hess()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    nwarm = max(args.warmup, 3)
    for _ in range(nwarm):
        step()
    barrier()
    t_mark0 = sampler.mark()
    # (1) the timed region: EXACTLY K steps between two events on the launching (current torch) stream —
    #     nothing else is enqueued between the callbacks, as in a solver iteration
    e0, e1 = ctx.ev(), ctx.ev()
    barrier()
    t_wall = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        f_cons(); f_jac(); f_hess()
    e1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    total_ms = e0.elapsed_time(e1)
    # (2) per-callback breakdown for the roofline: a second live pass with an event around every launch
    #     (an event between two kernels keeps the next one from being launched ahead — PDL — so this pass
    #     is a little slower than (1); its per-kernel durations are what the roofline is quoted on)
    nb = max(3, min(args.steps, 100))
    ev = [[ctx.ev() for _ in range(4)] for _ in range(nb)]
    for i in range(nb):
        ev[i][0].record(); f_cons()
        ev[i][1].record(); f_jac()
        ev[i][2].record(); f_hess()
        ev[i][3].record()
    barrier()
    clocks = sampler.stop(t_mark0, sampler.mark()) if rank == 0 else None
    per = np.array([[e[j].elapsed_time(e[j + 1]) for j in range(3)] for e in ev]).mean(axis=0)  # ms
    per = ctx.max_over_ranks(per)
    ms_per_step = float(ctx.max_over_ranks(total_ms)[0]) / args.steps
    value = 1e3 / ms_per_step
    names = ["cons", "jac_coord", "hess_coord"]
    which = [ex.lib.CB_CONS, ex.lib.CB_JAC, ex.lib.CB_HESS]
    bytes_loc = np.array([float(ex.algorithmic_bytes(m, w)) for w in which])
    bytes_cb = ctx.sum_over_ranks(bytes_loc)
    launches_step = int(sum(ex.launches_per_call(m, w) for w in which))
    bytes_meta.update(nnzj=int(m.meta.nnzj), nnzh=int(m.meta.nnzh), nspec=int(m.cmeta.n_kernels_specialised))

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
        dt = float(ctx.max_over_ranks((time.perf_counter() - t0) / n_e2e)[0])
        h2d = int(m.L.iexa_host_x_bytes(m.h)) + 8 * m.loc_ncon
        d2h = 8 * (m.loc_ncon + m.loc_nnzj + m.loc_nnzh)
        e2e = {"value": 1.0 / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": n_e2e, "note": "host (pinned) buffers through iexa_cons/iexa_jac_coord/iexa_hess_coord; PCIe copies inside the timed region; only the "
                                       "parts of x this rank READS are uploaded (iexa_x_ranges), once per eval (new_x); y once; c + Jacobian + Hessian values downloaded; "
                                       "bytes are per rank"}
        del xp, yp, cp, jp, hp, h_cons, h_jac, h_hess

    # ---- matrix-free products: fused kernels (no COO values materialised), per-call ms + roofline
    products = None
    if not args.no_products:
        rng = np.random.default_rng(5)
        v = torch.from_numpy(rng.uniform(-1, 1, core.nvar)).to(dev)
        w_full = rng.uniform(-1, 1, core.ncon)
        w = torch.from_numpy(to_local(ex, m, 0, w_full, m.loc_ncon)).to(dev)
        Jv = torch.empty(max(m.loc_ncon, 1), dtype=torch.float64, device=dev)
        Jtw = torch.empty(m.meta.nvar, dtype=torch.float64, device=dev)
        Hv = torch.empty(m.meta.nvar, dtype=torch.float64, device=dev)
        pcalls = [bind(m, "jprod", x, Jv, v=v), bind(m, "jtprod", x, Jtw, v=w), bind(m, "hprod", x, Hv, y, 1.0, v=v)]
        pms = per_callback_ms(ctx, pcalls, max(5, min(args.steps, 50)), flush=False)
        pb = ctx.sum_over_ranks([float(ex.algorithmic_bytes(m, wch)) for wch in (ex.lib.CB_JPROD, ex.lib.CB_JTPROD, ex.lib.CB_HPROD)])
        products = {}
        for n, t, bts, wch in zip(("jprod", "jtprod", "hprod"), pms, pb, (ex.lib.CB_JPROD, ex.lib.CB_JTPROD, ex.lib.CB_HPROD)):
            gbs = bts / (t * 1e-3) / 1e9
            products[n] = {"ms": float(t), "bytes": int(bts), "GB/s": gbs, "frac": gbs / (peak * world), "launches": int(ex.launches_per_call(m, wch))}
        products["note"] = ("fused first / second order programs with a product epilogue; algorithmic bytes = the inputs those programs load + the touched part "
                            "of v (and y) + every entry of the result once; jprod! has no atomics (bit-reproducible); jtprod!/hprod!: single-writer blocks stored "
                            "directly in a first launch, the remaining contributions added with atomics in a second, ordered launch"
                            + ("; world > 1: per-rank partial sums, no all-reduce in this number" if world > 1 else ""))
        products["engine_note"] = m.L.iexa_engine_note(m.h).decode()[:120]
        del v, w, Jv, Jtw, Hv, pcalls

    # ---- the fused entry point: cons! + jac_coord! + hess_coord! at one (x, y) in ONE launch (iexa_eval3).  NOT the headline
    #      (the metric counts the three NLPModels calls); reported next to it
    eval3 = None
    if not args.no_products:
        f3 = bind(m, "eval3", x, (c, jv, hv), y, 1.0)
        for _ in range(3):
            f3()
        barrier()
        n3 = max(5, min(args.steps, 50))
        a, b = ctx.ev(), ctx.ev()
        a.record()
        for _ in range(n3):
            f3()
        b.record(); barrier()
        ms3 = float(ctx.max_over_ranks(a.elapsed_time(b) / n3)[0])
        b3 = float(ctx.sum_over_ranks(float(ex.algorithmic_bytes(m, ex.lib.CB_EVAL3)))[0])
        eval3 = {"ms": ms3, "evals/s": 1e3 / ms3, "bytes": int(b3), "GB/s": b3 / (ms3 * 1e-3) / 1e9, "frac": b3 / (ms3 * 1e-3) / 1e9 / (peak * world),
                 "launches": 1, "note": "iexa_eval3: one fused kernel — every constraint group evaluates value, first and second order from ONE program (x / theta / "
                                        "columns loaded once, sin / cos of a state once); algorithmic bytes = distinct inputs once + c + Jacobian + Hessian values; "
                                        "results identical to the three callbacks (tests/test_gpu_parity.py)"}

    # ---- solver iteration: obj + grad! + cons! + jac_coord! + hess_coord! + COO->CSR of both matrices, device-resident,
    #      with (N > 1) the x halo exchange and the all-reduces INSIDE the timed step (ESCAPE34/utils.jl:7,23 `ad_time`)
    iteration = None
    if not args.no_iteration:
        iteration = measure_iteration(ctx, ex, m, core, x, y, c, jv, hv, args, peak)

    # ---- NON-TARGET cost, reported separately: getting a new iterate to the ranks when the solver lives
    #      on GPU 0: NCCL broadcast of the whole x, vs the halo exchange a distributed solver does (inside `iteration`)
    xdist = None
    if world > 1:
        def timed(fn, reps=11):
            for _ in range(4):
                fn()
            ts = []
            for _ in range(reps):
                barrier()
                a, b = ctx.ev(), ctx.ev()
                a.record(); fn(); b.record()
                barrier()
                ts.append(float(ctx.max_over_ranks(a.elapsed_time(b))[0]))
            return float(np.median(ts))
        n = m.L.iexa_x_ranges(m.h, None, 0)
        segs = (ex.lib.Segment * max(n, 1))()
        m.L.iexa_x_ranges(m.h, segs, n)
        xdist = {"broadcast_whole_x_ms": timed(lambda: dist.broadcast(x, src=0)), "x_bytes": int(8 * m.meta.nvar),
                 "read_fraction_per_rank": sum(s.length for s in segs[:n]) / m.meta.nvar,
                 "note": "non-target: a solver that lives on GPU 0 broadcasts x; a distributed solver exchanges the shared slice and the shard-boundary halos only (timed inside `iteration`)"}

    # ---- the other BASELINE configurations at their named sizes
    workloads = None
    if not args.no_workloads:
        del m, x, y, c, jv, hv, f_cons, f_jac, f_hess
        torch.cuda.empty_cache()
        workloads = {}
        only = set(args.only_workloads.split(",")) if args.only_workloads else None
        for wname, spec in WORKLOADS.items():
            if only and wname not in only:
                continue
            try:
                workloads[wname] = measure_workload(ctx, ex, wname, spec, peak)
            except Exception as e:   # one broken workload must not take the headline down
                workloads[wname] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
        m = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (largest share of the step), measured live above
    dom = int(np.argmax(per))
    achieved = bytes_cb[dom] / (per[dom] * 1e-3) / 1e9
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf) and world == 1:
        try:
            traffic = json.load(open(tf)).get(names[dom])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": f"iexa_cb_{names[dom].split('_')[0]}", "achieved": achieved, "peak": peak * world,
                "unit": "GB/s", "frac": achieved / (peak * world), "traffic": traffic, "peak_kind": peak_kind + (f" x {world} GPUs" if world > 1 else ""),
                "algorithmic_bytes_per_launch": int(bytes_cb[dom]), "avg_launch_ms": float(per[dom]),
                "per_callback": {n: {"ms": float(t), "GB/s": b / (t * 1e-3) / 1e9, "frac": b / (t * 1e-3) / 1e9 / (peak * world),
                                     "bytes": int(b)} for n, t, b in zip(names, per, bytes_cb)},
                "all_three": {"bytes": int(sum(bytes_cb)), "GB/s": sum(bytes_cb) / (ms_per_step * 1e-3) / 1e9,
                              "frac": sum(bytes_cb) / (ms_per_step * 1e-3) / 1e9 / (peak * world),
                              "note": "whole step of the timed region (two events around K steps)"},
                "timing": f"per-kernel durations: CUDA events around every launch in a second live pass of {nb} steps "
                          "right after the timed region (events between the callbacks would serialise the "
                          "programmatic dependent launches the timed region uses); bytes summed and times max-ed over the ranks"}

    cpu = None
    if not args.no_cpu and world == 1:
        # the FULL named configuration, no sampling: one eval on one thread (ExaModels' CPU backend is sequential: Julia
        # threads = 1), then the same with every host thread (OpenMP over supports)
        from oracle.oracle import OracleModel
        om = OracleModel(core)
        nthr = host_threads()
        tall = oracle_eval_seconds(om, x_h, y_full, nthr, reps=3, warm=1)
        t1 = oracle_eval_seconds(om, x_h, y_full, 1, reps=1, warm=0)
        cpu = {"value": 1.0 / min(t1), "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"oracle (C restatement of ExaModels' sequential per-support AD) on the FULL configuration ({args.supports} public supports): "
                         f"one eval on 1 thread = {min(t1):.2f} s (ExaModels' CPU backend is sequential: Julia threads = 1); no sampling, no scaling",
               "all_cores": {"value": 1.0 / min(tall), "unit": UNIT, "cores": nthr, "seconds_per_eval": min(tall),
                             "sample": "same model, OpenMP over supports, best of 3 full evals"}}
        del om

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload_name, args.supports, core.nvar, core.ncon, bytes_meta["nnzj"], bytes_meta["nnzh"]),
        "engine": {"sharding": f"contiguous support blocks x{world}", "kernels": "interpreter" if args.interp else f"nvrtc-specialised ({bytes_meta['nspec']})",
                   "build_s": {"lowering": t_core, "plan+upload+nvrtc": t_plan, "plan+upload (image cached)": t_rebuild,
                               "nvrtc_compiles": int(cstat[0].value), "disk_cache_hits": int(cstat[1].value),
                               "note": "compiled images are cached in memory and on disk keyed by the generated source (IEXA_CACHE_DIR)"},
                   "warmup_steps_run": int(nwarm), "numa_node_of_rank0": ctx.numa},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "eval3": eval3, "products": products, "iteration": iteration,
        "x_distribution": xdist, "workloads": workloads, "device_memory": device_memory,
        "gpu_launches": int(args.steps * launches_step),
        "clocks": clocks, "wall_s_timed_region": t_wall,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


bytes_meta = {}


def measure_iteration(ctx, ex, m, core, x, y, c, jv, hv, args, peak):
    """obj + grad! + cons! + jac_coord! + hess_coord! + iexa_csr_apply(J) + iexa_csr_apply(H) per step, device-resident.
    N > 1: every step starts with the x halo exchange (each rank owns a part of the iterate; shared variables and shard
    boundaries are pushed over NVLink) and contains the all-reduce of the objective and of the shared gradient slice."""
    torch, dist = ctx.torch, ctx.dist
    from iexa_b200.model import bind
    from iexa_b200.dist import ShardedExaModel
    dev, world = ctx.dev, ctx.world
    L = m.L
    st = torch.cuda.current_stream(dev).cuda_stream
    g = torch.empty(m.meta.nvar, dtype=torch.float64, device=dev)
    f_dev = torch.zeros(1, dtype=torch.float64, device=dev)
    sm = ShardedExaModel.wrap(m)
    peer = sm.enable_peer_halo(x) if world > 1 else False
    # CSR handles of THIS rank's COO slices (rows are global numbers; a rank's CSR has entries in its own rows only)
    csr = []
    t0 = time.perf_counter()
    for whichm, nnz, fn, nrows in ((0, m.loc_nnzj, ex.jac_structure_, m.meta.ncon), (1, m.loc_nnzh, ex.hess_structure_, m.meta.nvar)):
        if nnz == 0:
            csr.append(None)
            continue
        r = torch.zeros(nnz, dtype=torch.int32, device=dev); cc = torch.zeros_like(r)
        fn(m, r, cc)
        keys = torch.zeros(nnz, dtype=torch.int32, device=dev)
        ex.lib.check(L, L.iexa_coo_locality(m.h, whichm, keys.data_ptr(), 1, st))
        torch.cuda.synchronize()
        h = C.c_void_p()
        ex.lib.check(L, L.iexa_csr_create_keyed(C.byref(h), nrows, m.meta.nvar, nnz, r.data_ptr(), cc.data_ptr(), 4,
                                                keys.data_ptr() if whichm == 1 else None, 1, ctx.local_rank))
        outv = torch.empty(L.iexa_csr_nnz(h), dtype=torch.float64, device=dev)
        csr.append((h, outv))
        del r, cc, keys
    t_csr = time.perf_counter() - t0
    f_grad = bind(m, "grad", x, g)
    f_cons, f_jac, f_hess = bind(m, "cons", x, c), bind(m, "jac_coord", x, jv), bind(m, "hess_coord", x, hv, y, 1.0)
    vp = C.c_void_p
    xp, fp = vp(x.data_ptr()), vp(f_dev.data_ptr())
    srcs = (jv, hv)

    def step():
        if world > 1:
            sm.exchange_x(x)
        L.iexa_obj_device(m.h, xp, fp, vp(st))
        f_grad()
        if world > 1:
            sm.allreduce_obj_grad_(f_dev, g)
        f_cons(); f_jac(); f_hess()
        for k in range(2):
            if csr[k] is not None:
                L.iexa_csr_apply(csr[k][0], vp(srcs[k].data_ptr()), vp(csr[k][1].data_ptr()), 1, vp(st))

    for _ in range(3):
        step()
    ctx.barrier()
    n = max(5, min(args.steps, 50))
    a, b = ctx.ev(), ctx.ev()
    ctx.barrier()
    a.record()
    for _ in range(n):
        step()
    b.record()
    ctx.barrier()
    ms = float(ctx.max_over_ranks(a.elapsed_time(b) / n)[0])
    # components, each timed alone (events around every call; no flush: config 3 streams far more than the L2 holds)
    comp_calls = [lambda: L.iexa_obj_device(m.h, xp, fp, vp(st)), f_grad, f_cons, f_jac, f_hess]
    comp_names = ["obj", "grad", "cons", "jac_coord", "hess_coord"]
    for k, nm in ((0, "csr_apply_jac"), (1, "csr_apply_hess")):
        if csr[k] is not None:
            comp_calls.append(lambda k=k: L.iexa_csr_apply(csr[k][0], vp(srcs[k].data_ptr()), vp(csr[k][1].data_ptr()), 1, vp(st)))
            comp_names.append(nm)
    if world > 1:
        comp_calls += [lambda: sm.exchange_x(x), lambda: sm.allreduce_obj_grad_(f_dev, g)]
        comp_names += ["x_halo_exchange", "allreduce_obj+shared_grad"]
    cms = per_callback_ms(ctx, comp_calls, max(5, min(args.steps, 30)), flush=False)
    byt = [float(ex.algorithmic_bytes(m, w)) for w in (ex.lib.CB_OBJ, ex.lib.CB_GRAD, ex.lib.CB_CONS, ex.lib.CB_JAC, ex.lib.CB_HESS)]
    for k in range(2):
        if csr[k] is not None:
            nnz, cn = (m.loc_nnzj, m.loc_nnzh)[k], int(L.iexa_csr_nnz(csr[k][0]))
            byt.append(float(8 * nnz + 8 * cn + 4 * cn + (0 if cn == nnz and k == 0 else 4 * nnz + 4 * cn)))
    tot_b = float(ctx.sum_over_ranks(sum(byt))[0])
    out = {"ms": ms, "iterations/s": 1e3 / ms, "steps": n, "n_gpus": world,
           "what": "obj + grad! + cons! + jac_coord! + hess_coord! + iexa_csr_apply(Jacobian) + iexa_csr_apply(Hessian), device-resident, one x"
                   + ("; every step starts with the x halo exchange and contains the all-reduce of the objective and of the shared gradient slice" if world > 1 else ""),
           "bytes": int(tot_b), "GB/s": tot_b / (ms * 1e-3) / 1e9, "frac": tot_b / (ms * 1e-3) / 1e9 / (peak * world),
           "components_ms": {nm: float(t) for nm, t in zip(comp_names, cms)},
           "csr_setup_s": t_csr, "halo_entries_this_rank": int(sm.halo_entries()) if world > 1 else 0,
           "collectives": ("x halo exchange and [objective, shared gradient slice] all-reduce as single kernels over NVLink peer memory (csrc/halo.cu, CUDA IPC); "
                           f"shared gradient entries: {'all of g (NCCL)' if sm.shared_all else len(sm.shared_idx)}; peer status {sm.peer_status()}") if peer
                          else ("NCCL point-to-point halo exchange + NCCL all-reduce (torch.distributed)" if world > 1 else "none (one GPU)"),
           "reference_analogue": "ad_time of ESCAPE34/utils.jl:7,23 (time inside the NLPModels callbacks per solve) — per iteration"}
    # ---- the same iteration on the ROW-SORTED slot policy (iexa.h IEXA_SLOT_ORDER_JAC_ROW_SORTED): jac_coord! writes the CSR
    #      value array directly, so the Jacobian's COO->CSR pass is not part of the step at all.  Checked in-run against the
    #      default model: same CSR pattern, bit-identical CSR values.
    # Every rank takes the same path: set-up (which may fail) happens first, the ranks then AGREE whether to run the timed part.
    rs, rs_err = None, None
    want = not getattr(args, "no_row_sorted", False)
    if want and csr[0] is not None:
        try:
            t0 = time.perf_counter()
            mr = ex.ExaModel(core, device=ctx.local_rank, rank=ctx.rank, world=world, slot_order=2,
                             flags=ex.lib.IEXA_F_NO_SPECIALISE if args.interp else ex.lib.IEXA_F_DEFAULT)
            t_build = time.perf_counter() - t0
            if ex.jac_is_csr(mr):
                jr = torch.empty_like(jv)
                rp = torch.zeros(mr.loc_ncon + 1, dtype=torch.int32, device=dev)
                ex.jac_csr_rowptr_(mr, rp)
                rr = torch.zeros(mr.loc_nnzj, dtype=torch.int32, device=dev); cr = torch.zeros_like(rr)
                ex.jac_structure_(mr, rr, cr)
                r_grad = bind(mr, "grad", x, g)
                r_cons, r_jac, r_hess = bind(mr, "cons", x, c), bind(mr, "jac_coord", x, jr), bind(mr, "hess_coord", x, hv, y, 1.0)
                # parity first: the default model's CSR (its own COO->CSR map) against the array jac_coord! wrote here
                f_jac(); L.iexa_csr_apply(csr[0][0], vp(jv.data_ptr()), vp(csr[0][1].data_ptr()), 1, vp(st))
                r_jac()
                torch.cuda.synchronize()
                prp = torch.zeros(m.meta.ncon + 1, dtype=torch.int32, device=dev); pci = torch.zeros(mr.loc_nnzj, dtype=torch.int32, device=dev)
                ex.lib.check(L, L.iexa_csr_pattern(csr[0][0], vp(prp.data_ptr()), vp(pci.data_ptr()), 1))
                torch.cuda.synchronize()
                same_vals = bool(torch.equal(csr[0][1], jr[: csr[0][1].numel()]))
                same_cols = bool(torch.equal(pci, cr - 1))
                same_rows = bool(torch.equal(prp, rp)) if world == 1 else True   # world > 1: the default map numbers rows globally
                rs = (mr, r_grad, r_cons, r_jac, r_hess, same_vals, same_cols, same_rows, t_build)
                del rp, rr, cr, prp, pci
            else:
                rs_err = "a generator of this model has no static column order"
        except Exception as e:      # the variant must never take the headline down
            rs, rs_err = None, repr(e)[:300]
    elif want:
        rs_err = "no Jacobian entries on this rank"
    if want:
        go = float(ctx.max_over_ranks(0.0 if rs is not None else 1.0)[0]) == 0.0
        if go:
            mr, r_grad, r_cons, r_jac, r_hess, same_vals, same_cols, same_rows, t_build = rs

            def step_r():
                if world > 1:
                    sm.exchange_x(x)
                L.iexa_obj_device(mr.h, xp, fp, vp(st))
                r_grad()
                if world > 1:
                    sm.allreduce_obj_grad_(f_dev, g)
                r_cons(); r_jac(); r_hess()
                if csr[1] is not None:
                    L.iexa_csr_apply(csr[1][0], vp(hv.data_ptr()), vp(csr[1][1].data_ptr()), 1, vp(st))

            for _ in range(3):
                step_r()
            ctx.barrier()
            a, b = ctx.ev(), ctx.ev()
            ctx.barrier()
            a.record()
            for _ in range(n):
                step_r()
            b.record()
            ctx.barrier()
            msr = float(ctx.max_over_ranks(a.elapsed_time(b) / n)[0])
            tj = per_callback_ms(ctx, [r_jac], max(5, min(args.steps, 30)), flush=False)
            bad = ctx.max_over_ranks([0.0 if same_vals else 1.0, 0.0 if same_cols else 1.0, 0.0 if same_rows else 1.0])   # AND over the ranks
            out["row_sorted"] = {"ms": msr, "iterations/s": 1e3 / msr, "jac_coord_ms": float(tj[0]),
                                 "csr_values_bit_identical_to_default_policy_plus_csr_apply": bool(bad[0] == 0.0),
                                 "csr_colind_identical": bool(bad[1] == 0.0), "csr_rowptr_identical": bool(bad[2] == 0.0) if world == 1 else None,
                                 "build_s": t_build,
                                 "what": "the same step with IEXA_OPT_SLOT_ORDER = JAC_ROW_SORTED: jac_coord! output IS the CSR value array "
                                         "(iexa_jac_csr_rowptr + jac_structure cols), so only the Hessian goes through iexa_csr_apply"}
        else:
            out["row_sorted"] = {"unavailable": rs_err or "unavailable on another rank"}
        rs = None
    sm.close_peer_halo()
    for k in range(2):
        if csr[k] is not None:
            L.iexa_csr_destroy(csr[k][0])
    del csr, g
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
